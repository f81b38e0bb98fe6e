// Implicit-GEMM convolution on tcgen05 tensor cores (sm_100a).
//
// One persistent, warp-specialised kernel serves every dense contraction of the
// two hot paths: 3x3 spatial conv, temporal Conv1d k3 (+ fused 1x1 skip conv as
// extra K), 1x1 qkv / proj, and the policy's Conv1d k5/k3/k1 and Linear layers.
//
//   rows (M)  = 128 output pixels: a 4-D box of the channels-last output grid
//   cols (N)  = block_n output channels
//   K         = a "tap program": for each tap (source tensor, coordinate
//               offset) a run of 64-channel chunks.  The A tile of one K step
//               is ONE TMA box load at (pixel box + tap offset): out-of-bounds
//               coordinates are zero-filled by the TMA unit, which is exactly
//               the convolution's zero padding, so no im2col buffer exists.
//
// Precision: operands are stored as bf16 (hi, lo) pairs; each K step issues
// hi*hi + lo*hi + hi*lo into the same fp32 TMEM accumulator (error ~2^-16,
// i.e. fp32-class; see DESIGN.md "precision").  passes==1 keeps only hi*hi.
//
// Pipeline (per CTA, 256 threads):
//   warp 0 lane 0 : TMA producer  (smem ring: full/empty mbarriers)
//   warp 1 lane 0 : tcgen05.mma issuer, accumulators double-buffered in TMEM
//   warp 2        : TMEM alloc / dealloc
//   warps 4..7    : epilogue (tcgen05.ld -> bias/rowvec/residual -> global,
//                   optional hi/lo split output, optional GroupNorm partial sums)
#include "common.cuh"
#include "../../include/v2a_b200.h"

#include <atomic>
#include <cstdarg>
#include <cstring>
#include <mutex>
#include <string>

namespace v2a {

// ---------------------------------------------------------------------------
// host error state
// ---------------------------------------------------------------------------
static thread_local std::string g_err;
void set_error(const char* fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    g_err = buf;
}
const char* get_error() { return g_err.c_str(); }
std::atomic<int64_t> g_launches{0};

// ---------------------------------------------------------------------------
// kernel parameters
// ---------------------------------------------------------------------------
constexpr int kThreads = 384;                               // 4 control warps + 8 epilogue warps
constexpr int kEpilogueThreads = 256;
constexpr int kTileM = 128;
constexpr int kChunkK = 64;                                 // bf16 elements per K step (128 B rows)
constexpr int kATileBytes = kTileM * kChunkK * 2;           // 16 KB
constexpr int kTmemCols = 512;
constexpr int kAccStride = 256;                             // TMEM columns per accumulator stage
constexpr int kSlabStride = 20;                             // words per staged row: 16 + 4 (bank spread, 16 B aligned)

struct alignas(64) IgemmParams {
    CUtensorMap a_hi[V2A_MAX_SRC];
    CUtensorMap a_lo[V2A_MAX_SRC];
    CUtensorMap b_hi;
    CUtensorMap b_lo;
    CUtensorMap bh_hi;    // cluster mode: box of block_n/2 rows (each CTA of the pair loads one half, multicast)
    CUtensorMap bh_lo;
    int tap_src[V2A_MAX_TAPS];
    int tap_d[V2A_MAX_TAPS][4];
    int tap_chunks[V2A_MAX_TAPS];
    int ntaps;
    int k_iters;          // sum of tap_chunks
    int tile_log2[4];
    int out_dims[4];
    int ntile[4];         // tiles along each output dim
    int num_m_tiles, num_n_tiles;
    int block_n, passes, stages;
    uint32_t stage_bytes, b_tile_bytes;
    int cout, ldc;
    float* out_f32;
    __nv_bfloat16* out_hi;
    __nv_bfloat16* out_lo;
    const float* bias;
    const float* rowvec;
    int ld_rowvec;
    int rowvec_mul[4];
    const float* residual;
    int ld_res;
    double* stats;
    int stats_mul[4];
    int stats_ld;
    int stats_replicas;
    long long stats_rep_stride;
    int cluster;          // 2: CTA pairs share every B (weight) tile through TMA multicast
    int iters_per_cta;    // cluster mode: tile iterations of every CTA (ghost tiles pad the last ones)
    int k_splits;         // >1: the K loop of one output tile is shared by k_splits CTAs (fp32 atomics)
    long long split_stride;  // > 0 with k_splits > 1: split s stores its partial sums at out_f32 + s * split_stride
    int kps;              // K iterations per split
    int a_fp16, b_fp16;   // operand planes hold fp16 (hi, lo) pairs instead of bf16 ones
    int nacc_log2;        // accumulator ring: 2 x 256 TMEM columns (1) or 4 x 128 (2: block_n <= 128, unfused)
    int cta2;             // CTA pairs issue ONE cta_group::2 MMA (M = 256, B split between the two SMs)
    int fuse2;            // N <= 128: a_hi x [b_hi | b_lo] as ONE N = 2*block_n MMA (two accumulator halves)
    long long out_mul[4]; // output row of grid point c = out_off + sum c[d] * out_mul[d] (dense by default; the
    long long out_off;    // sub-pixel phases of the upsample conv write every other row / column of a finer grid)
    int debug;  // V2A_IGEMM_DEBUG bits: 1 skip stats, 2 skip stores, 4 skip residual (timing experiments only)
    int linear; // tile m = rows [128 m, 128 m + 128) of the dense row space, every tile full, one embedding row and one
                // GroupNorm instance per tile: the epilogue's per-tile set-up needs no votes / shuffles / row decode
};

__device__ __forceinline__ void decode_tile(const IgemmParams& p, int tile, int& n_idx, int o[4], int& split,
                                            int* m_out = nullptr) {
    if (m_out) {   // M-tile index (>= num_m_tiles for a ghost tile)
        *m_out = p.cluster > 1 ? ((tile >> 1) / p.num_n_tiles) * 2 + (tile & 1)
                               : (tile >= p.num_m_tiles * p.num_n_tiles * p.k_splits ? p.num_m_tiles
                                                                                      : tile / p.k_splits / p.num_n_tiles);
    }
    if (p.cluster > 1) {
        // CTA pairs: tiles 2q and 2q + 1 are the two M tiles of pair-tile q, which share one N tile (so the pair
        // shares / splits ONE weight tile); q walks the N tiles fastest.  (k_splits == 1 in pair mode.)
        const int q = tile >> 1;
        n_idx = q % p.num_n_tiles;
        int m = (q / p.num_n_tiles) * 2 + (tile & 1);
        split = 0;
        if (m >= p.num_m_tiles) {     // ghost tile: a box wholly outside the tensor -> TMA zero fill, no valid row
            o[0] = o[1] = o[2] = 0;
            o[3] = p.ntile[3] << p.tile_log2[3];
            return;
        }
#pragma unroll
        for (int d = 0; d < 4; ++d) {
            int j = m % p.ntile[d];
            m /= p.ntile[d];
            o[d] = j << p.tile_log2[d];
        }
        return;
    }
    if (tile >= p.num_m_tiles * p.num_n_tiles * p.k_splits) {
        // ghost tile (cluster mode pads every CTA to the same iteration count): a box wholly outside the
        // tensor -> TMA zero fill, no row is valid in the epilogue
        n_idx = 0;
        split = 0;
        o[0] = o[1] = o[2] = 0;
        o[3] = p.ntile[3] << p.tile_log2[3];
        return;
    }
    split = tile % p.k_splits;     // splits of one output tile are adjacent: they run concurrently
    tile /= p.k_splits;
    n_idx = tile % p.num_n_tiles;
    int m = tile / p.num_n_tiles;
#pragma unroll
    for (int d = 0; d < 4; ++d) {
        int j = m % p.ntile[d];
        m /= p.ntile[d];
        o[d] = j << p.tile_log2[d];
    }
}

// Tiles of one CTA (tile = first + i * step).
struct TileRange {
    int first, step, count;
};
__device__ __forceinline__ TileRange cta_tiles(const IgemmParams& p, int total) {
    TileRange r;
    // round-robin: the CTAs of the grid work on neighbouring tiles at the same time, so conv halos and the
    // three temporal taps of a pixel block are re-read from L2, not from HBM
    r.first = blockIdx.x;
    r.step = gridDim.x;
    if (p.cluster > 1) r.count = p.iters_per_cta;   // CTA pairs stay in lockstep (ghost tiles pad the tail)
    else r.count = total > r.first ? (total - r.first + r.step - 1) / r.step : 0;
    return r;
}

#ifndef V2A_IGEMM_MAXNREG
#define V2A_IGEMM_MAXNREG 168
#endif
// kCta2: the cta_group::2 (CTA-pair MMA) paths exist only in that instantiation -- a kernel that contains
// cta_group::2 instructions must be launched as a cluster of 2
//
// kEpi: what the epilogue writes, fixed at compile time (bit 0: fp32 output, bit 1: hi/lo planes, bit 2: GroupNorm
// sums; such launches have no split-K and no debug switches) or -1 = decided at run time.  ncu source counters
// (round 2) showed the run-time version executing ~900 warp instructions per 16-column chunk, ~150 of which do the
// work -- the rest re-evaluate feature tests, predicates and 64-bit addresses per chunk (IMAD + ISETP + BRA = 36 %
// of everything executed) -- and with only the 8 epilogue warps resident that instruction stream, not TMEM, shared
// memory or HBM, bounded every short-K launch (8.2 us per 128 x 128 tile against 2.5 us of MMA for a temporal conv).
template <bool kCta2, int kEpi>
__global__ void __maxnreg__(V2A_IGEMM_MAXNREG) igemm_kernel(const __grid_constant__ IgemmParams p) {
    constexpr bool kFixed = kEpi >= 0;
    extern __shared__ uint8_t smem_raw[];
    // SWIZZLE_128B tiles need 1024-byte alignment.  The offset is added to the __shared__ pointer itself: a
    // round trip through uintptr_t loses the address space and turned every epilogue slab access into a
    // generic LD/ST (89 LD + 31 ST and not one LDS/STS in the SASS of the previous version).
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int S = p.stages;

    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)S * p.stage_bytes);
    uint64_t* full_bar = bars;             // [S]
    uint64_t* empty_bar = bars + S;        // [S]
    uint64_t* tfull_bar = bars + 2 * S;    // [4]
    uint64_t* tempty_bar = bars + 2 * S + 4;  // [4]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * S + 8);
    float* warp_add = reinterpret_cast<float*>(bars) + 64;  // 8 epilogue warps x 256 floats, after the 256 B barrier block
    float* stage_slab = warp_add + 8 * 256;   // 8 epilogue warps x [32 rows][kSlabStride] coalescing slabs

    pdl_trigger();
    if (warp == 0 && lane == 0) {
        for (int s = 0; s < p.stages; ++s) {
            mbar_init(&full_bar[s], 1);
            // multicast pairs: both CTAs' MMA warps release a stage; cta_group::2 pairs: the leader's commit does
            mbar_init(&empty_bar[s], (p.cluster > 1 && !kCta2) ? 2 : 1);
        }
        for (int s = 0; s < 4; ++s) {
            mbar_init(&tfull_bar[s], 1);
            mbar_init(&tempty_bar[s], kCta2 ? 2 * kEpilogueThreads : kEpilogueThreads);   // cta2: both epilogues
        }
        fence_mbar_init();
    } else if (warp == 1 && lane == 0) {
        tma_prefetch_desc(&p.a_hi[0]);
        tma_prefetch_desc(&p.b_hi);
        if (p.passes == 3) {
            tma_prefetch_desc(&p.a_lo[0]);
            tma_prefetch_desc(&p.b_lo);
        }
    } else if (warp == 2) {
        if constexpr (kCta2) tmem_alloc_2sm(tmem_slot, kTmemCols);
        else tmem_alloc(tmem_slot, kTmemCols);
    }
    tc_fence_before();
    __syncthreads();
    if (p.cluster > 1) cluster_sync_all();   // the peer's barriers exist before anything is multicast at them
    tc_fence_after();
    // programmatic dependent launch: everything above (barriers, TMEM, descriptor prefetch) overlaps the previous
    // kernel's tail; nothing below may run before that kernel's results are visible
    pdl_wait();
    const uint32_t tmem_base = *tmem_slot;
    const int nacc_mask = (1 << p.nacc_log2) - 1;
    const uint32_t acc_stride = kTmemCols >> p.nacc_log2;

    const int total_tiles = p.num_m_tiles * p.num_n_tiles * p.k_splits;
    const TileRange tr = cta_tiles(p, total_tiles);
    const uint32_t crank = p.cluster > 1 ? cluster_ctarank() : 0;

    if (warp == 0 && lane == 0) {
        // ===================== TMA producer =====================
        int stage = 0;
        uint32_t phase = 0;
        for (int ti = 0; ti < tr.count; ++ti) {
            const int tile = tr.first + ti * tr.step;
            int n_idx, o[4], split;
            decode_tile(p, tile, n_idx, o, split);
            const int n0 = n_idx * p.block_n;
            const int kb = split * p.kps;
            const int ke = min(kb + p.kps, p.k_iters);
            int kit = 0;
            for (int e = 0; e < p.ntaps; ++e) {
                const int src = p.tap_src[e];
                const int c1 = o[0] + p.tap_d[e][0], c2 = o[1] + p.tap_d[e][1];
                const int c3 = o[2] + p.tap_d[e][2], c4 = o[3] + p.tap_d[e][3];
                if (kit + p.tap_chunks[e] <= kb || kit >= ke) { kit += p.tap_chunks[e]; continue; }
                for (int ch = 0; ch < p.tap_chunks[e]; ++ch, ++kit) {
                    if (kit < kb || kit >= ke) continue;
                    mbar_wait(&empty_bar[stage], phase ^ 1, 100 + stage);
                    uint8_t* st = smem + (size_t)stage * p.stage_bytes;
                    if constexpr (kCta2) {
                        // each CTA stages its own A tile and ITS HALF of the weight rows in its own shared memory;
                        // all bytes of the pair are counted on the leader's barrier (the leader alone issues MMAs)
                        const uint32_t lead_full = mapa_shared(smem_u32(&full_bar[stage]), 0);
                        if (crank == 0) mbar_arrive_expect_tx(&full_bar[stage], 2 * p.stage_bytes);
                        tma_load_5d_2sm(st, &p.a_hi[src], lead_full, ch * kChunkK, c1, c2, c3, c4);
                        tma_load_5d_2sm(st + kATileBytes, &p.a_lo[src], lead_full, ch * kChunkK, c1, c2, c3, c4);
                        const uint32_t half_rows = p.block_n >> 1, half_bytes = p.b_tile_bytes >> 1;
                        uint8_t* sb = st + 2 * kATileBytes;
                        tma_load_2d_2sm(sb, &p.bh_hi, lead_full, kit * kChunkK, n0 + crank * half_rows);
                        tma_load_2d_2sm(sb + half_bytes, &p.bh_lo, lead_full, kit * kChunkK, n0 + crank * half_rows);
                        if (++stage == S) { stage = 0; phase ^= 1; }
                        continue;
                    }
                    mbar_arrive_expect_tx(&full_bar[stage], p.stage_bytes);
                    if (p.passes == 3 && p.cluster > 1) {
                        tma_load_5d(st, &p.a_hi[src], &full_bar[stage], ch * kChunkK, c1, c2, c3, c4);
                        tma_load_5d(st + kATileBytes, &p.a_lo[src], &full_bar[stage], ch * kChunkK,
                                    c1, c2, c3, c4);
                        // this CTA fetches its half of the weight tile for BOTH CTAs of the pair
                        const uint32_t half_rows = p.block_n >> 1, half_bytes = p.b_tile_bytes >> 1;
                        uint8_t* sb = st + 2 * kATileBytes + crank * half_bytes;
                        tma_load_2d_mc(sb, &p.bh_hi, &full_bar[stage], kit * kChunkK, n0 + crank * half_rows, 3);
                        tma_load_2d_mc(sb + p.b_tile_bytes, &p.bh_lo, &full_bar[stage], kit * kChunkK,
                                       n0 + crank * half_rows, 3);
                    } else if (p.passes == 3) {
                        tma_load_5d(st, &p.a_hi[src], &full_bar[stage], ch * kChunkK, c1, c2, c3, c4);
                        tma_load_5d(st + kATileBytes, &p.a_lo[src], &full_bar[stage], ch * kChunkK,
                                    c1, c2, c3, c4);
                        uint8_t* sb = st + 2 * kATileBytes;
                        tma_load_2d(sb, &p.b_hi, &full_bar[stage], kit * kChunkK, n0);
                        tma_load_2d(sb + p.b_tile_bytes, &p.b_lo, &full_bar[stage], kit * kChunkK, n0);
                    } else {
                        tma_load_5d(st, &p.a_hi[src], &full_bar[stage], ch * kChunkK, c1, c2, c3, c4);
                        tma_load_2d(st + kATileBytes, &p.b_hi, &full_bar[stage], kit * kChunkK, n0);
                    }
                    if (++stage == S) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (kCta2 && warp == 1 && lane == 0) {
        // ===================== MMA issuer of a cta_group::2 pair (leader CTA only) =====================
        // M = 256 (this CTA's tile + the peer's), N rows of B split between the two shared memories.  With the
        // fused split product each CTA's weight stage is [b_hi half | b_lo half], so the N = 2*block_n MMA's
        // columns come out as [hi*hi(c<h) | hi*lo(c<h) | hi*hi(c>=h) | hi*lo(c>=h)] (h = block_n/2) and the
        // a_lo x b_hi MMA (N = block_n -> [lo*hi(c<h) | lo*hi(c>=h)]) is aimed at column h, where both of its
        // halves land on accumulators of the same logical columns.  Shared-memory operand reads per SM per k step:
        // 14 KB instead of 20 KB (24 KB unfused).
        if constexpr (kCta2) if (crank == 0) {
            const int hb = p.block_n >> 1;
            const uint32_t idesc2 = umma_idesc_16(256, p.fuse2 ? 2 * p.block_n : p.block_n, p.a_fp16, p.b_fp16);
            const uint32_t idesc1 = umma_idesc_16(256, p.block_n, p.a_fp16, p.b_fp16);
            int stage = 0;
            uint32_t phase = 0;
            for (int it = 0; it < tr.count; ++it) {
                const int acc = it & nacc_mask;
                const uint32_t acc_phase = (it >> p.nacc_log2) & 1;
                mbar_wait(&tempty_bar[acc], acc_phase ^ 1, 200 + acc);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + acc * acc_stride;
                for (int kit = 0; kit < p.k_iters; ++kit) {
                    mbar_wait(&full_bar[stage], phase, 300 + stage);
                    tc_fence_after();
                    const uint32_t sa = smem_u32(smem + (size_t)stage * p.stage_bytes);
                    const uint64_t a_hi = umma_desc_sw128(sa);
                    const uint64_t a_lo = umma_desc_sw128(sa + kATileBytes);
                    const uint64_t b_hi = umma_desc_sw128(sa + 2 * kATileBytes);
                    if (p.fuse2) {
#pragma unroll
                        for (int k = 0; k < 4; ++k)
                            umma_bf16_2sm(d_tmem, a_hi + 2 * k, b_hi + 2 * k, idesc2, (kit | k) != 0);
#pragma unroll
                        for (int k = 0; k < 4; ++k)
                            umma_bf16_2sm(d_tmem + hb, a_lo + 2 * k, b_hi + 2 * k, idesc1, 1);
                    } else {
                        // block_n > 128: three M = 256 MMAs, columns in natural order (this CTA's half | the peer's)
                        const uint64_t b_lo = umma_desc_sw128(sa + 2 * kATileBytes + (p.b_tile_bytes >> 1));
#pragma unroll
                        for (int k = 0; k < 4; ++k)
                            umma_bf16_2sm(d_tmem, a_lo + 2 * k, b_hi + 2 * k, idesc1, (kit | k) != 0);
#pragma unroll
                        for (int k = 0; k < 4; ++k) umma_bf16_2sm(d_tmem, a_hi + 2 * k, b_lo + 2 * k, idesc1, 1);
#pragma unroll
                        for (int k = 0; k < 4; ++k) umma_bf16_2sm(d_tmem, a_hi + 2 * k, b_hi + 2 * k, idesc1, 1);
                    }
                    umma_commit_2sm_mc(&empty_bar[stage], 3);
                    if (kit == p.k_iters - 1) umma_commit_2sm_mc(&tfull_bar[acc], 3);
                    if (++stage == S) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1 && lane == 0) {
        // ===================== MMA issuer =====================
        const uint32_t idesc = umma_idesc_16(kTileM, p.block_n, p.a_fp16, p.b_fp16);
        int stage = 0;
        uint32_t phase = 0;
        for (int it = 0; it < tr.count; ++it) {
            const int tile = tr.first + it * tr.step;
            const int acc = it & nacc_mask;
            const uint32_t acc_phase = (it >> p.nacc_log2) & 1;
            mbar_wait(&tempty_bar[acc], acc_phase ^ 1, 200 + acc);
            tc_fence_after();
            const uint32_t d_tmem = tmem_base + acc * acc_stride;
            const int kb = (tile % p.k_splits) * p.kps;
            const int n_it = min(kb + p.kps, p.k_iters) - kb;
            for (int kit = 0; kit < n_it; ++kit) {
                mbar_wait(&full_bar[stage], phase, 300 + stage);
                tc_fence_after();
                const uint32_t sa = smem_u32(smem + (size_t)stage * p.stage_bytes);
                if (p.passes == 3 && p.fuse2) {
                    // The weight stage holds [b_hi rows | b_lo rows] back to back = one K-major tile of 2*block_n
                    // rows, so a_hi x b_hi and a_hi x b_lo are ONE MMA with N = 2*block_n writing two accumulator
                    // halves (summed in the epilogue), and a_lo x b_hi adds into the first half.  Same FLOPs, but
                    // 2 instead of 3 MMAs per k step and 20 KB instead of 24 KB of shared-memory operand reads
                    // (at N = 128 the three-MMA form reads 128 B/clk = the whole shared-memory bandwidth, which the
                    // epilogue's staging slabs also need).
                    const uint64_t a_hi = umma_desc_sw128(sa);
                    const uint64_t a_lo = umma_desc_sw128(sa + kATileBytes);
                    const uint64_t b_hi = umma_desc_sw128(sa + 2 * kATileBytes);
                    const uint32_t idesc2 = umma_idesc_16(kTileM, 2 * p.block_n, p.a_fp16, p.b_fp16);
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        umma_bf16(d_tmem, a_hi + 2 * k, b_hi + 2 * k, idesc2, (kit | k) != 0);
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        umma_bf16(d_tmem, a_lo + 2 * k, b_hi + 2 * k, idesc, 1);
                } else if (p.passes == 3) {
                    const uint64_t a_hi = umma_desc_sw128(sa);
                    const uint64_t a_lo = umma_desc_sw128(sa + kATileBytes);
                    const uint64_t b_hi = umma_desc_sw128(sa + 2 * kATileBytes);
                    const uint64_t b_lo = umma_desc_sw128(sa + 2 * kATileBytes + p.b_tile_bytes);
                    // small cross terms first, dominant hi*hi last
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        umma_bf16(d_tmem, a_lo + 2 * k, b_hi + 2 * k, idesc, (kit | k) != 0);
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        umma_bf16(d_tmem, a_hi + 2 * k, b_lo + 2 * k, idesc, 1);
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        umma_bf16(d_tmem, a_hi + 2 * k, b_hi + 2 * k, idesc, 1);
                } else {
                    const uint64_t a_hi = umma_desc_sw128(sa);
                    const uint64_t b_hi = umma_desc_sw128(sa + kATileBytes);
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        umma_bf16(d_tmem, a_hi + 2 * k, b_hi + 2 * k, idesc, (kit | k) != 0);
                }
                // smem slot reusable once these MMAs retire (pair mode: the peer multicasts into it too)
                if (p.cluster > 1) umma_commit_mc(&empty_bar[stage], 3);
                else umma_commit(&empty_bar[stage]);
                if (kit == n_it - 1) umma_commit(&tfull_bar[acc]);
                if (++stage == S) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp >= 4) {
        // ===================== epilogue =====================
        const int quad = warp & 3;            // TMEM lane quarter this warp may read
        const int row = quad * 32 + lane;     // row of the 128-row tile
        const int half = (warp - 4) >> 2;     // two warps share a lane quarter and split the columns
        float* addv = warp_add + (warp - 4) * 256;  // this warp's private bias(+rowvec) staging
        const int nch = p.block_n >> 4;
        const int c_begin = half == 0 ? 0 : ((nch + 1) >> 1) << 4;
        const int c_end = half == 0 ? ((nch + 1) >> 1) << 4 : p.block_n;
        // same-address atomic contention is spread over `stats_replicas` copies of the sums
        const bool has_f32 = kFixed ? (kEpi & 1) != 0 : p.out_f32 != nullptr;
        const bool has_hl = kFixed ? (kEpi & 2) != 0 : p.out_hi != nullptr;
        const bool has_stats = kFixed ? (kEpi & 4) != 0 : (p.stats && !(p.debug & 1));
        const bool splitk = kFixed ? false : p.k_splits > 1;
        const int dbg = kFixed ? 0 : p.debug;
        double* const stats = has_stats ? p.stats + (long long)(blockIdx.x % p.stats_replicas) * p.stats_rep_stride : nullptr;
        // Per-warp staging slab [32 rows][16 words + 4 pad].  TMEM hands each lane one ROW (16 consecutive
        // columns); storing that directly makes every 16-byte store instruction touch 32 different 128-byte
        // lines and bounded the whole kernel on the narrow layers.  Staged through the slab, 4 lanes write one
        // row's 64 bytes (8 rows per instruction), the residual is read the same way, and the GroupNorm column
        // sums are plain shared-memory column walks instead of a 31-shuffle tree.
        float* slab = stage_slab + (warp - 4) * (32 * kSlabStride);
        const int srow0 = lane >> 2;          // store phase: row (srow0 + 8 i), 16-byte piece (lane & 3)
        const int piece = lane & 3;
        for (int it = 0; it < tr.count; ++it) {
            const int tile = tr.first + it * tr.step;
            const int acc = it & nacc_mask;
            const uint32_t acc_phase = (it >> p.nacc_log2) & 1;
            int n_idx, o[4], split, mt;
            decode_tile(p, tile, n_idx, o, split, &mt);
            const int n0 = n_idx * p.block_n;
            const bool lead = split == 0;   // bias / rowvec / residual enter the sum exactly once
            bool valid, any_valid, rv_uniform, inst_uniform;
            int rv, inst, rv0, inst0;
            int64_t pix, spix[4];
            bool svalid[4];
            if (p.linear) {
                // dense, contiguous tiling (every large layer): everything about the tile is warp-uniform
                valid = any_valid = mt < p.num_m_tiles;
                pix = (int64_t)mt * kTileM + row;
                rv = inst = 0;
#pragma unroll
                for (int d = 0; d < 4; ++d) {
                    rv += o[d] * p.rowvec_mul[d];
                    inst += o[d] * p.stats_mul[d];
                }
                rv0 = rv;
                inst0 = inst;
                rv_uniform = true;
                inst_uniform = stats != nullptr && valid;
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    spix[i] = (int64_t)mt * kTileM + quad * 32 + srow0 + 8 * i;
                    svalid[i] = valid;
                }
            } else {
            // row -> output pixel
            int r = row, coord[4];
            valid = true;
#pragma unroll
            for (int d = 0; d < 4; ++d) {
                coord[d] = o[d] + (r & ((1 << p.tile_log2[d]) - 1));
                r >>= p.tile_log2[d];
                valid = valid && (coord[d] < p.out_dims[d]);
            }
            pix = p.out_off + coord[0] * p.out_mul[0] + coord[1] * p.out_mul[1] +
                  coord[2] * p.out_mul[2] + coord[3] * p.out_mul[3];
            rv = 0, inst = 0;
#pragma unroll
            for (int d = 0; d < 4; ++d) {
                rv += coord[d] * p.rowvec_mul[d];
                inst += coord[d] * p.stats_mul[d];
            }
            any_valid = __any_sync(0xffffffffu, valid);   // false for a whole ghost / padding warp
            // warp-uniform row group / stats instance?  (true for every large layer)
            rv0 = __shfl_sync(0xffffffffu, rv, 0);
            rv_uniform = p.rowvec == nullptr || !lead || __all_sync(0xffffffffu, !valid || rv == rv0);
            inst0 = __shfl_sync(0xffffffffu, valid ? inst : -1, 0);
            inst_uniform = stats != nullptr && __all_sync(0xffffffffu, !valid || inst == inst0) && inst0 >= 0;
            // the 4 rows this lane stores: their pixel index / validity come from the lanes that own them
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                spix[i] = __shfl_sync(0xffffffffu, pix, srow0 + 8 * i);
                svalid[i] = __shfl_sync(0xffffffffu, (int)valid, srow0 + 8 * i) != 0;
            }
            }
            // stage bias (+ the shared rowvec row) for this N tile: overlaps the tile's MMAs
            __syncwarp();
            for (int c = c_begin + lane; c < c_end; c += 32) {
                float a = 0.0f;
                if (lead && any_valid && n0 + c < p.cout) {
                    if (p.bias) a = __ldg(&p.bias[n0 + c]);
                    if (p.rowvec && rv_uniform) a += __ldg(&p.rowvec[(int64_t)rv0 * p.ld_rowvec + n0 + c]);
                }
                addv[c] = a;
            }
            __syncwarp();
            const bool use_res = p.residual != nullptr && lead && !(dbg & 4);
            // The residual used to be loaded where it is consumed: an exposed HBM round trip per 16-column
            // chunk that made the residual-carrying temporal convs 2x slower than the same launch without it
            // (1.20 vs 0.71 ms at 1.8 M rows x 128 channels).  Now (a) the NEXT tile's residual window of this
            // warp is pulled into L2 one whole tile period ahead, and (b) the chunk's values travel one chunk
            // ahead in registers (`rres`), the first chunk being requested before the wait on the MMAs.
            if (p.residual != nullptr && !(dbg & 4) && it + 1 < tr.count) {
                int n_idx2, o2[4], split2, mt2;
                decode_tile(p, tile + tr.step, n_idx2, o2, split2, &mt2);
                const int cpf = n_idx2 * p.block_n + c_begin + 16 * piece;   // 4 lanes x 64 B = this warp's 64 columns
                if (p.linear) {
                    if (split2 == 0 && mt2 < p.num_m_tiles && cpf < p.cout && c_begin + 16 * piece < c_end) {
#pragma unroll
                        for (int i = 0; i < 4; ++i)
                            prefetch_l2(p.residual + ((int64_t)mt2 * kTileM + quad * 32 + srow0 + 8 * i) * p.ld_res + cpf);
                    }
                } else if (split2 == 0) {
                    int r2 = row;
                    int64_t pix2 = p.out_off;
                    bool valid2 = true;
#pragma unroll
                    for (int d = 0; d < 4; ++d) {
                        const int cd = o2[d] + (r2 & ((1 << p.tile_log2[d]) - 1));
                        r2 >>= p.tile_log2[d];
                        valid2 = valid2 && (cd < p.out_dims[d]);
                        pix2 += cd * p.out_mul[d];
                    }
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const int64_t sp2 = __shfl_sync(0xffffffffu, pix2, srow0 + 8 * i);
                        const bool sv2 = __shfl_sync(0xffffffffu, (int)valid2, srow0 + 8 * i) != 0;
                        if (sv2 && cpf < p.cout && c_begin + 16 * piece < c_end)
                            prefetch_l2(p.residual + sp2 * p.ld_res + cpf);
                    }
                }
            }
            float4 rres[4];
            auto load_res = [&](int c) {
                const int n = n0 + c;
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    rres[i] = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (svalid[i] && n < p.cout)
                        rres[i] = ld_nc_f4(p.residual + spix[i] * p.ld_res + n + 4 * piece);
                }
            };
            if (use_res && c_begin < c_end) load_res(c_begin);

            mbar_wait(&tfull_bar[acc], acc_phase, 400 + acc);
            tc_fence_after();
            const uint32_t t_row = tmem_base + ((uint32_t)(quad * 32) << 16) + acc * acc_stride;

            auto process = [&](uint32_t (&raw)[16], int c) {
                const int n = n0 + c;
                if (n >= p.cout) return;  // warp-uniform
                float v[16];
                {   // bias (+ shared rowvec row) of these 16 columns: 4 broadcast LDS.128 instead of 16 scalar reads
                    const float4* a4 = reinterpret_cast<const float4*>(addv + c);
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const float4 a = a4[q];
                        v[4 * q] = __uint_as_float(raw[4 * q]) + a.x;
                        v[4 * q + 1] = __uint_as_float(raw[4 * q + 1]) + a.y;
                        v[4 * q + 2] = __uint_as_float(raw[4 * q + 2]) + a.z;
                        v[4 * q + 3] = __uint_as_float(raw[4 * q + 3]) + a.w;
                    }
                }
                if (valid && p.rowvec && !rv_uniform && lead) {
                    const float* rp = p.rowvec + (int64_t)rv * p.ld_rowvec + n;
#pragma unroll
                    for (int j = 0; j < 16; ++j)
                        if (n + j < p.cout) v[j] += __ldg(&rp[j]);
                }
                const bool need_f32_phase = has_f32 || use_res || has_stats;
                if (need_f32_phase) {
                    // A: own row -> slab (rows / columns that do not exist are staged as 0 for the column sums)
                    float4* srow = reinterpret_cast<float4*>(slab + lane * kSlabStride);
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        float4 x = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
                        if (!valid) x = make_float4(0.f, 0.f, 0.f, 0.f);
                        srow[q] = x;
                    }
                    __syncwarp();
                    // C: coalesced residual read + output store: 4 lanes per row, 8 rows per instruction.  The same
                    // 4 x float4 give this lane's share of the GroupNorm column sums (columns 4*piece .. +3 over
                    // its 4 rows) for free: the sums used to be a second walk over the slab, 32 shared-memory loads
                    // per lane and chunk (ncu: the short-K launches spend 45 % of the LSU shared-memory pipe in the
                    // epilogue, next to the MMAs' operand reads).
                    const bool fast_stats = has_stats && inst_uniform;
                    const bool row_back = use_res && (has_hl || (has_stats && !inst_uniform));
                    float cs[4] = {0.f, 0.f, 0.f, 0.f}, cq[4] = {0.f, 0.f, 0.f, 0.f};
                    if (has_f32 || use_res || fast_stats) {
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            if (!svalid[i]) continue;
                            float4* sp = reinterpret_cast<float4*>(slab + (srow0 + 8 * i) * kSlabStride) + piece;
                            float4 x = *sp;
                            if (use_res) {
                                const float4 rr = rres[i];
                                x.x += rr.x; x.y += rr.y; x.z += rr.z; x.w += rr.w;
                                if (row_back) *sp = x;
                            }
                            if (has_f32 && !(dbg & 2)) {
                                float4* op = reinterpret_cast<float4*>(p.out_f32 + spix[i] * p.ldc + n) + piece;
                                if (splitk && p.split_stride > 0)      // own slice, plain store: the caller reduces
                                    *reinterpret_cast<float4*>(reinterpret_cast<float*>(op) + split * p.split_stride) = x;
                                else if (splitk) atomicAdd(op, x);     // partial sum of this K range
                                else *op = x;
                            }
                            cs[0] += x.x; cs[1] += x.y; cs[2] += x.z; cs[3] += x.w;
                            cq[0] = fmaf(x.x, x.x, cq[0]); cq[1] = fmaf(x.y, x.y, cq[1]);
                            cq[2] = fmaf(x.z, x.z, cq[2]); cq[3] = fmaf(x.w, x.w, cq[3]);
                        }
                        if (use_res) {
                            if (c + 16 < c_end) load_res(c + 16);   // next chunk's residual: in flight during the rest
                            if (row_back) {                         // own row again, residual included
                                __syncwarp();
#pragma unroll
                                for (int q = 0; q < 4; ++q) {
                                    const float4 x = srow[q];
                                    v[4 * q] = x.x; v[4 * q + 1] = x.y; v[4 * q + 2] = x.z; v[4 * q + 3] = x.w;
                                }
                            }
                        }
                    }
                    // D: GroupNorm partial sums of the 32 x 16 block
                    if (has_stats) {
                        if (inst_uniform) {
                            // recursive halving over the 8 lanes that share `piece` (lane bits 4, 3, 2): 7 shuffles
                            // leave every lane with ONE finished value -- (sum | sum of squares) of one column
                            const bool b4 = (lane & 16) != 0, b3 = (lane & 8) != 0, b2 = (lane & 4) != 0;
                            float k4[4];
#pragma unroll
                            for (int j = 0; j < 4; ++j) {
                                const float send = b4 ? cs[j] : cq[j];
                                const float recv = __shfl_xor_sync(0xffffffffu, send, 16);
                                k4[j] = (b4 ? cq[j] : cs[j]) + recv;
                            }
                            float k2[2];
#pragma unroll
                            for (int j = 0; j < 2; ++j) {
                                const float send = b3 ? k4[j] : k4[2 + j];
                                const float recv = __shfl_xor_sync(0xffffffffu, send, 8);
                                k2[j] = (b3 ? k4[2 + j] : k4[j]) + recv;
                            }
                            const float send = b2 ? k2[0] : k2[1];
                            const float recv = __shfl_xor_sync(0xffffffffu, send, 4);
                            const float tot = (b2 ? k2[1] : k2[0]) + recv;
                            const int col = 4 * piece + (b3 ? 2 : 0) + (b2 ? 1 : 0);
                            if (n + col < p.cout)
                                atomicAdd(&stats[((int64_t)inst0 * p.stats_ld + n + col) * 2 + (b4 ? 1 : 0)],
                                          (double)tot);
                        } else if (valid) {
#pragma unroll
                            for (int j = 0; j < 16; ++j)
                                if (n + j < p.cout) {
                                    double* sp = &stats[((int64_t)inst * p.stats_ld + n + j) * 2];
                                    atomicAdd(sp, (double)v[j]);
                                    atomicAdd(sp + 1, (double)v[j] * (double)v[j]);
                                }
                        }
                    }
                    __syncwarp();
                }
                if (has_hl && !(dbg & 2)) {
                    // bf16 planes: one slab row = [16 hi | 16 lo] = 64 bytes; pieces 0,1 -> hi plane, 2,3 -> lo
                    uint4 h0, l0, h1, l1;
                    split8(v, h0, l0);
                    split8(v + 8, h1, l1);
                    uint4* srow = reinterpret_cast<uint4*>(slab + lane * kSlabStride);
                    srow[0] = h0; srow[1] = h1; srow[2] = l0; srow[3] = l1;
                    __syncwarp();
                    __nv_bfloat16* plane = piece < 2 ? p.out_hi : p.out_lo;
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        if (!svalid[i]) continue;
                        const uint4 x = *(reinterpret_cast<const uint4*>(slab + (srow0 + 8 * i) * kSlabStride) + piece);
                        *reinterpret_cast<uint4*>(plane + spix[i] * p.ldc + n + 8 * (piece & 1)) = x;
                    }
                    __syncwarp();
                }
            };

            // TMEM loads are software pipelined: chunk c+16 is in flight while chunk c is processed
            uint32_t ra[16], rb[16];
            if (p.fuse2) {
                // two accumulator halves per chunk: [hi*hi + lo*hi | hi*lo].  single CTA: [sum | hi*lo]; cta_group::2
                // pair: [., . | ., .] per half of the columns (see the issuer).  The TMEM reads of chunk c + 16 are in
                // flight while chunk c is processed (three register arrays: current, next, shared second half) -- an
                // unpipelined load -> wait -> process chain made the epilogue latency-bound (~7 us per 128 x 128 tile).
                const int hb = p.block_n >> 1;
                auto col_a = [&](int c) { return kCta2 ? (c >= hb ? p.block_n + (c - hb) : c) : c; };
                auto col_b = [&](int c) { return kCta2 ? col_a(c) + hb : p.block_n + c; };
                uint32_t rt[16];
                if (c_begin < c_end) {
                    tmem_ld16(t_row + col_a(c_begin), ra);
                    tmem_ld16(t_row + col_b(c_begin), rt);
                }
                for (int c = c_begin; c < c_end; c += 32) {
                    tmem_ld_wait16(ra);
                    tmem_ld_wait16(rt);
#pragma unroll
                    for (int j = 0; j < 16; ++j) ra[j] = __float_as_uint(__uint_as_float(ra[j]) + __uint_as_float(rt[j]));
                    if (c + 16 < c_end) {
                        tmem_ld16(t_row + col_a(c + 16), rb);
                        tmem_ld16(t_row + col_b(c + 16), rt);
                    }
                    process(ra, c);
                    if (c + 16 < c_end) {
                        tmem_ld_wait16(rb);
                        tmem_ld_wait16(rt);
#pragma unroll
                        for (int j = 0; j < 16; ++j) rb[j] = __float_as_uint(__uint_as_float(rb[j]) + __uint_as_float(rt[j]));
                        if (c + 32 < c_end) {
                            tmem_ld16(t_row + col_a(c + 32), ra);
                            tmem_ld16(t_row + col_b(c + 32), rt);
                        }
                        process(rb, c + 16);
                    }
                }
            } else {
            if (c_begin < c_end) tmem_ld16(t_row + c_begin, ra);
            for (int c = c_begin; c < c_end; c += 32) {
                tmem_ld_wait16(ra);
                if (c + 16 < c_end) tmem_ld16(t_row + c + 16, rb);
                process(ra, c);
                if (c + 16 < c_end) {
                    tmem_ld_wait16(rb);
                    if (c + 32 < c_end) tmem_ld16(t_row + c + 32, ra);
                    process(rb, c + 16);
                }
            }
            }
            tc_fence_before();
            if (kCta2 && crank != 0) mbar_arrive_cluster(mapa_shared(smem_u32(&tempty_bar[acc]), 0));
            else mbar_arrive(&tempty_bar[acc]);
        }
    }

    tc_fence_before();
    __syncthreads();
    if constexpr (kCta2) cluster_sync_all();   // the leader's MMAs read the peer's shared memory and write its TMEM
    if (warp == 2) {
        tc_fence_after();
        if constexpr (kCta2) tmem_dealloc_2sm(tmem_base, kTmemCols);
        else tmem_dealloc(tmem_base, kTmemCols);
    }
    if (p.cluster > 1) cluster_sync_all();   // no CTA leaves while its peer can still signal its barriers
}


// ===========================================================================
// Conv3d as ONE launch: the spatial conv and the temporal conv that consumes it, interleaved tile by tile
// (guided_diffusion/nn.py:53-87: Conv2d on every frame, then a zero-padded Conv1d(k = 3) over the frame axis).
//
// The temporal conv of a narrow layer (Cout <= 128) is bound by its epilogue, not by its 6-12 k-steps of MMA
// (0.9 ms against 0.39 ms of tensor time at 1.8 M rows), and the spatial conv before it is bound by its MMAs.  Run
// back to back as two launches the two phases cannot hide each other.  Here every CTA pair walks ONE sequence of
// work items that alternates a spatial pair-tile S(i) with the temporal pair-tile T(i - lag): the long spatial
// K loops keep the tensor pipe busy while the epilogue warps drain the temporal tiles (fp32 store, GroupNorm sums,
// residual / embedding add), and the short temporal K loops ride on the same operand ring.
//
//   * both programs are cta_group::2 pair MMAs with the fused split product (block_n <= 128), same stage layout;
//   * both tile the SAME dense row space into the same 128-row blocks (the host checks this), so temporal tile m
//     reads the spatial output rows of tiles m - tpf, m, m + tpf (tpf = tiles per frame) and nothing else;
//   * the intermediate y (bf16 hi/lo planes) goes through global memory, but `lag` keeps only ~30 MB of it in
//     flight, i.e. it is produced into and consumed from the 126 MB L2;
//   * ordering: the spatial epilogue publishes tile m with bar.sync (its 8 warps) -> __threadfence ->
//     st.release.gpu flags[m]; the temporal producer acquires the three flags it depends on and issues
//     fence.proxy.async before the TMA reads.  lag >= tpf/2 + (pairs in the grid) makes every dependency an item
//     of an EARLIER iteration of some pair, and spatial items never wait, so the walk cannot deadlock; a bounded
//     spin traps instead of hanging if that reasoning is ever wrong.
// ===========================================================================
struct alignas(64) DualParams {
    IgemmParams g[2];      // 0 = spatial (writes y planes), 1 = temporal (reads them)
    unsigned int* flags;   // [tiles] spatial tile m is complete (zeroed before every launch)
    int lag;               // pair-tiles between S(i) and the temporal item issued with it
    int tpf;               // 128-row tiles per (sample, frame): dependency distance of the temporal taps
    int frames;            // F
    int tiles;             // 128-row tiles of the row space (both programs)
    int mp;                // pair-tiles = ceil(tiles / 2)
    int iters;             // iterations of every pair: ceil((mp + lag) / pairs)
    int acc4;              // four 128-column accumulators (two per program, unfused MMAs) instead of two fused 256-column
                           // ones shared by strict S / T alternation -- with those each program effectively owns ONE
                           // accumulator and MMA(S k+1) waits for the epilogue of S k
};

__device__ __forceinline__ void dual_tile_origin(const IgemmParams& p, int m, int o[4]) {
    if (m >= p.num_m_tiles) {   // ghost tile of an odd tile count: a box wholly outside the tensor
        o[0] = o[1] = o[2] = 0;
        o[3] = p.ntile[3] << p.tile_log2[3];
        return;
    }
#pragma unroll
    for (int d = 0; d < 4; ++d) {
        const int j = m % p.ntile[d];
        m /= p.ntile[d];
        o[d] = j << p.tile_log2[d];
    }
}

__device__ __forceinline__ unsigned int ld_acquire_gpu(const unsigned int* p) {
    unsigned int v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_gpu(unsigned int* p, unsigned int v) {
    asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void fence_proxy_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }

__device__ __forceinline__ unsigned int ld_relaxed_gpu(const unsigned int* p) {
    unsigned int v;
    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
// Wait until the spatial tiles a temporal tile reads (frames f - 1, f, f + 1 of the same pixels) are published.
// The three flags are read with independent relaxed loads (one L2 round trip, not three) and ordered by ONE acquire
// fence; in steady state they were set iterations ago, so this costs ~1 us of the producer thread's slack.
__device__ __forceinline__ void dual_wait_deps(const DualParams& dp, int m) {
    const int f = (m / dp.tpf) % dp.frames;
    const unsigned int* f0 = dp.flags + m;
    const unsigned int* f1 = f > 0 ? f0 - dp.tpf : f0;
    const unsigned int* f2 = f + 1 < dp.frames ? f0 + dp.tpf : f0;
    const long long t0 = clock64();
    for (;;) {
        const unsigned int a = ld_relaxed_gpu(f0), b = ld_relaxed_gpu(f1), c = ld_relaxed_gpu(f2);
        if (a & b & c) break;
        __nanosleep(64);
        if (clock64() - t0 > 4000000000LL) {
            printf("v2a: dual-conv dependency timeout tile=%d block=%d\n", m, (int)blockIdx.x);
            __trap();
        }
    }
    asm volatile("fence.acq_rel.gpu;" ::: "memory");
    fence_proxy_async_all();      // generic-proxy stores of other CTAs -> this thread's TMA reads
}

__global__ void __maxnreg__(V2A_IGEMM_MAXNREG) igemm_dual_kernel(const __grid_constant__ DualParams dp) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int S = dp.g[0].stages;
    const uint32_t stage_bytes = dp.g[0].stage_bytes;

    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)S * stage_bytes);
    uint64_t* full_bar = bars;
    uint64_t* empty_bar = bars + S;
    uint64_t* tfull_bar = bars + 2 * S;
    uint64_t* tempty_bar = bars + 2 * S + 4;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * S + 8);
    float* warp_add = reinterpret_cast<float*>(bars) + 64;
    float* stage_slab = warp_add + 8 * 256;

    pdl_trigger();
    if (warp == 0 && lane == 0) {
        for (int s = 0; s < S; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], 1);
        }
        for (int s = 0; s < 4; ++s) {
            mbar_init(&tfull_bar[s], 1);
            mbar_init(&tempty_bar[s], 2 * kEpilogueThreads);
        }
        fence_mbar_init();
    } else if (warp == 1 && lane == 0) {
        for (int w = 0; w < 2; ++w) {
            tma_prefetch_desc(&dp.g[w].a_hi[0]);
            tma_prefetch_desc(&dp.g[w].a_lo[0]);
            tma_prefetch_desc(&dp.g[w].bh_hi);
            tma_prefetch_desc(&dp.g[w].bh_lo);
        }
    } else if (warp == 2) {
        tmem_alloc_2sm(tmem_slot, kTmemCols);
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    tc_fence_after();
    pdl_wait();                                        // see igemm_kernel
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t acc_stride = kTmemCols >> 1;       // two accumulators of 256 columns
    const uint32_t crank = cluster_ctarank();
    const int pairs = gridDim.x >> 1;
    const int pr = blockIdx.x >> 1;

    if (warp == 0 && lane == 0) {
        // ===================== TMA producer =====================
        int stage = 0;
        uint32_t phase = 0;
        for (int k = 0; k < dp.iters; ++k) {
            const int i = pr + k * pairs;
            // the temporal item of this iteration: its dependencies are checked while the spatial item's loads are
            // queued behind a full operand ring (the producer thread has nothing else to do there), not in the gap
            // between the two items where every microsecond of this thread stalls the MMAs
            const int qt = i - dp.lag;
            const int mt = 2 * qt + (int)crank;
            bool t_pending = qt >= 0 && qt < dp.mp && mt < dp.tiles;
            for (int which = 0; which < 2; ++which) {
                const int q = which == 0 ? i : qt;
                if (q < 0 || q >= dp.mp) continue;
                const IgemmParams& p = dp.g[which];
                const int m = 2 * q + (int)crank;
                int o[4];
                dual_tile_origin(p, m, o);
                if (which == 1 && t_pending) {
                    dual_wait_deps(dp, mt);
                    t_pending = false;
                }
                const uint32_t half_rows = p.block_n >> 1, half_bytes = p.b_tile_bytes >> 1;
                int kit = 0;
                for (int e = 0; e < p.ntaps; ++e) {
                    const int src = p.tap_src[e];
                    const int c1 = o[0] + p.tap_d[e][0], c2 = o[1] + p.tap_d[e][1];
                    const int c3 = o[2] + p.tap_d[e][2], c4 = o[3] + p.tap_d[e][3];
                    for (int ch = 0; ch < p.tap_chunks[e]; ++ch, ++kit) {
                        mbar_wait(&empty_bar[stage], phase ^ 1, 100 + stage);
                        uint8_t* st = smem + (size_t)stage * stage_bytes;
                        const uint32_t lead_full = mapa_shared(smem_u32(&full_bar[stage]), 0);
                        if (crank == 0) mbar_arrive_expect_tx(&full_bar[stage], 2 * stage_bytes);
                        tma_load_5d_2sm(st, &p.a_hi[src], lead_full, ch * kChunkK, c1, c2, c3, c4);
                        tma_load_5d_2sm(st + kATileBytes, &p.a_lo[src], lead_full, ch * kChunkK, c1, c2, c3, c4);
                        uint8_t* sb = st + 2 * kATileBytes;
                        tma_load_2d_2sm(sb, &p.bh_hi, lead_full, kit * kChunkK, crank * half_rows);
                        tma_load_2d_2sm(sb + half_bytes, &p.bh_lo, lead_full, kit * kChunkK, crank * half_rows);
                        if (++stage == S) { stage = 0; phase ^= 1; }
                        if (which == 0 && t_pending && kit == 1) {
                            dual_wait_deps(dp, mt);
                            t_pending = false;
                        }
                    }
                }
            }
        }
    } else if (warp == 1 && lane == 0) {
        // ===================== MMA issuer (leader CTA of the pair) =====================
        if (crank == 0) {
            int stage = 0, it = 0, its[2] = {0, 0};
            uint32_t phase = 0;
            for (int k = 0; k < dp.iters; ++k) {
                const int i = pr + k * pairs;
                for (int which = 0; which < 2; ++which) {
                    const int q = which == 0 ? i : i - dp.lag;
                    if (q < 0 || q >= dp.mp) continue;
                    const IgemmParams& p = dp.g[which];
                    const int hb = p.block_n >> 1;
                    const uint32_t idesc2 = umma_idesc_16(256, 2 * p.block_n, 0, 0);
                    const uint32_t idesc1 = umma_idesc_16(256, p.block_n, 0, 0);
                    const int cnt = dp.acc4 ? its[which] : it;
                    const int acc = dp.acc4 ? 2 * which + (cnt & 1) : (cnt & 1);
                    const uint32_t acc_phase = (cnt >> 1) & 1;
                    ++it;
                    ++its[which];
                    mbar_wait(&tempty_bar[acc], acc_phase ^ 1, 200 + acc);
                    tc_fence_after();
                    const uint32_t d_tmem = tmem_base + acc * (dp.acc4 ? 128u : acc_stride);
                    for (int kit = 0; kit < p.k_iters; ++kit) {
                        mbar_wait(&full_bar[stage], phase, 300 + stage);
                        tc_fence_after();
                        const uint32_t sa = smem_u32(smem + (size_t)stage * stage_bytes);
                        const uint64_t a_hi = umma_desc_sw128(sa);
                        const uint64_t a_lo = umma_desc_sw128(sa + kATileBytes);
                        const uint64_t b_hi = umma_desc_sw128(sa + 2 * kATileBytes);
                        if (dp.acc4) {
                            // three M = 256 MMAs into ONE 128-column accumulator, columns in natural order (this CTA's
                            // half of the weight rows | the peer's)
                            const uint64_t b_lo = umma_desc_sw128(sa + 2 * kATileBytes + (p.b_tile_bytes >> 1));
#pragma unroll
                            for (int kk = 0; kk < 4; ++kk)
                                umma_bf16_2sm(d_tmem, a_lo + 2 * kk, b_hi + 2 * kk, idesc1, (kit | kk) != 0);
#pragma unroll
                            for (int kk = 0; kk < 4; ++kk) umma_bf16_2sm(d_tmem, a_hi + 2 * kk, b_lo + 2 * kk, idesc1, 1);
#pragma unroll
                            for (int kk = 0; kk < 4; ++kk) umma_bf16_2sm(d_tmem, a_hi + 2 * kk, b_hi + 2 * kk, idesc1, 1);
                        } else {
#pragma unroll
                        for (int kk = 0; kk < 4; ++kk)
                            umma_bf16_2sm(d_tmem, a_hi + 2 * kk, b_hi + 2 * kk, idesc2, (kit | kk) != 0);
#pragma unroll
                        for (int kk = 0; kk < 4; ++kk)
                            umma_bf16_2sm(d_tmem + hb, a_lo + 2 * kk, b_hi + 2 * kk, idesc1, 1);
                        }
                        umma_commit_2sm_mc(&empty_bar[stage], 3);
                        if (kit == p.k_iters - 1) umma_commit_2sm_mc(&tfull_bar[acc], 3);
                        if (++stage == S) { stage = 0; phase ^= 1; }
                    }
                }
            }
        }
    } else if (warp >= 4) {
        // ===================== epilogue (both CTAs; the code of igemm_kernel's epilogue, per work item) ==========
        const int quad = warp & 3;
        const int row = quad * 32 + lane;
        const int half = (warp - 4) >> 2;
        float* addv = warp_add + (warp - 4) * 256;
        float* slab = stage_slab + (warp - 4) * (32 * kSlabStride);
        const int srow0 = lane >> 2;
        const int piece = lane & 3;
        int it = 0, its[2] = {0, 0};
        for (int k = 0; k < dp.iters; ++k) {
            const int i = pr + k * pairs;
#pragma unroll
            for (int which = 0; which < 2; ++which) {     // unrolled: each program's epilogue is specialised below
                const int q = which == 0 ? i : i - dp.lag;
                if (q < 0 || q >= dp.mp) continue;
                const IgemmParams& p = dp.g[which];
                // the spatial program writes hi/lo planes and nothing else, the temporal one fp32 (+ sums): fixed by
                // the plan, so the feature tests below fold at compile time (see igemm_kernel's kEpi)
                const bool has_f32 = which == 1, has_hl = which == 0;
                const bool has_stats = which == 1 && p.stats != nullptr;
                const int m = 2 * q + (int)crank;
                const int cnt = dp.acc4 ? its[which] : it;
                const int acc = dp.acc4 ? 2 * which + (cnt & 1) : (cnt & 1);
                const uint32_t acc_phase = (cnt >> 1) & 1;
                ++it;
                ++its[which];
                const int nch = p.block_n >> 4;
                const int c_begin = half == 0 ? 0 : ((nch + 1) >> 1) << 4;
                const int c_end = half == 0 ? ((nch + 1) >> 1) << 4 : p.block_n;
                double* const stats = has_stats ? p.stats + (long long)(blockIdx.x % p.stats_replicas) * p.stats_rep_stride
                                                : nullptr;
                int o[4];
                dual_tile_origin(p, m, o);
                int r = row, coord[4];
                bool valid = true;
#pragma unroll
                for (int d = 0; d < 4; ++d) {
                    coord[d] = o[d] + (r & ((1 << p.tile_log2[d]) - 1));
                    r >>= p.tile_log2[d];
                    valid = valid && (coord[d] < p.out_dims[d]);
                }
                const int64_t pix = p.out_off + coord[0] * p.out_mul[0] + coord[1] * p.out_mul[1] +
                                    coord[2] * p.out_mul[2] + coord[3] * p.out_mul[3];
                int rv = 0, inst = 0;
#pragma unroll
                for (int d = 0; d < 4; ++d) {
                    rv += coord[d] * p.rowvec_mul[d];
                    inst += coord[d] * p.stats_mul[d];
                }
                const bool any_valid = __any_sync(0xffffffffu, valid);
                const int rv0 = __shfl_sync(0xffffffffu, rv, 0);
                const bool rv_uniform = p.rowvec == nullptr || __all_sync(0xffffffffu, !valid || rv == rv0);
                const int inst0 = __shfl_sync(0xffffffffu, valid ? inst : -1, 0);
                const bool inst_uniform =
                    stats != nullptr && __all_sync(0xffffffffu, !valid || inst == inst0) && inst0 >= 0;
                int64_t spix[4];
                bool svalid[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    spix[j] = __shfl_sync(0xffffffffu, pix, srow0 + 8 * j);
                    svalid[j] = __shfl_sync(0xffffffffu, (int)valid, srow0 + 8 * j) != 0;
                }
                __syncwarp();
                for (int c = c_begin + lane; c < c_end; c += 32) {
                    float a = 0.0f;
                    if (any_valid && c < p.cout) {
                        if (p.bias) a = __ldg(&p.bias[c]);
                        if (p.rowvec && rv_uniform) a += __ldg(&p.rowvec[(int64_t)rv0 * p.ld_rowvec + c]);
                    }
                    addv[c] = a;
                }
                __syncwarp();
                const bool use_res = which == 1 && p.residual != nullptr;
                if (use_res) {      // pull the residual window of this warp's NEXT temporal tile into L2
                    const int m2 = m + 2 * pairs;
                    if (m2 < p.num_m_tiles) {
                        int o2[4];
                        dual_tile_origin(p, m2, o2);
                        int r2 = row;
                        int64_t pix2 = p.out_off;
                        bool valid2 = true;
#pragma unroll
                        for (int d = 0; d < 4; ++d) {
                            const int cd = o2[d] + (r2 & ((1 << p.tile_log2[d]) - 1));
                            r2 >>= p.tile_log2[d];
                            valid2 = valid2 && (cd < p.out_dims[d]);
                            pix2 += cd * p.out_mul[d];
                        }
                        const int cpf = c_begin + 16 * piece;
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            const int64_t sp2 = __shfl_sync(0xffffffffu, pix2, srow0 + 8 * j);
                            const bool sv2 = __shfl_sync(0xffffffffu, (int)valid2, srow0 + 8 * j) != 0;
                            if (sv2 && cpf < p.cout && cpf < c_end) prefetch_l2(p.residual + sp2 * p.ld_res + cpf);
                        }
                    }
                }
                float4 rres[4];
                auto load_res = [&](int c) {
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        rres[j] = make_float4(0.f, 0.f, 0.f, 0.f);
                        if (svalid[j] && c < p.cout) rres[j] = ld_nc_f4(p.residual + spix[j] * p.ld_res + c + 4 * piece);
                    }
                };
                if (use_res && c_begin < c_end) load_res(c_begin);

                mbar_wait(&tfull_bar[acc], acc_phase, 400 + acc);
                tc_fence_after();
                const uint32_t t_row = tmem_base + ((uint32_t)(quad * 32) << 16) + acc * (dp.acc4 ? 128u : acc_stride);

                auto process = [&](uint32_t (&raw)[16], int c) {
                    const int n = c;
                    if (n >= p.cout) return;
                    float v[16];
                    {
                        const float4* a4 = reinterpret_cast<const float4*>(addv + c);
#pragma unroll
                        for (int qq = 0; qq < 4; ++qq) {
                            const float4 a = a4[qq];
                            v[4 * qq] = __uint_as_float(raw[4 * qq]) + a.x;
                            v[4 * qq + 1] = __uint_as_float(raw[4 * qq + 1]) + a.y;
                            v[4 * qq + 2] = __uint_as_float(raw[4 * qq + 2]) + a.z;
                            v[4 * qq + 3] = __uint_as_float(raw[4 * qq + 3]) + a.w;
                        }
                    }
                    if (which == 1 && valid && p.rowvec && !rv_uniform) {
                        const float* rp = p.rowvec + (int64_t)rv * p.ld_rowvec + n;
#pragma unroll
                        for (int j = 0; j < 16; ++j)
                            if (n + j < p.cout) v[j] += __ldg(&rp[j]);
                    }
                    const bool need_f32_phase = has_f32 || use_res || has_stats;
                    if (need_f32_phase) {
                        float4* srow = reinterpret_cast<float4*>(slab + lane * kSlabStride);
#pragma unroll
                        for (int qq = 0; qq < 4; ++qq) {
                            float4 x = make_float4(v[4 * qq], v[4 * qq + 1], v[4 * qq + 2], v[4 * qq + 3]);
                            if (!valid) x = make_float4(0.f, 0.f, 0.f, 0.f);
                            srow[qq] = x;
                        }
                        __syncwarp();
                        const bool fast_stats = has_stats && inst_uniform;
                        const bool row_back = use_res && (has_hl || (has_stats && !inst_uniform));
                        float cs[4] = {0.f, 0.f, 0.f, 0.f}, cq[4] = {0.f, 0.f, 0.f, 0.f};
                        if (has_f32 || use_res || fast_stats) {
#pragma unroll
                            for (int j = 0; j < 4; ++j) {
                                if (!svalid[j]) continue;
                                float4* sp = reinterpret_cast<float4*>(slab + (srow0 + 8 * j) * kSlabStride) + piece;
                                float4 x = *sp;
                                if (use_res) {
                                    const float4 rr = rres[j];
                                    x.x += rr.x; x.y += rr.y; x.z += rr.z; x.w += rr.w;
                                    if (row_back) *sp = x;
                                }
                                if (has_f32)
                                    *(reinterpret_cast<float4*>(p.out_f32 + spix[j] * p.ldc + n) + piece) = x;
                                cs[0] += x.x; cs[1] += x.y; cs[2] += x.z; cs[3] += x.w;
                                cq[0] = fmaf(x.x, x.x, cq[0]); cq[1] = fmaf(x.y, x.y, cq[1]);
                                cq[2] = fmaf(x.z, x.z, cq[2]); cq[3] = fmaf(x.w, x.w, cq[3]);
                            }
                            if (use_res) {
                                if (c + 16 < c_end) load_res(c + 16);
                                if (row_back) {
                                    __syncwarp();
#pragma unroll
                                    for (int qq = 0; qq < 4; ++qq) {
                                        const float4 x = srow[qq];
                                        v[4 * qq] = x.x; v[4 * qq + 1] = x.y; v[4 * qq + 2] = x.z; v[4 * qq + 3] = x.w;
                                    }
                                }
                            }
                        }
                        if (has_stats) {
                            if (inst_uniform) {
                                const bool b4 = (lane & 16) != 0, b3 = (lane & 8) != 0, b2 = (lane & 4) != 0;
                                float k4[4];
#pragma unroll
                                for (int j = 0; j < 4; ++j) {
                                    const float send = b4 ? cs[j] : cq[j];
                                    const float recv = __shfl_xor_sync(0xffffffffu, send, 16);
                                    k4[j] = (b4 ? cq[j] : cs[j]) + recv;
                                }
                                float k2[2];
#pragma unroll
                                for (int j = 0; j < 2; ++j) {
                                    const float send = b3 ? k4[j] : k4[2 + j];
                                    const float recv = __shfl_xor_sync(0xffffffffu, send, 8);
                                    k2[j] = (b3 ? k4[2 + j] : k4[j]) + recv;
                                }
                                const float send = b2 ? k2[0] : k2[1];
                                const float recv = __shfl_xor_sync(0xffffffffu, send, 4);
                                const float tot = (b2 ? k2[1] : k2[0]) + recv;
                                const int col = 4 * piece + (b3 ? 2 : 0) + (b2 ? 1 : 0);
                                if (n + col < p.cout)
                                    atomicAdd(&stats[((int64_t)inst0 * p.stats_ld + n + col) * 2 + (b4 ? 1 : 0)],
                                              (double)tot);
                            } else if (valid) {
#pragma unroll
                                for (int j = 0; j < 16; ++j)
                                    if (n + j < p.cout) {
                                        double* sp = &stats[((int64_t)inst * p.stats_ld + n + j) * 2];
                                        atomicAdd(sp, (double)v[j]);
                                        atomicAdd(sp + 1, (double)v[j] * (double)v[j]);
                                    }
                            }
                        }
                        __syncwarp();
                    }
                    if (has_hl) {
                        uint4 h0, l0, h1, l1;
                        split8(v, h0, l0);
                        split8(v + 8, h1, l1);
                        uint4* srow = reinterpret_cast<uint4*>(slab + lane * kSlabStride);
                        srow[0] = h0; srow[1] = h1; srow[2] = l0; srow[3] = l1;
                        __syncwarp();
                        __nv_bfloat16* plane = piece < 2 ? p.out_hi : p.out_lo;
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            if (!svalid[j]) continue;
                            const uint4 x = *(reinterpret_cast<const uint4*>(slab + (srow0 + 8 * j) * kSlabStride) + piece);
                            *reinterpret_cast<uint4*>(plane + spix[j] * p.ldc + n + 8 * (piece & 1)) = x;
                        }
                        __syncwarp();
                    }
                };

                // fused split product of a cta_group::2 pair: columns [hi*hi + lo*hi | hi*lo] per half of the N rows; the
                // TMEM reads of the next chunk are in flight while this one is processed (see igemm_kernel)
                uint32_t ra[16], rb[16], rt[16];
                const int hb = p.block_n >> 1;
                auto col_a = [&](int c) { return c >= hb ? p.block_n + (c - hb) : c; };
                if (dp.acc4) {      // plain 128-column accumulator: chunk c + 16 in flight while chunk c is processed
                    if (c_begin < c_end) tmem_ld16(t_row + c_begin, ra);
                    for (int c = c_begin; c < c_end; c += 32) {
                        tmem_ld_wait16(ra);
                        if (c + 16 < c_end) tmem_ld16(t_row + c + 16, rb);
                        process(ra, c);
                        if (c + 16 < c_end) {
                            tmem_ld_wait16(rb);
                            if (c + 32 < c_end) tmem_ld16(t_row + c + 32, ra);
                            process(rb, c + 16);
                        }
                    }
                } else {
                if (c_begin < c_end) {
                    tmem_ld16(t_row + col_a(c_begin), ra);
                    tmem_ld16(t_row + col_a(c_begin) + hb, rt);
                }
                for (int c = c_begin; c < c_end; c += 32) {
                    tmem_ld_wait16(ra);
                    tmem_ld_wait16(rt);
#pragma unroll
                    for (int j = 0; j < 16; ++j) ra[j] = __float_as_uint(__uint_as_float(ra[j]) + __uint_as_float(rt[j]));
                    if (c + 16 < c_end) {
                        tmem_ld16(t_row + col_a(c + 16), rb);
                        tmem_ld16(t_row + col_a(c + 16) + hb, rt);
                    }
                    process(ra, c);
                    if (c + 16 < c_end) {
                        tmem_ld_wait16(rb);
                        tmem_ld_wait16(rt);
#pragma unroll
                        for (int j = 0; j < 16; ++j) rb[j] = __float_as_uint(__uint_as_float(rb[j]) + __uint_as_float(rt[j]));
                        if (c + 32 < c_end) {
                            tmem_ld16(t_row + col_a(c + 32), ra);
                            tmem_ld16(t_row + col_a(c + 32) + hb, rt);
                        }
                        process(rb, c + 16);
                    }
                }
                }
                tc_fence_before();
                if (crank != 0) mbar_arrive_cluster(mapa_shared(smem_u32(&tempty_bar[acc]), 0));
                else mbar_arrive(&tempty_bar[acc]);
                if (which == 0) {
                    // publish the spatial tile: every epilogue warp's stores are done (bar.sync over the 256 epilogue
                    // threads), made visible GPU-wide, then the flag is released
                    asm volatile("bar.sync 1, 256;" ::: "memory");
                    if (warp == 4 && lane == 0 && m < dp.tiles) {
                        __threadfence();
                        fence_proxy_async_all();
                        st_release_gpu(dp.flags + m, 1u);
                    }
                }
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc_2sm(tmem_base, kTmemCols);
    }
    cluster_sync_all();
}


// ===========================================================================
// Small-M backend: the same tap program on CUDA cores for GEMMs of at most 32 output rows.
//
// `predict_action` runs the policy UNet at batch 1 between simulator steps (diffusion_unet_image_policy.py:88-201):
// M = B*T = 4 .. 16 rows against 65 M weights.  A 128-row tensor-core tile is then > 87 % padding and every launch
// costs ~10-40 us of fixed pipeline set-up (TMEM allocation, barrier rings, split-K memset + atomics) for a few
// microseconds of weight streaming.  Here one warp owns one output channel: it streams that channel's K-major weight
// row (hi + lo planes, 16-byte loads) exactly once and multiplies it with the im2col'd activation rows staged in shared
// memory (fp32, exact hi + lo), fp32 FMA accumulation -- the activation operand is tiny, the weights are the traffic.
// Same descriptor, same epilogue subset (bias, residual, fp32 and/or hi/lo output); chosen by the plan, not the caller.
// ===========================================================================
constexpr int kSmallMaxRows = 16;
constexpr int kSmallWarps = 8;                // warps per block = (channel groups per block) x (K slices per group)

struct SmallMParams {
    const __nv_bfloat16* a_hi[V2A_MAX_SRC];
    const __nv_bfloat16* a_lo[V2A_MAX_SRC];
    int src_ch[V2A_MAX_SRC];
    int tap_src[V2A_MAX_TAPS];
    int tap_chunk0[V2A_MAX_TAPS + 1];         // first 64-wide K chunk of every tap (prefix sums)
    int ntaps;
    const __nv_bfloat16* w_hi;
    const __nv_bfloat16* w_lo;
    int ktot, wrows;
    int rows, cout, ldc;
    float* out_f32;
    __nv_bfloat16* out_hi;
    __nv_bfloat16* out_lo;
    const float* bias;
    const float* residual;
    int ld_res;
    int slices;                               // K slices per channel group (1, 2, 4 or 8 warps share a group)
    int slice_groups;                         // 8-element groups per slice
    // the im2col shift of every (output row, tap) as an element offset into the tap's source (-1 = zero padding) and
    // the output row of every GEMM row: computed once by the host at plan time (they depend on the shape only)
    int off[kSmallMaxRows][V2A_MAX_TAPS];
    int out_row[kSmallMaxRows];
};

__device__ __forceinline__ void bf16x8_to_f32(const uint4& h, const uint4& l, float* f) {
    const uint32_t hv[4] = {h.x, h.y, h.z, h.w}, lv[4] = {l.x, l.y, l.z, l.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        f[2 * i] = __uint_as_float(hv[i] << 16) + __uint_as_float(lv[i] << 16);
        f[2 * i + 1] = __uint_as_float(hv[i] & 0xffff0000u) + __uint_as_float(lv[i] & 0xffff0000u);
    }
}
__device__ __forceinline__ uint4 ld_stream_u4(const uint4* p) {   // read-once weights: do not displace x in L1
    uint4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];"
                 : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
                 : "l"(p));
    return v;
}

// One warp = (kCh consecutive output channels, one K slice): it streams that slice of the channels' K-major weight
// rows once (2 * kCh independent 16-byte loads in flight per lane, the next group's requested before the current one
// is used) and reads the matching activation elements of every output row straight from global memory (a few KB, L1
// resident: every warp of the SM reads the same rows) ONCE for its kCh channels.  The first weight loads are issued
// before the block's only barrier (the offset table's way into shared memory), so HBM latency overlaps the set-up.
template <int kRows, int kCh>
__global__ void __launch_bounds__(kSmallWarps * 32) igemm_smallm_kernel(const __grid_constant__ SmallMParams p) {
    __shared__ int s_off[kSmallMaxRows][V2A_MAX_TAPS];
    __shared__ float s_part[kSmallWarps][kRows][kCh];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int gpb = kSmallWarps / p.slices;                       // channel groups per block
    const int n0 = (blockIdx.x * gpb + warp / p.slices) * kCh;
    const int slice = warp % p.slices;
    const int g_begin = slice * p.slice_groups;
    const int g_end = min(g_begin + p.slice_groups, p.ktot >> 3);
    const bool live = n0 < p.cout;
    const uint4* wh[kCh];
    const uint4* wl[kCh];
#pragma unroll
    for (int j = 0; j < kCh; ++j) {
        const int n = min(n0 + j, p.wrows - 1);                   // rows past cout: re-read a valid row, never stored
        wh[j] = reinterpret_cast<const uint4*>(p.w_hi + (size_t)n * p.ktot);
        wl[j] = reinterpret_cast<const uint4*>(p.w_lo + (size_t)n * p.ktot);
    }
    pdl_trigger();       // the next kernel of the chain may start streaming ITS weights while this one computes
    uint4 h[kCh], l[kCh];
    int g = g_begin + lane;
    if (live && g < g_end) {
#pragma unroll
        for (int j = 0; j < kCh; ++j) {
            h[j] = ld_stream_u4(wh[j] + g);
            l[j] = ld_stream_u4(wl[j] + g);
        }
    }
    for (int idx = threadIdx.x; idx < kSmallMaxRows * V2A_MAX_TAPS; idx += blockDim.x)
        (&s_off[0][0])[idx] = (&p.off[0][0])[idx];
    __syncthreads();
    pdl_wait();          // activations / residual come from the previous kernels; the weights above do not
    float acc[kRows][kCh];
#pragma unroll
    for (int m = 0; m < kRows; ++m)
#pragma unroll
        for (int j = 0; j < kCh; ++j) acc[m][j] = 0.0f;
    if (live) {
        for (; g < g_end; g += 32) {
            float w[kCh][8];
#pragma unroll
            for (int j = 0; j < kCh; ++j) bf16x8_to_f32(h[j], l[j], w[j]);
            if (g + 32 < g_end) {
#pragma unroll
                for (int j = 0; j < kCh; ++j) {
                    h[j] = ld_stream_u4(wh[j] + g + 32);
                    l[j] = ld_stream_u4(wl[j] + g + 32);
                }
            }
            const int k = g << 3, chunk = k >> 6;
            int e = 0;
            while (e + 1 < p.ntaps && chunk >= p.tap_chunk0[e + 1]) ++e;
            const int c = ((chunk - p.tap_chunk0[e]) << 6) + (k & 63);
            const int src = p.tap_src[e];
            if (c >= p.src_ch[src]) continue;                // zero padding of the last chunk of a tap
            const __nv_bfloat16* ah = p.a_hi[src] + c;
            const __nv_bfloat16* al = p.a_lo[src] + c;
            // rows in chunks of four, branch-free: the eight activation loads of a chunk are issued back to back (a
            // `continue` per padded row kept them from being hoisted and made every row a dependent L1 round trip)
#pragma unroll
            for (int m0 = 0; m0 < kRows; m0 += 4) {
                uint4 xh[4], xl[4];
#pragma unroll
                for (int r = 0; r < 4; ++r) {
                    const int off = m0 + r < p.rows ? s_off[m0 + r][e] : -1;
                    const int o = off < 0 ? 0 : off;
                    xh[r] = __ldg(reinterpret_cast<const uint4*>(ah + o));
                    xl[r] = __ldg(reinterpret_cast<const uint4*>(al + o));
                    if (off < 0) xh[r] = xl[r] = make_uint4(0u, 0u, 0u, 0u);      // a select, not a branch
                }
#pragma unroll
                for (int r = 0; r < 4; ++r) {
                    float x[8];
                    bf16x8_to_f32(xh[r], xl[r], x);
#pragma unroll
                    for (int j = 0; j < kCh; ++j) {
                        float a = acc[m0 + r][j];
#pragma unroll
                        for (int i = 0; i < 8; ++i) a = fmaf(x[i], w[j][i], a);
                        acc[m0 + r][j] = a;
                    }
                }
            }
        }
    }
    // ---- reduce over the lanes (recursive halving: lanes trade halves of their value list, 2N instead of 5N
    // shuffles), then over the K slices of the channel group ----
    {
        constexpr int N = kRows * kCh;
        float* v = &acc[0][0];
        int n = N, o = 16, k = 0;
#pragma unroll
        for (; o >= 1 && n > 1; o >>= 1, ++k) {
            n >>= 1;
            const bool hi = (lane & o) != 0;
#pragma unroll
            for (int i = 0; i < N / 2; ++i) {
                if (i < n) {
                    const float send = hi ? v[i] : v[i + n];
                    const float recv = __shfl_xor_sync(0xffffffffu, send, o);
                    v[i] = (hi ? v[i + n] : v[i]) + recv;
                }
            }
        }
#pragma unroll
        for (; o >= 1; o >>= 1) v[0] += __shfl_xor_sync(0xffffffffu, v[0], o);
        // lane L now holds the finished sums of value indices (L >> (5 - k)) * n .. + n - 1   (n = N >> k)
        if ((lane & ((32 >> k) - 1)) == 0) {
            const int base = (lane >> (5 - k)) * n;
#pragma unroll
            for (int i = 0; i < (N >= 32 ? N / 32 : 1); ++i) (&s_part[warp][0][0])[base + i] = v[i];
        }
    }
    __syncthreads();
    if (slice != 0 || !live) return;
    // lanes walk the (row, channel) outputs of the group
    for (int idx = lane; idx < p.rows * kCh; idx += 32) {
        const int m = idx / kCh, j = idx - m * kCh;
        const int n = n0 + j;
        if (n >= p.cout) continue;
        float v = 0.0f;
        for (int s2 = 0; s2 < p.slices; ++s2) v += s_part[warp + s2][m][j];
        const long long row = p.out_row[m];
        v += p.bias ? __ldg(p.bias + n) : 0.0f;
        if (p.residual) v += p.residual[row * p.ld_res + n];
        if (p.out_f32) p.out_f32[row * p.ldc + n] = v;
        if (p.out_hi) {
            __nv_bfloat16 hh, ll;
            split_bf16(v, hh, ll);
            p.out_hi[row * p.ldc + n] = hh;
            p.out_lo[row * p.ldc + n] = ll;
        }
    }
}

// ---------------------------------------------------------------------------
// host side: tensor maps + plan
// ---------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                  const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
    static EncodeTiledFn fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) ==
                cudaSuccess && qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    });
    return fn;
}

// bf16 tensor, dims[0] innermost (contiguous); box[0] must be 64 (128 B rows)
static int make_map(CUtensorMap* m, const void* base, int rank, const uint64_t* dims,
                    const uint32_t* box) {
    EncodeTiledFn fn = get_encode_fn();
    V2A_REQUIRE(fn != nullptr, "cuTensorMapEncodeTiled entry point unavailable (no CUDA driver?)");
    cuuint64_t gdim[5];
    cuuint64_t gstride[4];
    cuuint32_t bdim[5], estr[5];
    uint64_t stride = 2;  // bytes
    for (int i = 0; i < rank; ++i) {
        gdim[i] = dims[i];
        bdim[i] = box[i];
        estr[i] = 1;
        stride *= dims[i];
        if (i < rank - 1) gstride[i] = stride;
    }
    CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, const_cast<void*>(base),
                    gdim, gstride, bdim, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    V2A_REQUIRE(r == CUDA_SUCCESS,
                "cuTensorMapEncodeTiled failed (%d): rank %d dims [%llu %llu %llu %llu %llu] box "
                "[%u %u %u %u %u] base %p",
                (int)r, rank, (unsigned long long)dims[0], (unsigned long long)(rank > 1 ? dims[1] : 0),
                (unsigned long long)(rank > 2 ? dims[2] : 0), (unsigned long long)(rank > 3 ? dims[3] : 0),
                (unsigned long long)(rank > 4 ? dims[4] : 0), box[0], rank > 1 ? box[1] : 0,
                rank > 2 ? box[2] : 0, rank > 3 ? box[3] : 0, rank > 4 ? box[4] : 0, base);
    return 0;
}

// shared with wgrad.cu
int make_tensor_map_bf16(CUtensorMap* m, const void* base, int rank, const uint64_t* dims, const uint32_t* box) {
    return make_map(m, base, rank, dims, box);
}

struct IgemmPlan {
    IgemmParams p;
    int epi;              // compile-time epilogue class of the launch (bit 0 fp32, bit 1 hi/lo, bit 2 sums) or -1
    bool smallm;          // <= 16 output rows: the CUDA-core weight-streaming backend (igemm_smallm_kernel)
    int small_ch;         // output channels per warp there (1 or 4)
    SmallMParams sp;
    int grid;
    size_t smem;
    bool zero_out;        // split-K: clear the output window before the launch
    size_t zero_width;    // bytes per row to clear
    int64_t zero_rows;
};

static int g_num_sms = 0;
static int g_max_smem = 0;

typedef void (*IgemmKernelFn)(const IgemmParams);
static IgemmKernelFn igemm_fn(bool cta2, int epi) {
    switch (epi) {
        case 1: return cta2 ? igemm_kernel<true, 1> : igemm_kernel<false, 1>;   // fp32
        case 2: return cta2 ? igemm_kernel<true, 2> : igemm_kernel<false, 2>;   // hi/lo planes
        case 3: return cta2 ? igemm_kernel<true, 3> : igemm_kernel<false, 3>;   // fp32 + hi/lo planes
        case 5: return cta2 ? igemm_kernel<true, 5> : igemm_kernel<false, 5>;   // fp32 + GroupNorm sums
        default: return cta2 ? igemm_kernel<true, -1> : igemm_kernel<false, -1>;
    }
}

static int device_props() {
    if (g_num_sms) return 0;
    int dev = 0;
    V2A_CUDA_OK(cudaGetDevice(&dev));
    V2A_CUDA_OK(cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev));
    V2A_CUDA_OK(cudaDeviceGetAttribute(&g_max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
    return 0;
}

// K steps (of 64) from which a pair launch uses cta_group::2 MMAs (tuning probe: V2A_CTA2_MINK)
static int cta2_min_k() {
    static const int v = getenv("V2A_CTA2_MINK") ? atoi(getenv("V2A_CTA2_MINK")) : 12;   // same-box A/B: 16: 97.2 ms, 12: 96.1, 10: 97.0
    return v;
}

static int plan_create(const v2a_igemm_desc* d, IgemmPlan** out, bool force_cta2 = false) {
    if (int rc = device_props()) return rc;
    V2A_REQUIRE(d->nsrc >= 1 && d->nsrc <= V2A_MAX_SRC, "igemm: nsrc %d out of range", d->nsrc);
    V2A_REQUIRE(d->ntaps >= 1 && d->ntaps <= V2A_MAX_TAPS, "igemm: ntaps %d out of range", d->ntaps);
    V2A_REQUIRE(d->passes == 1 || d->passes == 3, "igemm: passes must be 1 or 3");
    V2A_REQUIRE(d->block_n >= 16 && d->block_n <= 256 && d->block_n % 16 == 0,
                "igemm: block_n %d must be a multiple of 16 in [16,256]", d->block_n);
    // 16-column chunks are stored whole: a partial last chunk needs the row padded to 16 columns
    V2A_REQUIRE(d->ldc % (d->out_hi ? 8 : 4) == 0 && d->ldc >= ((d->cout + 15) / 16) * 16,
                "igemm: ldc %d must be 16-byte aligned and >= cout %d rounded up to 16", d->ldc, d->cout);
    V2A_REQUIRE(d->out_f32 || d->out_hi, "igemm: no output tensor");
    V2A_REQUIRE(!d->out_hi || d->out_lo, "igemm: out_hi without out_lo");
    V2A_REQUIRE(!d->residual || d->ld_res % 4 == 0, "igemm: ld_res must be a multiple of 4");
    int tl = 0;
    for (int i = 0; i < 4; ++i) {
        V2A_REQUIRE(d->tile_log2[i] >= 0 && d->tile_log2[i] <= 7, "igemm: bad tile_log2");
        V2A_REQUIRE(d->out_dims[i] >= 1, "igemm: bad out_dims");
        tl += d->tile_log2[i];
    }
    V2A_REQUIRE(tl == 7, "igemm: tile box must hold 128 rows (sum tile_log2 = %d)", tl);

    IgemmPlan* pl = new IgemmPlan();
    memset(&pl->p, 0, sizeof(pl->p));
    IgemmParams& p = pl->p;
    int k_iters = 0;
    for (int e = 0; e < d->ntaps; ++e) {
        const v2a_igemm_tap& t = d->taps[e];
        if (!(t.src >= 0 && t.src < d->nsrc && t.nchunks >= 1)) {
            delete pl;
            V2A_REQUIRE(false, "igemm: bad tap %d", e);
        }
        p.tap_src[e] = t.src;
        for (int i = 0; i < 4; ++i) p.tap_d[e][i] = t.d[i];
        p.tap_chunks[e] = t.nchunks;
        k_iters += t.nchunks;
    }
    if (k_iters * kChunkK != d->ktot) {
        delete pl;
        V2A_REQUIRE(false, "igemm: ktot %d != 64 * sum(nchunks) %d", d->ktot, k_iters * kChunkK);
    }
    p.ntaps = d->ntaps;
    p.k_iters = k_iters;
    pl->smallm = false;
    {
        const int64_t rows = (int64_t)d->out_dims[0] * d->out_dims[1] * d->out_dims[2] * d->out_dims[3];
        const char* env = getenv("V2A_SMALLM");
        bool ok = rows <= 16 && d->passes == 3 && !d->stats && !d->rowvec && !d->a_fp16 && !d->b_fp16 &&
                  d->w_hi && d->w_lo && !force_cta2 && !(env && atoi(env) == 0);
        for (int s = 0; s < d->nsrc && ok; ++s) ok = d->src[s].channels % 8 == 0 && d->src[s].hi && d->src[s].lo;
        if (ok) {
            SmallMParams& sp = pl->sp;
            memset(&sp, 0, sizeof(sp));
            for (int s = 0; s < d->nsrc; ++s) {
                sp.a_hi[s] = reinterpret_cast<const __nv_bfloat16*>(d->src[s].hi);
                sp.a_lo[s] = reinterpret_cast<const __nv_bfloat16*>(d->src[s].lo);
                sp.src_ch[s] = d->src[s].channels;
            }
            int c0 = 0;
            for (int e = 0; e < d->ntaps; ++e) {
                sp.tap_src[e] = d->taps[e].src;
                sp.tap_chunk0[e] = c0;
                c0 += d->taps[e].nchunks;
            }
            sp.tap_chunk0[d->ntaps] = c0;
            sp.ntaps = d->ntaps;
            sp.w_hi = reinterpret_cast<const __nv_bfloat16*>(d->w_hi);
            sp.w_lo = reinterpret_cast<const __nv_bfloat16*>(d->w_lo);
            sp.ktot = d->ktot;
            sp.wrows = d->wrows;
            long long out_mul[4], out_off = 0, mul = 1;
            for (int i = 0; i < 4; ++i) {
                out_mul[i] = mul;
                mul *= d->out_dims[i];
            }
            if (d->out_pix_mul[0] | d->out_pix_mul[1] | d->out_pix_mul[2] | d->out_pix_mul[3]) {
                for (int i = 0; i < 4; ++i) out_mul[i] = d->out_pix_mul[i];
                out_off = d->out_pix_off;
            }
            // (row, tap) -> element offset into the tap's source, or -1 where the tap falls into the zero padding
            for (int m = 0; m < (int)rows; ++m) {
                int c[4], mm = m;
                for (int i = 0; i < 4; ++i) {
                    c[i] = mm % d->out_dims[i];
                    mm /= d->out_dims[i];
                }
                const long long orow = out_off + c[0] * out_mul[0] + c[1] * out_mul[1] + c[2] * out_mul[2] + c[3] * out_mul[3];
                V2A_REQUIRE(orow >= 0 && orow < ((long long)1 << 31), "igemm (small M): output row index out of range");
                sp.out_row[m] = (int)orow;
                for (int e = 0; e < d->ntaps; ++e) {
                    const v2a_igemm_src& sr = d->src[d->taps[e].src];
                    long long off = -1;
                    bool in = true;
                    int sc[4];
                    for (int i = 0; i < 4; ++i) {
                        sc[i] = c[i] + d->taps[e].d[i];
                        in = in && sc[i] >= 0 && sc[i] < sr.dims[i];
                    }
                    if (in) off = ((((long long)sc[3] * sr.dims[2] + sc[2]) * sr.dims[1] + sc[1]) * sr.dims[0] + sc[0]) * sr.channels;
                    V2A_REQUIRE(off < ((long long)1 << 31), "igemm (small M): source offset out of range");
                    sp.off[m][e] = (int)off;
                }
            }
            sp.rows = (int)rows;
            sp.cout = d->cout < d->wrows ? d->cout : d->wrows;
            sp.ldc = d->ldc;
            sp.out_f32 = d->out_f32;
            sp.out_hi = reinterpret_cast<__nv_bfloat16*>(d->out_hi);
            sp.out_lo = reinterpret_cast<__nv_bfloat16*>(d->out_lo);
            sp.bias = d->bias;
            sp.residual = d->residual;
            sp.ld_res = d->ld_res;
            // 1, 2, 4 or 8 warps share one channel group: at most ~2 weight groups (of 8 elements) per lane and slice
            int slices = 1;
            while (slices < kSmallWarps && d->ktot / 8 > 64 * slices) slices *= 2;
            sp.slices = slices;
            sp.slice_groups = ceil_div(d->ktot / 8, slices);
            // four channels per warp share every activation read (wide layers); one channel per warp keeps the grid
            // large on narrow ones
            pl->small_ch = sp.cout >= 256 ? 4 : 1;
            pl->smallm = true;
            pl->grid = ceil_div(ceil_div(sp.cout, pl->small_ch), kSmallWarps / slices);
            pl->smem = 0;
            pl->zero_out = false;
            p.k_splits = 1;
            *out = pl;
            return 0;
        }
    }
    p.num_m_tiles = 1;
    for (int i = 0; i < 4; ++i) {
        p.tile_log2[i] = d->tile_log2[i];
        p.out_dims[i] = d->out_dims[i];
        p.ntile[i] = ceil_div(d->out_dims[i], 1 << d->tile_log2[i]);
        p.num_m_tiles *= p.ntile[i];
        p.rowvec_mul[i] = d->rowvec_mul[i];
        p.stats_mul[i] = d->stats_mul[i];
    }
    p.num_n_tiles = ceil_div(d->cout, d->block_n);
    p.block_n = d->block_n;
    p.passes = d->passes;
    p.b_tile_bytes = (uint32_t)d->block_n * kChunkK * 2;
    p.stage_bytes = (uint32_t)d->passes == 3 ? 2 * (kATileBytes + p.b_tile_bytes)
                                             : (kATileBytes + p.b_tile_bytes);
    const size_t overhead = 1024 /*align*/ + 256 /*barriers*/ + 8192 /*epilogue bias staging*/ +
                            8 * 32 * kSlabStride * 4 /*epilogue coalescing slabs*/;
    int stages = (int)((g_max_smem - overhead) / p.stage_bytes);
    if (stages > 8) stages = 8;
    if (stages < 2) {
        delete pl;
        V2A_REQUIRE(false, "igemm: block_n %d leaves room for %d pipeline stages", d->block_n, stages);
    }
    p.stages = stages;
    pl->smem = (size_t)stages * p.stage_bytes + overhead;
    p.cout = d->cout;
    p.ldc = d->ldc;
    p.out_f32 = d->out_f32;
    p.out_hi = reinterpret_cast<__nv_bfloat16*>(d->out_hi);
    p.out_lo = reinterpret_cast<__nv_bfloat16*>(d->out_lo);
    p.bias = d->bias;
    p.rowvec = d->rowvec;
    p.ld_rowvec = d->ld_rowvec;
    p.residual = d->residual;
    p.ld_res = d->ld_res;
    p.stats = d->stats;
    p.stats_ld = d->stats_ld;
    p.stats_replicas = d->stats_replicas > 0 ? d->stats_replicas : 1;
    p.stats_rep_stride = d->stats_rep_stride;
    p.a_fp16 = d->a_fp16;
    p.b_fp16 = d->b_fp16;
    if (d->out_pix_mul[0] | d->out_pix_mul[1] | d->out_pix_mul[2] | d->out_pix_mul[3]) {
        for (int i = 0; i < 4; ++i) p.out_mul[i] = d->out_pix_mul[i];
        p.out_off = d->out_pix_off;
    } else {                       // dense: rows in grid order, D0 fastest
        long long m = 1;
        for (int i = 0; i < 4; ++i) {
            p.out_mul[i] = m;
            m *= d->out_dims[i];
        }
        p.out_off = 0;
    }
    {
        const char* env = getenv("V2A_FUSE2");
        p.fuse2 = (d->passes == 3 && d->block_n <= 128 && !(env && atoi(env) == 0)) ? 1 : 0;
    }
    {
        const char* dbg = getenv("V2A_IGEMM_DEBUG");
        p.debug = dbg ? atoi(dbg) : 0;
    }

    {
        // linear tiling: boxes divide their dims, and once a box is narrower than its dim all higher dims have box 1
        // => tile m is rows [128 m, 128 m + 128) of the dense row space; rows of one tile share the embedding row and
        // the GroupNorm instance when the multipliers vanish on every dim the box spans
        const char* env = getenv("V2A_LINEAR");
        bool lin = !(env && atoi(env) == 0) &&
                   !(d->out_pix_mul[0] | d->out_pix_mul[1] | d->out_pix_mul[2] | d->out_pix_mul[3]);
        bool partial = false;
        for (int i = 0; i < 4; ++i) {
            const int box = 1 << d->tile_log2[i];
            if (d->out_dims[i] % box != 0 || (partial && box != 1)) lin = false;
            if (box != d->out_dims[i]) partial = true;
            if (d->tile_log2[i] > 0 && ((d->rowvec && d->rowvec_mul[i] != 0) || (d->stats && d->stats_mul[i] != 0)))
                lin = false;
        }
        p.linear = lin ? 1 : 0;
    }
    // tensor maps: A boxes follow the output tile box; every source shares it
    int rc = 0;
    for (int s = 0; s < d->nsrc && !rc; ++s) {
        const v2a_igemm_src& src = d->src[s];
        if (src.channels % 8 != 0 || !src.hi || (d->passes == 3 && !src.lo)) {
            set_error("igemm: source %d needs channels %% 8 == 0 and hi/lo planes", s);
            rc = 2;
            break;
        }
        uint64_t dims[5] = {(uint64_t)src.channels, (uint64_t)src.dims[0], (uint64_t)src.dims[1],
                            (uint64_t)src.dims[2], (uint64_t)src.dims[3]};
        uint32_t box[5] = {kChunkK, 1u << d->tile_log2[0], 1u << d->tile_log2[1],
                           1u << d->tile_log2[2], 1u << d->tile_log2[3]};
        rc = make_map(&p.a_hi[s], src.hi, 5, dims, box);
        if (!rc && d->passes == 3) rc = make_map(&p.a_lo[s], src.lo, 5, dims, box);
    }
    if (!rc) {
        uint64_t dims[2] = {(uint64_t)d->ktot, (uint64_t)d->wrows};
        uint32_t box[2] = {kChunkK, (uint32_t)d->block_n};
        if (!d->w_hi || (d->passes == 3 && !d->w_lo)) {
            set_error("igemm: missing weight planes");
            rc = 2;
        }
        if (!rc) rc = make_map(&p.b_hi, d->w_hi, 2, dims, box);
        if (!rc && d->passes == 3) rc = make_map(&p.b_lo, d->w_lo, 2, dims, box);
    }
    if (rc) {
        delete pl;
        return rc;
    }
    // split-K: a GEMM with few output tiles (policy layers: M = B*T = 1024 rows) leaves most SMs idle; share
    // each tile's K loop between CTAs that add fp32 partial sums with vector atomics.  Needs a plain fp32
    // output (no hi/lo planes, no GroupNorm sums, which want the finished value).
    p.k_splits = 1;
    p.kps = p.k_iters;
    pl->zero_out = false;
    {
        const int tiles_mn = p.num_m_tiles * p.num_n_tiles;
        const char* env = getenv("V2A_SPLIT_K");
        const bool allowed = d->out_f32 && !d->out_hi && !d->stats && !(env && atoi(env) == 0);
        if (allowed && tiles_mn * 2 <= g_num_sms && p.k_iters >= 8) {
            int want = g_num_sms / tiles_mn;
            if (want > 16) want = 16;
            if (want > p.k_iters / 4) want = p.k_iters / 4;
            if (env && atoi(env) > 1) want = atoi(env);
            if (d->split_stride > 0 && want > 16) want = 16;   // callers size their slice buffers for 16
            if (want > 1) {
                p.kps = ceil_div(p.k_iters, want);
                p.k_splits = ceil_div(p.k_iters, p.kps);
            }
        }
        p.split_stride = p.k_splits > 1 && d->split_stride > 0 ? d->split_stride : 0;
        if (p.split_stride > 0) {
            V2A_REQUIRE(!d->residual, "igemm: split_stride slices take no residual (the lead split would hold it alone)");
        } else if (p.k_splits > 1) {
            // accumulate-in-place (residual aliases the output): the atomics add onto what is there
            const bool in_place = d->residual == d->out_f32 && d->ld_res == d->ldc;
            if (in_place) p.residual = nullptr;
            pl->zero_out = !in_place;
            pl->zero_width = (size_t)(((d->cout + 15) / 16) * 16) * sizeof(float);
            pl->zero_rows = (int64_t)d->out_dims[0] * d->out_dims[1] * d->out_dims[2] * d->out_dims[3];
        }
    }
    const int total = p.num_m_tiles * p.num_n_tiles * p.k_splits;
    pl->grid = total < g_num_sms ? total : g_num_sms;
    // CTA pairs (thread-block clusters of 2) share each weight tile through TMA multicast: halves the B part of
    // the L2 -> shared-memory traffic, which bounds the Cout <= 256 layers.  One N tile only (every tile of
    // the launch then uses the same B), 3-pass plans, no split-K.
    p.cluster = 1;
    p.iters_per_cta = 0;
    {
        const char* env = getenv("V2A_CLUSTER");
        const bool want = !(env && atoi(env) == 0);
        // pairs over several N tiles are new with the cta_group::2 path and follow its switch
        const char* env2 = getenv("V2A_CTA2");
        const int mode2 = env2 ? atoi(env2) : 1;
        const bool multi_n = mode2 != 0 && mode2 != 3 && p.k_iters >= cta2_min_k() && d->block_n % 32 == 0;
        if (want && (p.num_n_tiles == 1 || multi_n) && p.k_splits == 1 && d->passes == 3 && d->block_n % 16 == 0 &&
            p.num_m_tiles >= 4 && g_num_sms >= 2) {
            int grid = (g_num_sms / 2) * 2;
            const int pairs_needed = ceil_div(p.num_m_tiles, 2) * p.num_n_tiles;
            if (grid > 2 * pairs_needed) grid = 2 * pairs_needed;
            p.cluster = 2;
            p.iters_per_cta = ceil_div(pairs_needed, grid / 2);
            pl->grid = grid;
            uint64_t dims[2] = {(uint64_t)d->ktot, (uint64_t)d->wrows};
            uint32_t box[2] = {kChunkK, (uint32_t)d->block_n / 2};
            rc = make_map(&p.bh_hi, d->w_hi, 2, dims, box);
            if (!rc) rc = make_map(&p.bh_lo, d->w_lo, 2, dims, box);
            if (rc) {
                delete pl;
                return rc;
            }
        }
    }
    // cta_group::2 pairs: ONE MMA spans the two SMs of a pair (M = 256), each SM reads its own A tile and only
    // HALF of the weight rows from shared memory: 14 KB of operand reads per SM per k step instead of 20 KB, which
    // turns the shared-memory-bound N = 128 layers tensor-bound (B = 16, K = 3456: 385 -> 530 TFLOP/s).  Needs the
    // fused split product (block_n <= 128); short-K launches (temporal / 1x1 convs, whose time is their epilogue)
    // lose to the lockstep of the two epilogues and stay on multicast pairs.  V2A_CTA2=0 disables, =2 forces.
    p.cta2 = 0;
    p.nacc_log2 = 1;
    {
        const char* env = getenv("V2A_CTA2");
        const int mode = env ? atoi(env) : 1;
        if (mode == 4 && p.cluster == 2 && d->block_n <= 128 && d->block_n % 32 == 0 && p.k_iters < 16) {
            // experiment: short-K launches as UNFUSED pairs with a 4-deep accumulator ring (4 x 128 TMEM columns),
            // so the two epilogues of a pair run up to 3 tiles behind the MMAs instead of in lockstep with them
            p.cta2 = 1;
            p.fuse2 = 0;
            p.nacc_log2 = 2;
            p.stage_bytes = 2 * (kATileBytes + p.b_tile_bytes / 2);
            int stages = (int)((g_max_smem - overhead) / p.stage_bytes);
            if (stages > 8) stages = 8;
            p.stages = stages;
            pl->smem = (size_t)stages * p.stage_bytes + overhead;
        }
        if (!p.cta2 && (mode != 0 || force_cta2) && p.cluster == 2 && d->block_n % 32 == 0 &&
            (p.k_iters >= cta2_min_k() || mode == 2 || force_cta2) &&
            (p.fuse2 || mode != 3)) {      // mode 3: fused (block_n <= 128) layers only (A/B probes)
            p.cta2 = 1;
            p.stage_bytes = 2 * (kATileBytes + p.b_tile_bytes / 2);
            int stages = (int)((g_max_smem - overhead) / p.stage_bytes);
            if (stages > 8) stages = 8;
            p.stages = stages;
            pl->smem = (size_t)stages * p.stage_bytes + overhead;
        }
    }
    {
        // epilogue class: fixed at compile time when nothing run-time-only is involved
        const char* fe = getenv("V2A_FAST_EPILOGUE");
        const int bits = (p.out_f32 ? 1 : 0) | (p.out_hi ? 2 : 0) | (p.stats ? 4 : 0);
        const bool fixed = p.k_splits == 1 && p.debug == 0 && (bits == 1 || bits == 2 || bits == 3 || bits == 5) &&
                           !(fe && atoi(fe) == 0);
        pl->epi = fixed ? bits : -1;
    }
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaSuccess;
        for (int c2 = 0; c2 < 2 && e == cudaSuccess; ++c2)
            for (int epi : {-1, 1, 2, 3, 5})
                if (e == cudaSuccess)
                    e = cudaFuncSetAttribute(igemm_fn(c2 != 0, epi), cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             g_max_smem);
        if (e == cudaSuccess)
            e = cudaFuncSetAttribute(igemm_dual_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, g_max_smem);
        if (e != cudaSuccess) {
            delete pl;
            V2A_CUDA_OK(e);
        }
        attr_set = true;
    }
    *out = pl;
    return 0;
}


struct DualPlan {
    DualParams dp;
    int grid;
    size_t smem;
};

static int dual_plan_create(const v2a_igemm_desc* ds, const v2a_igemm_desc* dt, int frames, int tiles_per_frame,
                            unsigned int* flags, DualPlan** out) {
    V2A_REQUIRE(flags != nullptr && frames >= 1 && tiles_per_frame >= 1, "igemm_dual: missing flags / geometry");
    IgemmPlan *ps = nullptr, *pt = nullptr;
    int rc = plan_create(ds, &ps, true);
    if (!rc) rc = plan_create(dt, &pt, true);
    auto fail = [&](const char* why) {
        delete ps;
        delete pt;
        set_error("igemm_dual: %s", why);
        return 2;
    };
    if (rc) {
        delete ps;
        delete pt;
        return rc;
    }
    const IgemmParams &a = ps->p, &b = pt->p;
    for (const IgemmParams* p : {&a, &b})
        if (!(p->cta2 && p->fuse2 && p->cluster == 2 && p->num_n_tiles == 1 && p->k_splits == 1 && p->passes == 3 &&
              p->nacc_log2 == 1 && !p->a_fp16 && !p->b_fp16))
            return fail("both programs must be cta_group::2 pair launches with the fused split product "
                        "(3 passes, block_n <= 128 and a multiple of 32, one N tile, no split-K)");
    if (a.block_n != b.block_n || a.stage_bytes != b.stage_bytes || a.stages != b.stages)
        return fail("the two programs must share block_n (one operand ring serves both)");
    if (a.num_m_tiles != b.num_m_tiles) return fail("the two programs must tile the same row space");
    if (a.num_m_tiles % (tiles_per_frame * frames) != 0)
        return fail("tiles must be whole (sample, frame) blocks");
    if (!a.out_hi || a.out_f32 || a.stats || a.residual || a.rowvec || !b.out_f32 || b.out_hi)
        return fail("the spatial program writes hi/lo planes only (bias allowed), the temporal program fp32 "
                    "(+ bias / embedding row / residual / GroupNorm sums): the kernel's epilogues are specialised on that");
    DualPlan* pl = new DualPlan();
    memset(&pl->dp, 0, sizeof(pl->dp));
    pl->dp.g[0] = a;
    pl->dp.g[1] = b;
    pl->dp.flags = flags;
    pl->dp.tpf = tiles_per_frame;
    pl->dp.frames = frames;
    pl->dp.tiles = a.num_m_tiles;
    pl->dp.mp = ceil_div(a.num_m_tiles, 2);
    int grid = (g_num_sms / 2) * 2;
    if (grid > 2 * pl->dp.mp) grid = 2 * pl->dp.mp;
    const int pairs = grid / 2;
    // every dependency of T(q) (pair-tiles up to q + ceil(tpf / 2)) must belong to an EARLIER iteration of its pair
    // than the one that issues T(q): lag >= ceil(tpf / 2) + pairs; one more round of pairs gives the producers'
    // epilogues time to publish, so the temporal producer normally finds its flags set
    const char* env = getenv("V2A_DUAL_LAG");
    const int rounds = env ? atoi(env) : 2;
    pl->dp.lag = (tiles_per_frame + 1) / 2 + 1 + pairs * (rounds < 1 ? 1 : rounds);
    pl->dp.iters = ceil_div(pl->dp.mp + pl->dp.lag, pairs);
    {
        // measured (gpurun_out/r2c12_ab.txt, B = 16 forward, same box): fused two-accumulator form 90.28 ms, this
        // four-accumulator unfused form 91.79 ms (the N = 128 layers are shared-memory-operand bound: three MMAs per k
        // step read 72 KB per SM instead of 56 KB) -- kept as an opt-in probe
        const char* a4 = getenv("V2A_DUAL_ACC4");
        pl->dp.acc4 = (a4 && atoi(a4) == 1) ? 1 : 0;
    }
    pl->grid = grid;
    pl->smem = ps->smem;
    delete ps;
    delete pt;
    *out = pl;
    return 0;
}

}  // namespace v2a

extern "C" {

int v2a_igemm_dual_plan_create(const v2a_igemm_desc* spatial, const v2a_igemm_desc* temporal, int frames,
                               int tiles_per_frame, void* flags, void** plan_out) {
    v2a::DualPlan* pl = nullptr;
    int rc = v2a::dual_plan_create(spatial, temporal, frames, tiles_per_frame, (unsigned int*)flags, &pl);
    if (rc) return rc;
    *plan_out = pl;
    return 0;
}

int v2a_igemm_dual_plan_run(void* plan, void* stream) {
    v2a::DualPlan* pl = reinterpret_cast<v2a::DualPlan*>(plan);
    V2A_CUDA_OK(cudaMemsetAsync(pl->dp.flags, 0, sizeof(unsigned int) * (size_t)pl->dp.tiles, (cudaStream_t)stream));
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)pl->grid);
    cfg.blockDim = dim3(v2a::kThreads);
    cfg.dynamicSmemBytes = pl->smem;
    cfg.stream = (cudaStream_t)stream;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = v2a::pdl_enabled() ? 2 : 1;
    V2A_CUDA_OK(cudaLaunchKernelEx(&cfg, v2a::igemm_dual_kernel, pl->dp));
    V2A_CUDA_OK(cudaGetLastError());
    v2a::g_launches.fetch_add(1);
    return 0;
}

void v2a_igemm_dual_plan_destroy(void* plan) { delete reinterpret_cast<v2a::DualPlan*>(plan); }

const char* v2a_last_error(void) { return v2a::get_error(); }
int v2a_version(void) { return 100; }
int64_t v2a_launch_count(void) { return v2a::g_launches.load(); }

int v2a_igemm_plan_create(const v2a_igemm_desc* desc, void** plan_out) {
    v2a::IgemmPlan* pl = nullptr;
    int rc = v2a::plan_create(desc, &pl);
    if (rc) return rc;
    *plan_out = pl;
    return 0;
}

int v2a_igemm_plan_run(void* plan, void* stream) {
    v2a::IgemmPlan* pl = reinterpret_cast<v2a::IgemmPlan*>(plan);
    if (pl->smallm) {
        const int rc = pl->sp.rows <= 4 ? 4 : (pl->sp.rows <= 8 ? 8 : 16);
        const dim3 blk(v2a::kSmallWarps * 32);
        cudaStream_t st = (cudaStream_t)stream;
        cudaError_t le;
        if (pl->small_ch == 4) {
            if (rc == 4) le = launch_maybe_pdl(v2a::igemm_smallm_kernel<4, 4>, dim3(pl->grid), blk, 0, st, pl->sp);
            else if (rc == 8) le = launch_maybe_pdl(v2a::igemm_smallm_kernel<8, 4>, dim3(pl->grid), blk, 0, st, pl->sp);
            else le = launch_maybe_pdl(v2a::igemm_smallm_kernel<16, 4>, dim3(pl->grid), blk, 0, st, pl->sp);
        } else {
            if (rc == 4) le = launch_maybe_pdl(v2a::igemm_smallm_kernel<4, 1>, dim3(pl->grid), blk, 0, st, pl->sp);
            else if (rc == 8) le = launch_maybe_pdl(v2a::igemm_smallm_kernel<8, 1>, dim3(pl->grid), blk, 0, st, pl->sp);
            else le = launch_maybe_pdl(v2a::igemm_smallm_kernel<16, 1>, dim3(pl->grid), blk, 0, st, pl->sp);
        }
        V2A_CUDA_OK(le);
        V2A_CUDA_OK(cudaGetLastError());
        v2a::g_launches.fetch_add(1);
        return 0;
    }
    if (pl->zero_out)
        V2A_CUDA_OK(cudaMemset2DAsync(pl->p.out_f32, (size_t)pl->p.ldc * sizeof(float), 0, pl->zero_width,
                                      (size_t)pl->zero_rows, (cudaStream_t)stream));
    if (pl->p.cluster > 1) {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3((unsigned)pl->grid);
        cfg.blockDim = dim3(v2a::kThreads);
        cfg.dynamicSmemBytes = pl->smem;
        cfg.stream = (cudaStream_t)stream;
        cudaLaunchAttribute attr[2];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = 2;
        attr[0].val.clusterDim.y = 1;
        attr[0].val.clusterDim.z = 1;
        attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[1].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = attr;
        cfg.numAttrs = v2a::pdl_enabled() ? 2 : 1;
        V2A_CUDA_OK(cudaLaunchKernelEx(&cfg, v2a::igemm_fn(pl->p.cta2 != 0, pl->epi), pl->p));
    } else {
        V2A_CUDA_OK(launch_maybe_pdl(v2a::igemm_fn(false, pl->epi), dim3(pl->grid), dim3(v2a::kThreads), pl->smem,
                                     (cudaStream_t)stream, pl->p));
    }
    V2A_CUDA_OK(cudaGetLastError());
    v2a::g_launches.fetch_add(1);
    return 0;
}

int v2a_igemm_plan_k_splits(void* plan) { return reinterpret_cast<v2a::IgemmPlan*>(plan)->p.k_splits; }

void v2a_igemm_plan_destroy(void* plan) { delete reinterpret_cast<v2a::IgemmPlan*>(plan); }

}  // extern "C"
