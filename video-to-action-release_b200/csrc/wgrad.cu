// Weight-gradient GEMM on tcgen05 tensor cores, straight from channels-last activations.
//
//   dW^T[(unit, ci), co] = sum_pixels  x[pixel + tap(unit), chunk(unit)*64 + ci] * dy[pixel, co]
//
// The reduction index is the PIXEL, which is the slow index of both channels-last operands, so both
// are fed to the tensor core as MN-major shared-memory tiles (instruction descriptor a_major = b_major
// = 1): a TMA box of (64 channels x 64 pixels) IS the canonical MN-major SWIZZLE_128B tile -- one
// 128-byte row per pixel (= per k), 8-row swizzle atoms 1024 B apart (SBO), further 64-channel blocks
// 8 KB apart (LBO).  No transposed copy of dy and no im2col of x ever exists: the tap shift is the TMA
// box coordinate, and out-of-bounds coordinates are zero-filled = the convolution's zero padding.
//
//   M tile = 128 rows = two "units" (a unit = one tap x one 64-channel chunk of one source)
//   N tile = block_n output channels (multiple of 64, <= 256)
//   K      = all output pixels, 64 per pipeline step, shared between `k_splits` CTAs (fp32 RED)
//
// replaces the autograd weight gradient of
//   torchvision ResNet18 convs of the observation encoder   diffusion_policy/common/vision_nets.py:29-39
//   SpatialSoftmax keypoint conv                            diffusion_policy/common/base_nets.py:183
#include "common.cuh"
#include "../../include/v2a_b200.h"

#include <atomic>
#include <cstring>

namespace v2a {
extern std::atomic<int64_t> g_launches;
int make_tensor_map_bf16(CUtensorMap* m, const void* base, int rank, const uint64_t* dims, const uint32_t* box);

constexpr int kWgThreads = 256;      // 4 control warps + 4 epilogue warps
constexpr int kWgUnitBytes = 64 * 128;   // 64 pixels x 64 channels bf16
constexpr int kWgMaxUnits = V2A_WGRAD_MAX_UNITS;

struct alignas(64) WgradParams {
    CUtensorMap a_hi[V2A_MAX_SRC];
    CUtensorMap a_lo[V2A_MAX_SRC];
    CUtensorMap b_hi;
    CUtensorMap b_lo;
    int nunits;
    short unit_src[kWgMaxUnits];
    short unit_chunk[kWgMaxUnits];
    short unit_d[kWgMaxUnits][4];
    int box_log2[4];
    int nbox[4];
    int k_iters, kps, k_splits;
    int num_m_tiles, num_n_tiles;
    int block_n, nb, passes, stages;
    uint32_t stage_bytes;
    int cout;
    float* out;
    int ld_out;
    int x_fp16;     // x AND dy planes are fp16 (hi, lo) pairs: one MMA takes a single operand format
};

// MN-major, 128-byte-swizzled operand: rows of 64 channels (128 B) per k, 8-row atoms SBO = 1024 B apart,
// 64-channel blocks LBO = 8192 B apart.
__device__ __forceinline__ uint64_t umma_desc_mn_sw128(uint32_t smem_addr, uint32_t lbo_bytes) {
    const uint64_t hi = (uint64_t)(1024u >> 4) | (1ull << 14) | (2ull << 29);
    const uint64_t lo = (uint64_t)((smem_addr & 0x3FFFFu) >> 4) | ((uint64_t)(lbo_bytes >> 4) << 16);
    return (hi << 32) | lo;
}

__global__ void __launch_bounds__(kWgThreads, 1) wgrad_kernel(const __grid_constant__ WgradParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int S = p.stages;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)S * p.stage_bytes);
    uint64_t* full_bar = bars;
    uint64_t* empty_bar = bars + S;
    uint64_t* tfull_bar = bars + 2 * S;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * S + 1);

    int t = blockIdx.x;
    const int split = t % p.k_splits;
    t /= p.k_splits;
    const int n_idx = t % p.num_n_tiles;
    const int m_idx = t / p.num_n_tiles;
    const int n0 = n_idx * p.block_n;
    const int kb = split * p.kps;
    const int ke = min(kb + p.kps, p.k_iters);
    const int units_here = min(2, p.nunits - 2 * m_idx);
    const uint32_t a_plane = 2 * kWgUnitBytes;                 // both units of one plane
    const uint32_t b_plane = (uint32_t)p.nb * kWgUnitBytes;

    pdl_trigger();
    if (warp == 0 && lane == 0) {
        for (int s = 0; s < S; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], 1);
        }
        mbar_init(tfull_bar, 1);
        fence_mbar_init();
    } else if (warp == 1 && lane == 0) {
        tma_prefetch_desc(&p.a_hi[0]);
        tma_prefetch_desc(&p.b_hi);
    } else if (warp == 2) {
        tmem_alloc(tmem_slot, 256);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    pdl_wait();          // set-up above overlaps the previous kernel's tail (programmatic dependent launch)
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0 && lane == 0) {
        // ===================== TMA producer =====================
        const uint32_t tx_bytes = (uint32_t)(units_here + p.nb) * kWgUnitBytes * (p.passes == 3 ? 2 : 1);
        int stage = 0;
        uint32_t phase = 0;
        for (int kit = kb; kit < ke; ++kit) {
            int o[4], idx = kit;
#pragma unroll
            for (int d = 0; d < 4; ++d) {
                o[d] = (idx % p.nbox[d]) << p.box_log2[d];
                idx /= p.nbox[d];
            }
            mbar_wait(&empty_bar[stage], phase ^ 1, 500 + stage);
            uint8_t* st = smem + (size_t)stage * p.stage_bytes;
            mbar_arrive_expect_tx(&full_bar[stage], tx_bytes);
            for (int u = 0; u < units_here; ++u) {
                const int unit = 2 * m_idx + u;
                const int src = p.unit_src[unit];
                const int c0 = p.unit_chunk[unit] * 64;
                const int c1 = o[0] + p.unit_d[unit][0], c2 = o[1] + p.unit_d[unit][1];
                const int c3 = o[2] + p.unit_d[unit][2], c4 = o[3] + p.unit_d[unit][3];
                tma_load_5d(st + u * kWgUnitBytes, &p.a_hi[src], &full_bar[stage], c0, c1, c2, c3, c4);
                if (p.passes == 3)
                    tma_load_5d(st + a_plane + u * kWgUnitBytes, &p.a_lo[src], &full_bar[stage], c0, c1, c2, c3, c4);
            }
            uint8_t* sb = st + (p.passes == 3 ? 2 : 1) * a_plane;   // one-pass stages hold no lo planes
            for (int j = 0; j < p.nb; ++j) {
                tma_load_5d(sb + j * kWgUnitBytes, &p.b_hi, &full_bar[stage], n0 + 64 * j, o[0], o[1], o[2], o[3]);
                if (p.passes == 3)
                    tma_load_5d(sb + b_plane + j * kWgUnitBytes, &p.b_lo, &full_bar[stage], n0 + 64 * j, o[0], o[1],
                                o[2], o[3]);
            }
            if (++stage == S) { stage = 0; phase ^= 1; }
        }
    } else if (warp == 1 && lane == 0) {
        // ===================== MMA issuer =====================
        const uint32_t idesc = umma_idesc_16(128, p.block_n, p.x_fp16, p.x_fp16) | (1u << 15) | (1u << 16);   // A, B MN-major
        int stage = 0;
        uint32_t phase = 0;
        for (int kit = kb; kit < ke; ++kit) {
            mbar_wait(&full_bar[stage], phase, 600 + stage);
            tc_fence_after();
            const uint32_t sa = smem_u32(smem + (size_t)stage * p.stage_bytes);
            const uint32_t sb = sa + (p.passes == 3 ? 2 : 1) * a_plane;
            const uint64_t a_hi = umma_desc_mn_sw128(sa, kWgUnitBytes);
            const uint64_t a_lo = umma_desc_mn_sw128(sa + a_plane, kWgUnitBytes);
            const uint64_t b_hi = umma_desc_mn_sw128(sb, kWgUnitBytes);
            const uint64_t b_lo = umma_desc_mn_sw128(sb + b_plane, kWgUnitBytes);
            // one k step = 16 pixels = 16 rows of 128 B = 2048 B (>> 4 = 128 in the descriptor's address field)
            if (p.passes == 3 && p.block_n <= 128) {
                // [b_hi blocks | b_lo blocks] are consecutive 64-channel MN atoms (LBO apart): a_hi x b_hi and
                // a_hi x b_lo are ONE MMA with N = 2*block_n filling two accumulator halves (summed in the epilogue)
                const uint32_t idesc2 = umma_idesc_16(128, 2 * p.block_n, p.x_fp16, p.x_fp16) | (1u << 15) | (1u << 16);
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    umma_bf16(tmem_base, a_hi + 128 * k, b_hi + 128 * k, idesc2, (kit > kb || k != 0));
#pragma unroll
                for (int k = 0; k < 4; ++k) umma_bf16(tmem_base, a_lo + 128 * k, b_hi + 128 * k, idesc, 1);
            } else if (p.passes == 3) {
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    umma_bf16(tmem_base, a_lo + 128 * k, b_hi + 128 * k, idesc, (kit > kb || k != 0));
#pragma unroll
                for (int k = 0; k < 4; ++k) umma_bf16(tmem_base, a_hi + 128 * k, b_lo + 128 * k, idesc, 1);
#pragma unroll
                for (int k = 0; k < 4; ++k) umma_bf16(tmem_base, a_hi + 128 * k, b_hi + 128 * k, idesc, 1);
            } else {
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    umma_bf16(tmem_base, a_hi + 128 * k, b_hi + 128 * k, idesc, (kit > kb || k != 0));
            }
            umma_commit(&empty_bar[stage]);
            if (kit == ke - 1) umma_commit(tfull_bar);
            if (++stage == S) { stage = 0; phase ^= 1; }
        }
    } else if (warp >= 4 && kb < ke) {
        // ===================== epilogue: fp32 RED into the (zeroed / accumulating) output =====================
        const int quad = warp & 3;
        const int row = quad * 32 + lane;            // TMEM lane = M index: unit (row / 64), channel (row % 64)
        const int unit = 2 * m_idx + (row >> 6);
        const bool valid = unit < p.nunits;
        float* orow = p.out + ((int64_t)unit * 64 + (row & 63)) * p.ld_out + n0;
        mbar_wait(tfull_bar, 0, 700);
        tc_fence_after();
        const uint32_t t_row = tmem_base + ((uint32_t)(quad * 32) << 16);
        const bool fused = p.passes == 3 && p.block_n <= 128;
        for (int c = 0; c < p.block_n; c += 16) {
            uint32_t r[16];
            tmem_ld16(t_row + c, r);
            if (fused) {
                uint32_t r2[16];
                tmem_ld16(t_row + p.block_n + c, r2);
                tmem_ld_wait16(r);
                tmem_ld_wait16(r2);
#pragma unroll
                for (int j = 0; j < 16; ++j) r[j] = __float_as_uint(__uint_as_float(r[j]) + __uint_as_float(r2[j]));
            } else {
                tmem_ld_wait16(r);
            }
            if (valid) {
#pragma unroll
                for (int q = 0; q < 4; ++q)
                    if (n0 + c + 4 * q < p.cout)
                        atomicAdd(reinterpret_cast<float4*>(orow + c) + q,
                                  make_float4(__uint_as_float(r[4 * q]), __uint_as_float(r[4 * q + 1]),
                                              __uint_as_float(r[4 * q + 2]), __uint_as_float(r[4 * q + 3])));
            }
        }
        tc_fence_before();
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 256);
    }
}

// dW[co][ci][tap] (+)= dWT[(tap * nchunk + chunk) * 64 + ci % 64][co]   (ci = chunk * 64 + ci % 64); dW rows are
// ld_dw apart (a column window of a wider [cout][cin_total * taps] gradient when the conv input is a concat)
__global__ void wgrad_scatter_kernel(const float* __restrict__ wt, int ld, int cout, int cin, int ntaps, int nchunk,
                                     float* __restrict__ dw, int64_t ld_dw, int accumulate) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t per = (int64_t)cin * ntaps;
    if (i >= cout * per) return;
    const int co = (int)(i / per);
    const int rem = (int)(i - co * per);
    const int tap = rem % ntaps;
    const int ci = rem / ntaps;
    const int64_t r = ((int64_t)tap * nchunk + (ci >> 6)) * 64 + (ci & 63);
    const float v = wt[r * ld + co];
    float* d = dw + co * ld_dw + rem;
    *d = accumulate ? *d + v : v;
}

struct WgradPlan {
    WgradParams p;
    int grid;
    size_t smem;
};

static int wgrad_plan_create(const v2a_wgrad_desc* d, WgradPlan** out) {
    V2A_REQUIRE(d->nsrc >= 1 && d->nsrc <= V2A_MAX_SRC, "wgrad: nsrc %d out of range", d->nsrc);
    V2A_REQUIRE(d->nunits >= 1 && d->nunits <= kWgMaxUnits, "wgrad: nunits %d out of range", d->nunits);
    V2A_REQUIRE(d->passes == 1 || d->passes == 3, "wgrad: passes must be 1 or 3");
    V2A_REQUIRE(d->cout >= 4 && d->cout % 4 == 0 && d->ld_out % 4 == 0 && d->ld_out >= d->cout,
                "wgrad: cout %d / ld_out %d must be multiples of 4", d->cout, d->ld_out);
    V2A_REQUIRE(d->out != nullptr && d->dy.hi && (d->passes == 1 || d->dy.lo), "wgrad: missing buffers");
    int bl = 0;
    for (int i = 0; i < 4; ++i) {
        V2A_REQUIRE(d->box_log2[i] >= 0 && d->box_log2[i] <= 6 && d->dy.dims[i] >= 1, "wgrad: bad box / dims");
        bl += d->box_log2[i];
    }
    V2A_REQUIRE(bl == 6, "wgrad: the K box must hold 64 pixels (sum box_log2 = %d)", bl);
    WgradPlan* pl = new WgradPlan();
    memset(&pl->p, 0, sizeof(pl->p));
    WgradParams& p = pl->p;
    p.nunits = d->nunits;
    for (int u = 0; u < d->nunits; ++u) {
        const v2a_wgrad_unit& un = d->units[u];
        if (!(un.src >= 0 && un.src < d->nsrc && un.chunk >= 0 && un.chunk * 64 < d->src[un.src].channels)) {
            delete pl;
            V2A_REQUIRE(false, "wgrad: bad unit %d", u);
        }
        p.unit_src[u] = (short)un.src;
        p.unit_chunk[u] = (short)un.chunk;
        for (int i = 0; i < 4; ++i) p.unit_d[u][i] = (short)un.d[i];
    }
    p.k_iters = 1;
    for (int i = 0; i < 4; ++i) {
        p.box_log2[i] = d->box_log2[i];
        p.nbox[i] = ceil_div(d->dy.dims[i], 1 << d->box_log2[i]);
        p.k_iters *= p.nbox[i];
    }
    p.passes = d->passes;
    p.cout = d->cout;
    p.out = d->out;
    p.ld_out = d->ld_out;
    p.x_fp16 = d->x_fp16;
    int bn = ((d->cout + 63) / 64) * 64;
    if (bn > 256) bn = 256;
    p.block_n = bn;
    p.nb = bn / 64;
    p.num_n_tiles = ceil_div(d->cout, bn);
    p.num_m_tiles = ceil_div(d->nunits, 2);
    p.stage_bytes = (uint32_t)(2 + p.nb) * kWgUnitBytes * (d->passes == 3 ? 2 : 1);
    int dev = 0, sms = 0, max_smem = 0;
    V2A_CUDA_OK(cudaGetDevice(&dev));
    V2A_CUDA_OK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    V2A_CUDA_OK(cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
    const size_t overhead = 1024 + 256;
    int stages = (int)((max_smem - overhead) / p.stage_bytes);
    if (stages > 6) stages = 6;
    if (stages < 2) {
        delete pl;
        V2A_REQUIRE(false, "wgrad: no room for 2 pipeline stages");
    }
    p.stages = stages;
    pl->smem = (size_t)stages * p.stage_bytes + overhead;
    // split the pixel reduction so that the grid covers the machine about twice
    const int tiles = p.num_m_tiles * p.num_n_tiles;
    int want = ceil_div(2 * sms, tiles);
    if (want > p.k_iters / 2) want = p.k_iters / 2;
    if (want < 1) want = 1;
    p.kps = ceil_div(p.k_iters, want);
    p.k_splits = ceil_div(p.k_iters, p.kps);
    pl->grid = tiles * p.k_splits;

    int rc = 0;
    uint32_t box[5] = {64, 1u << d->box_log2[0], 1u << d->box_log2[1], 1u << d->box_log2[2], 1u << d->box_log2[3]};
    for (int s = 0; s < d->nsrc && !rc; ++s) {
        const v2a_igemm_src& src = d->src[s];
        if (src.channels % 8 != 0 || !src.hi || (d->passes == 3 && !src.lo)) {
            set_error("wgrad: source %d needs channels %% 8 == 0 and hi/lo planes", s);
            rc = 2;
            break;
        }
        uint64_t dims[5] = {(uint64_t)src.channels, (uint64_t)src.dims[0], (uint64_t)src.dims[1],
                            (uint64_t)src.dims[2], (uint64_t)src.dims[3]};
        rc = make_tensor_map_bf16(&p.a_hi[s], src.hi, 5, dims, box);
        if (!rc && d->passes == 3) rc = make_tensor_map_bf16(&p.a_lo[s], src.lo, 5, dims, box);
    }
    if (!rc) {
        if (d->dy.channels % 8 != 0) {
            set_error("wgrad: dy channels must be a multiple of 8");
            rc = 2;
        }
        uint64_t dims[5] = {(uint64_t)d->dy.channels, (uint64_t)d->dy.dims[0], (uint64_t)d->dy.dims[1],
                            (uint64_t)d->dy.dims[2], (uint64_t)d->dy.dims[3]};
        if (!rc) rc = make_tensor_map_bf16(&p.b_hi, d->dy.hi, 5, dims, box);
        if (!rc && d->passes == 3) rc = make_tensor_map_bf16(&p.b_lo, d->dy.lo, 5, dims, box);
    }
    if (rc) {
        delete pl;
        return rc;
    }
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem);
        if (e != cudaSuccess) {
            delete pl;
            V2A_CUDA_OK(e);
        }
        attr_set = true;
    }
    *out = pl;
    return 0;
}

}  // namespace v2a

extern "C" {

int v2a_wgrad_plan_create(const v2a_wgrad_desc* desc, void** plan_out) {
    v2a::WgradPlan* pl = nullptr;
    int rc = v2a::wgrad_plan_create(desc, &pl);
    if (rc) return rc;
    *plan_out = pl;
    return 0;
}

int v2a_wgrad_plan_run(void* plan, void* stream) {
    v2a::WgradPlan* pl = reinterpret_cast<v2a::WgradPlan*>(plan);
    V2A_CUDA_OK(launch_maybe_pdl(v2a::wgrad_kernel, dim3(pl->grid), dim3(v2a::kWgThreads), pl->smem, (cudaStream_t)stream,
                                 pl->p));
    V2A_CUDA_OK(cudaGetLastError());
    v2a::g_launches.fetch_add(1);
    return 0;
}

int v2a_wgrad_plan_k_splits(void* plan) { return reinterpret_cast<v2a::WgradPlan*>(plan)->p.k_splits; }

void v2a_wgrad_plan_destroy(void* plan) { delete reinterpret_cast<v2a::WgradPlan*>(plan); }

int v2a_wgrad_scatter(const float* wt, int ld, int cout, int cin, int ntaps, float* dw, int64_t ld_dw,
                      int accumulate, void* stream) {
    V2A_REQUIRE(cout >= 1 && cin >= 1 && ntaps >= 1 && ld_dw >= (int64_t)cin * ntaps, "wgrad_scatter: bad shape");
    const int64_t n = (int64_t)cout * cin * ntaps;
    v2a::wgrad_scatter_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        wt, ld, cout, cin, ntaps, (cin + 63) / 64, dw, ld_dw, accumulate);
    V2A_CUDA_OK(cudaGetLastError());
    v2a::g_launches.fetch_add(1);
    return 0;
}

}  // extern "C"
