// Shared device/host helpers for the v2a_b200 kernels (sm_100a only).
//
// Everything here is raw PTX for the Blackwell async machinery (mbarrier, TMA
// bulk-tensor loads, tcgen05 MMA / TMEM) plus the bf16 hi/lo split that every
// tensor-core contraction in this library uses to reach fp32-class accuracy.
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <utility>

namespace v2a {

// ----------------------------------------------------------------------------
// error plumbing (host): no exception ever crosses the C ABI
// ----------------------------------------------------------------------------
void set_error(const char* fmt, ...);
const char* get_error();

#define V2A_CUDA_OK(expr)                                                        \
    do {                                                                         \
        cudaError_t _e = (expr);                                                 \
        if (_e != cudaSuccess) {                                                 \
            v2a::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), \
                           __FILE__, __LINE__);                                  \
            return 1;                                                            \
        }                                                                        \
    } while (0)

#define V2A_REQUIRE(cond, ...)                                                   \
    do {                                                                         \
        if (!(cond)) {                                                           \
            v2a::set_error(__VA_ARGS__);                                         \
            return 2;                                                            \
        }                                                                        \
    } while (0)

static inline int ceil_div(int a, int b) { return (a + b - 1) / b; }

// ----------------------------------------------------------------------------
// bf16 hi/lo split: x ~= hi + lo with |x - hi - lo| <= 2^-17 |x|
// ----------------------------------------------------------------------------
__device__ __forceinline__ void split_bf16(float x, __nv_bfloat16& hi, __nv_bfloat16& lo) {
    hi = __float2bfloat16_rn(x);
    lo = __float2bfloat16_rn(x - __bfloat162float(hi));
}

__device__ __forceinline__ uint32_t pack_bf16x2(__nv_bfloat16 a, __nv_bfloat16 b) {
    return (uint32_t)__bfloat16_as_ushort(a) | ((uint32_t)__bfloat16_as_ushort(b) << 16);
}

// two floats -> packed bf16x2 (a in the low half) with ONE cvt.rn.bf16x2.f32: the scalar conversion is an
// XU-pipe instruction (quarter rate) and made the GroupNorm-apply kernel transcendental-bound
__device__ __forceinline__ uint32_t pack2_bf16_rn(float a, float b) {
    uint32_t r;
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a));
    return r;
}
// (hi, lo) planes of two floats: hi = rn(x), lo = rn(x - hi); a bf16 widens to fp32 by a 16-bit shift
__device__ __forceinline__ void split2(float a, float b, uint32_t& hi, uint32_t& lo) {
    hi = pack2_bf16_rn(a, b);
    const float ha = __uint_as_float(hi << 16), hb = __uint_as_float(hi & 0xffff0000u);
    lo = pack2_bf16_rn(a - ha, b - hb);
}
// split 8 floats -> one uint4 of hi and one uint4 of lo
__device__ __forceinline__ void split8(const float* v, uint4& hi, uint4& lo) {
    split2(v[0], v[1], hi.x, lo.x);
    split2(v[2], v[3], hi.y, lo.y);
    split2(v[4], v[5], hi.z, lo.z);
    split2(v[6], v[7], hi.w, lo.w);
}

// fp16 (hi, lo) planes: hi = rn_f16(x), lo = rn_f16(x - hi).  22+ significant bits (|x - hi - lo| <= 2^-23 |x| while
// lo stays a normal fp16, i.e. |x| >~ 0.25; absolute error <= 3e-8 below that): fp32-class operands for the
// observation encoder's FORWARD convolutions, whose inputs (GroupNorm outputs, weights) sit well inside the fp16
// range.  The bf16 split (8 + 8 bits, error 2^-17) leaves ~2e-5 of forward error, enough to flip ~1e-5 of the
// ReLU masks against the reference and move parameter gradients by > 1e-3; gradients themselves keep bf16 planes
// (range).  tcgen05 kind::f16 takes either format per operand (instruction-descriptor a_format / b_format).
__device__ __forceinline__ void split2_f16(float a, float b, uint32_t& hi, uint32_t& lo) {
    const __half2 h = __floats2half2_rn(a, b);
    const float2 hf = __half22float2(h);
    const __half2 l = __floats2half2_rn(a - hf.x, b - hf.y);
    hi = *reinterpret_cast<const uint32_t*>(&h);
    lo = *reinterpret_cast<const uint32_t*>(&l);
}
__device__ __forceinline__ void split8_f16(const float* v, uint4& hi, uint4& lo) {
    split2_f16(v[0], v[1], hi.x, lo.x);
    split2_f16(v[2], v[3], hi.y, lo.y);
    split2_f16(v[4], v[5], hi.z, lo.z);
    split2_f16(v[6], v[7], hi.w, lo.w);
}
// format-selecting split (fmt: 0 bf16 planes, 1 fp16 planes)
__device__ __forceinline__ void split8_fmt(const float* v, uint4& hi, uint4& lo, int fmt) {
    if (fmt) split8_f16(v, hi, lo);
    else split8(v, hi, lo);
}

__device__ __forceinline__ float silu_f(float x) { return x / (1.0f + expf(-x)); }

// mish(x) = x * tanh(softplus(x)); softplus threshold 20 as torch
__device__ __forceinline__ float mish_f(float x) {
    float sp = x > 20.0f ? x : log1pf(expf(x));
    return x * tanhf(sp);
}
// d mish / dx
__device__ __forceinline__ float mish_grad_f(float x) {
    float sp = x > 20.0f ? x : log1pf(expf(x));
    float th = tanhf(sp);
    float sg = 1.0f / (1.0f + expf(-x));
    return th + x * (1.0f - th * th) * sg;
}

// streaming 16-byte read-only load / L2 prefetch of global data that is read exactly once
__device__ __forceinline__ float4 ld_nc_f4(const float* p) {
    float4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0, %1, %2, %3}, [%4];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
                 : "l"(p));
    return v;
}
// Programmatic dependent launch (PDL): a kernel launched with cudaLaunchAttributeProgrammaticStreamSerialization may
// become resident while its predecessor in the stream is still running.  `pdl_trigger` (predecessor side) lets the
// successor's CTAs be scheduled once every CTA of this grid has called it or exited; `pdl_wait` (successor side) blocks
// until the predecessor grid has completed and its memory is visible -- everything before it may only touch data no
// earlier kernel of the chain writes (static weights).  Both are no-ops in a kernel launched the ordinary way.
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

__device__ __forceinline__ void prefetch_l2(const void* p) {
    asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
}

// ----------------------------------------------------------------------------
// shared-memory address + mbarrier
// ----------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return (uint32_t)__cvta_generic_to_shared(p);
}

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count)
                 : "memory");
}
__device__ __forceinline__ void fence_mbar_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
                 "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
// Bounded wait: a pipeline bug must abort the kernel (trap -> launch error the
// host reports) instead of hanging the GPU.  ~2 s at 2 GHz.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity, int tag) {
    if (mbar_try_wait(bar, parity)) return;
    long long t0 = clock64();
    while (!mbar_try_wait(bar, parity)) {
        if (clock64() - t0 > 4000000000LL) {
            printf("v2a: mbarrier timeout tag=%d block=%d thread=%d parity=%u\n", tag,
                   (int)blockIdx.x, (int)threadIdx.x, parity);
            __trap();
        }
    }
}

// ----------------------------------------------------------------------------
// TMA bulk-tensor loads (tiled mode), completion on an mbarrier
// ----------------------------------------------------------------------------
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* m) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(m)) : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* m, uint64_t* bar,
                                            int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes"
        " [%0], [%1, {%3, %4}], [%2];" ::"r"(smem_u32(smem_dst)),
        "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void tma_load_5d(void* smem_dst, const CUtensorMap* m, uint64_t* bar,
                                            int c0, int c1, int c2, int c3, int c4) {
    asm volatile(
        "cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes"
        " [%0], [%1, {%3, %4, %5, %6, %7}], [%2];" ::"r"(smem_u32(smem_dst)),
        "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2),
        "r"(c3), "r"(c4)
        : "memory");
}

// multicast variant: the box lands at the same shared-memory offset of every CTA in `mask` (and signals the
// mbarrier at the same offset in each of them)
__device__ __forceinline__ void tma_load_2d_mc(void* smem_dst, const CUtensorMap* m, uint64_t* bar,
                                               int c0, int c1, uint16_t mask) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster"
        " [%0], [%1, {%4, %5}], [%2], %3;" ::"r"(smem_u32(smem_dst)),
        "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "h"(mask), "r"(c0), "r"(c1)
        : "memory");
}

// ----------------------------------------------------------------------------
// thread-block cluster helpers
// ----------------------------------------------------------------------------
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}

// ----------------------------------------------------------------------------
// tcgen05: TMEM allocation, MMA, commit, TMEM loads, fences
// ----------------------------------------------------------------------------
__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                     smem_u32(dst_smem)),
                 "r"(ncols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols)
                 : "memory");
}
__device__ __forceinline__ void tc_fence_before() {
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after() {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
// D[tmem] (+)= A[smem] * B[smem]^T, bf16 inputs, fp32 accumulate, one CTA.
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b,
                                          uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
        "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
// mbarrier arrives once every tcgen05 op issued so far by this thread retires.
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile(
        "tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
            smem_u32(bar))
        : "memory");
}
// same, arriving on the mbarrier at this offset in every CTA of `mask` (stage release under TMA multicast)
__device__ __forceinline__ void umma_commit_mc(uint64_t* bar, uint16_t mask) {
    asm volatile(
        "tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
            smem_u32(bar)),
        "h"(mask)
        : "memory");
}
// 32 lanes x 16 consecutive fp32 columns: thread t of the warp gets lane
// (taddr.lane + t), columns taddr.col .. +15.
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t* v) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]),
          "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]),
          "=r"(v[14]), "=r"(v[15])
        : "r"(taddr)
        : "memory");
}
// Wait for outstanding tcgen05.ld; the loaded registers are tied to the asm ("+r") so the
// compiler cannot schedule their uses above the wait.
__device__ __forceinline__ void tmem_ld_wait16(uint32_t* v) {
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+r"(v[0]), "+r"(v[1]), "+r"(v[2]), "+r"(v[3]), "+r"(v[4]), "+r"(v[5]), "+r"(v[6]),
                   "+r"(v[7]), "+r"(v[8]), "+r"(v[9]), "+r"(v[10]), "+r"(v[11]), "+r"(v[12]), "+r"(v[13]),
                   "+r"(v[14]), "+r"(v[15])
                 :
                 : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() {
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// ----------------------------------------------------------------------------
// CTA-pair (cta_group::2) variants: one MMA spans two SMs (M = 256: each CTA's 128 rows of A and its own TMEM
// half; the N rows of B split between the two CTAs' shared memories).  Only the even ("leader") CTA issues.
// ----------------------------------------------------------------------------
__device__ __forceinline__ uint32_t mapa_shared(uint32_t cta_addr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(cta_addr), "r"(rank));
    return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void tmem_alloc_2sm(uint32_t* dst_smem, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)),
                 "r"(ncols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_2sm(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma_bf16_2sm(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                              uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
        "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
// arrives (once) on the mbarrier at this offset in every CTA of `mask` when the pair's MMAs issued so far retire
__device__ __forceinline__ void umma_commit_2sm_mc(uint64_t* bar, uint16_t mask) {
    asm volatile(
        "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
            smem_u32(bar)),
        "h"(mask)
        : "memory");
}
// TMA loads of a CTA pair: data lands in THIS CTA's shared memory, the bytes are counted on the mbarrier at
// cluster address `bar_cluster` (the leader's)
__device__ __forceinline__ void tma_load_2d_2sm(void* smem_dst, const CUtensorMap* m, uint32_t bar_cluster, int c0,
                                                int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes"
        " [%0], [%1, {%3, %4}], [%2];" ::"r"(smem_u32(smem_dst)),
        "l"(reinterpret_cast<uint64_t>(m)), "r"(bar_cluster), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void tma_load_5d_2sm(void* smem_dst, const CUtensorMap* m, uint32_t bar_cluster, int c0,
                                                int c1, int c2, int c3, int c4) {
    asm volatile(
        "cp.async.bulk.tensor.5d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes"
        " [%0], [%1, {%3, %4, %5, %6, %7}], [%2];" ::"r"(smem_u32(smem_dst)),
        "l"(reinterpret_cast<uint64_t>(m)), "r"(bar_cluster), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
        : "memory");
}

// K-major, 128-byte-swizzled operand tile: rows of 64 bf16 (128 B), 8-row
// swizzle atoms 1024 B apart.  Field layout as in the tcgen05 shared-memory
// matrix descriptor: start>>4 [0,14), LBO>>4 [16,30), SBO>>4 [32,46),
// version=1 [46,48), layout=SWIZZLE_128B(2) [61,64).
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t smem_addr) {
    const uint64_t hi = (uint64_t)(1024u >> 4) | (1ull << 14) | (2ull << 29);
    return (hi << 32) | (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
}
// kind::f16 instruction descriptor: D=f32, A=B=bf16, both K-major, M x N tile.
__host__ __device__ __forceinline__ uint32_t umma_idesc_bf16(int m, int n) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) |
           ((uint32_t)(m >> 4) << 24);
}
// same with the operand formats chosen per operand (a_format [7,10), b_format [10,13): 0 = fp16, 1 = bf16)
__host__ __device__ __forceinline__ uint32_t umma_idesc_16(int m, int n, int a_fp16, int b_fp16) {
    return (1u << 4) | ((a_fp16 ? 0u : 1u) << 7) | ((b_fp16 ? 0u : 1u) << 10) | ((uint32_t)(n >> 3) << 17) |
           ((uint32_t)(m >> 4) << 24);
}

}  // namespace v2a

// host: launch `kernel` with the PDL attribute (V2A_PDL=0 -> ordinary launch, A/B probe)
#ifdef __CUDACC__
namespace v2a {
static inline bool pdl_enabled() {
    static const bool on = [] { const char* e = getenv("V2A_PDL"); return !(e && atoi(e) == 0); }();
    return on;
}
}  // namespace v2a
template <typename... KArgs, typename... Args>
static inline cudaError_t launch_maybe_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st,
                                           Args&&... args) {
    const bool pdl = v2a::pdl_enabled();
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr;
    attr.id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr.val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = &attr;
    cfg.numAttrs = pdl ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...);
}
#endif
