// Per-frame spatial self-attention with the reference's "legacy" head layout
// (guided_diffusion/unet.py:341-358): qkv channels are grouped
// [head][q|k|v][32]; q and k are each scaled by 32^-1/4; softmax in fp32.
// 0.24 % of the UNet's FLOPs, L <= 256 and d = 32: K and V of one (frame, head)
// sit in shared memory, one query per thread, fp32 CUDA-core math with an
// online softmax; the result is written as bf16 hi/lo planes, the operand
// format of the proj_out tensor-core GEMM that follows.
#include "common.cuh"
#include "../../include/v2a_b200.h"

#include <atomic>
#include <cstdlib>

namespace v2a {
extern std::atomic<int64_t> g_launches;

constexpr int kHeadDim = 32;

// grid (N*heads, ceil(L / (blockDim * QPT))); smem: K[Lp][32], V[Lp][32] (Lp = L rounded up to 8, zero padded)
//
// QPT queries per thread share every K / V shared-memory read, keys are visited 8 at a time with ONE
// running-max rescale per block of 8 (the first version rescaled the 32 accumulators for every key and read
// K and V once per query: ~116 instructions per query-key, 0.83 ms per L=256 launch; this one ~78).
// exp is exp2 with log2(e) folded into the query scale.
template <int QPT>
__global__ void __launch_bounds__(128) attention_kernel(const float* __restrict__ qkv, int L, int heads,
                                                        __nv_bfloat16* __restrict__ out_hi,
                                                        __nv_bfloat16* __restrict__ out_lo) {
    extern __shared__ float4 kv_smem[];
    const int Lp = (L + 7) & ~7;
    float4* Ks = kv_smem;                 // [Lp][8] float4
    float4* Vs = kv_smem + (size_t)Lp * 8;
    const int n = blockIdx.x / heads, head = blockIdx.x % heads;
    const int C = heads * kHeadDim;
    const int ld = 3 * C;
    // (q * 32^-1/4) . (k * 32^-1/4) = (q . k) * 32^-1/2; softmax through exp2 -> one more factor log2(e)
    const float qscale = 0.17677669529663687f * 1.4426950408889634f;
    const float* base = qkv + (int64_t)n * L * ld + head * 3 * kHeadDim;
    for (int i = threadIdx.x; i < Lp * 8; i += blockDim.x) {
        const int s = i >> 3, part = i & 7;
        float4 k = make_float4(0.f, 0.f, 0.f, 0.f), v = k;
        if (s < L) {
            k = __ldg(reinterpret_cast<const float4*>(base + (int64_t)s * ld + kHeadDim) + part);
            v = __ldg(reinterpret_cast<const float4*>(base + (int64_t)s * ld + 2 * kHeadDim) + part);
        }
        Ks[i] = k;
        Vs[i] = v;
    }
    __syncthreads();
    const int t0 = blockIdx.y * blockDim.x * QPT + threadIdx.x;
    if (t0 >= L) return;
    float q[QPT][kHeadDim], acc[QPT][kHeadDim], m[QPT], l[QPT];
    bool live[QPT];
#pragma unroll
    for (int u = 0; u < QPT; ++u) {
        const int t = t0 + u * blockDim.x;
        live[u] = t < L;
        const float4* qp = reinterpret_cast<const float4*>(base + (int64_t)(live[u] ? t : t0) * ld);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const float4 v = __ldg(qp + j);
            q[u][4 * j] = v.x * qscale; q[u][4 * j + 1] = v.y * qscale;
            q[u][4 * j + 2] = v.z * qscale; q[u][4 * j + 3] = v.w * qscale;
        }
        m[u] = -INFINITY;
        l[u] = 0.0f;
#pragma unroll
        for (int j = 0; j < kHeadDim; ++j) acc[u][j] = 0.0f;
    }
    for (int s0 = 0; s0 < Lp; s0 += 8) {
        float d[QPT][8];
#pragma unroll
        for (int kk = 0; kk < 8; ++kk) {
#pragma unroll
            for (int u = 0; u < QPT; ++u) d[u][kk] = 0.0f;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const float4 k = Ks[(s0 + kk) * 8 + j];
#pragma unroll
                for (int u = 0; u < QPT; ++u)
                    d[u][kk] += q[u][4 * j] * k.x + q[u][4 * j + 1] * k.y + q[u][4 * j + 2] * k.z + q[u][4 * j + 3] * k.w;
            }
            if (s0 + kk >= L) {
#pragma unroll
                for (int u = 0; u < QPT; ++u) d[u][kk] = -INFINITY;
            }
        }
#pragma unroll
        for (int u = 0; u < QPT; ++u) {
            float bm = d[u][0];
#pragma unroll
            for (int kk = 1; kk < 8; ++kk) bm = fmaxf(bm, d[u][kk]);
            const float mn = fmaxf(m[u], bm);
            const float corr = exp2f(m[u] - mn);   // first block: exp2(-inf) = 0
            m[u] = mn;
            l[u] *= corr;
#pragma unroll
            for (int j = 0; j < kHeadDim; ++j) acc[u][j] *= corr;
#pragma unroll
            for (int kk = 0; kk < 8; ++kk) {
                d[u][kk] = exp2f(d[u][kk] - mn);
                l[u] += d[u][kk];
            }
        }
#pragma unroll
        for (int kk = 0; kk < 8; ++kk) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const float4 v = Vs[(s0 + kk) * 8 + j];
#pragma unroll
                for (int u = 0; u < QPT; ++u) {
                    acc[u][4 * j] += d[u][kk] * v.x;
                    acc[u][4 * j + 1] += d[u][kk] * v.y;
                    acc[u][4 * j + 2] += d[u][kk] * v.z;
                    acc[u][4 * j + 3] += d[u][kk] * v.w;
                }
            }
        }
    }
#pragma unroll
    for (int u = 0; u < QPT; ++u) {
        if (!live[u]) continue;
        const int t = t0 + u * blockDim.x;
        const float inv = 1.0f / l[u];
#pragma unroll
        for (int j = 0; j < kHeadDim; ++j) acc[u][j] *= inv;
        const int64_t o = ((int64_t)n * L + t) * C + head * kHeadDim;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            uint4 h, lo;
            split8(acc[u] + 8 * j, h, lo);
            *reinterpret_cast<uint4*>(out_hi + o + 8 * j) = h;
            *reinterpret_cast<uint4*>(out_lo + o + 8 * j) = lo;
        }
    }
}

// ---------------------------------------------------------------------------
// Tensor-core version (round 2) for L a multiple of 64 (the Libero UNet: L = 256 at 16x16, L = 64 at 8x8): one CTA per
// (frame, head), one warp per 16 queries, FlashAttention-2 dataflow on warp-level `mma.sync.m16n8k16` bf16 MMAs with
// the same 3-pass split product as every other contraction here (x ~ hi + lo; hi*hi + lo*hi + hi*lo, fp32 accumulate),
// so the result stays fp32-class.  Warp-level MMAs, not tcgen05: 0.24 % of the UNet's FLOPs in 16 x 32 x 256 problems
// per warp -- nothing here could fill a 128-row TMEM tile, while the CUDA-core version above spends ~78 instructions per
// query-key pair (3.5 ms per denoise step).
//   S = Q K^T   A = Q fragments (registers, pre-scaled by 32^-1/2 log2 e), B = K rows [key][d] in shared memory
//   P = exp2(S - max) per 64-key block with the usual running max / sum rescale
//   O += P V    A = P re-packed from the S accumulators (C-fragment layout == A-fragment layout), B = V^T [d][key]
// Shared-memory rows are padded (K: 40 halfs, V^T: L + 8) so the 32-bit fragment loads are bank-conflict free.
// ---------------------------------------------------------------------------
__device__ __forceinline__ void mma_bf16_16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile(
        "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, "
        "{%0, %1, %2, %3};"
        : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

constexpr int kKPad = 40;     // halfs per K row in shared memory (32 used)

__global__ void __launch_bounds__(512) attention_mma_kernel(const float* __restrict__ qkv, int L, int heads,
                                                            __nv_bfloat16* __restrict__ out_hi,
                                                            __nv_bfloat16* __restrict__ out_lo) {
    extern __shared__ __align__(16) unsigned char att_smem[];
    const int vpad = L + 8;
    __nv_bfloat16* Kh = reinterpret_cast<__nv_bfloat16*>(att_smem);          // [L][kKPad]
    __nv_bfloat16* Kl = Kh + (size_t)L * kKPad;
    __nv_bfloat16* Vh = Kl + (size_t)L * kKPad;                               // [32][vpad]  (V transposed)
    __nv_bfloat16* Vl = Vh + (size_t)kHeadDim * vpad;
    const int n = blockIdx.x / heads, head = blockIdx.x % heads;
    const int C = heads * kHeadDim;
    const int ld = 3 * C;
    const float* base = qkv + (int64_t)n * L * ld + head * 3 * kHeadDim;
    // ---- stage K (row major) and V (transposed) as bf16 (hi, lo) planes ----
    for (int i = threadIdx.x; i < L * 8; i += blockDim.x) {
        const int s = i >> 3, part = i & 7;
        const float4 k = __ldg(reinterpret_cast<const float4*>(base + (int64_t)s * ld + kHeadDim) + part);
        const float4 v = __ldg(reinterpret_cast<const float4*>(base + (int64_t)s * ld + 2 * kHeadDim) + part);
        uint32_t h0, l0, h1, l1;
        split2(k.x, k.y, h0, l0);
        split2(k.z, k.w, h1, l1);
        *reinterpret_cast<uint2*>(Kh + (size_t)s * kKPad + 4 * part) = make_uint2(h0, h1);
        *reinterpret_cast<uint2*>(Kl + (size_t)s * kKPad + 4 * part) = make_uint2(l0, l1);
        const float vv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            __nv_bfloat16 h, l;
            split_bf16(vv[j], h, l);
            Vh[(size_t)(4 * part + j) * vpad + s] = h;
            Vl[(size_t)(4 * part + j) * vpad + s] = l;
        }
    }
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int g = lane >> 2, t = lane & 3;
    const int q0 = warp * 16;                 // this warp's 16 queries
    if (q0 >= L) return;
    // (q * 32^-1/4) . (k * 32^-1/4) = (q . k) * 32^-1/2; softmax through exp2 -> one more factor log2(e)
    const float qscale = 0.17677669529663687f * 1.4426950408889634f;
    // ---- Q fragments: 2 k-slices of 16 channels, hi and lo ----
    uint32_t qh[2][4], ql[2][4];
#pragma unroll
    for (int ks = 0; ks < 2; ++ks) {
#pragma unroll
        for (int r = 0; r < 2; ++r) {         // rows g, g + 8
            const float* qp = base + (int64_t)(q0 + g + 8 * r) * ld + 16 * ks + 2 * t;
            const float2 a = __ldg(reinterpret_cast<const float2*>(qp));
            const float2 b = __ldg(reinterpret_cast<const float2*>(qp + 8));
            split2(a.x * qscale, a.y * qscale, qh[ks][r], ql[ks][r]);           // a0a1 (r = 0) / a2a3 (r = 1)
            split2(b.x * qscale, b.y * qscale, qh[ks][2 + r], ql[ks][2 + r]);   // a4a5 / a6a7
        }
    }
    float o[4][4];                            // 4 n-tiles of 8 channels
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
        for (int e = 0; e < 4; ++e) o[j][e] = 0.0f;
    float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.0f, l1 = 0.0f;      // rows g and g + 8
    for (int kb = 0; kb < L; kb += 64) {
        // ---- S = Q K^T for 64 keys: 8 n-tiles ----
        float sacc[8][4];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
#pragma unroll
            for (int e = 0; e < 4; ++e) sacc[j][e] = 0.0f;
            const __nv_bfloat16* kr_h = Kh + (size_t)(kb + 8 * j + g) * kKPad + 2 * t;
            const __nv_bfloat16* kr_l = Kl + (size_t)(kb + 8 * j + g) * kKPad + 2 * t;
#pragma unroll
            for (int ks = 0; ks < 2; ++ks) {
                const uint32_t bh0 = *reinterpret_cast<const uint32_t*>(kr_h + 16 * ks);
                const uint32_t bh1 = *reinterpret_cast<const uint32_t*>(kr_h + 16 * ks + 8);
                const uint32_t bl0 = *reinterpret_cast<const uint32_t*>(kr_l + 16 * ks);
                const uint32_t bl1 = *reinterpret_cast<const uint32_t*>(kr_l + 16 * ks + 8);
                mma_bf16_16816(sacc[j], ql[ks], bh0, bh1);     // small cross terms first
                mma_bf16_16816(sacc[j], qh[ks], bl0, bl1);
                mma_bf16_16816(sacc[j], qh[ks], bh0, bh1);
            }
        }
        // ---- online softmax over this block (rows g: elements 0,1; rows g + 8: elements 2,3) ----
        float bm0 = sacc[0][0], bm1 = sacc[0][2];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            bm0 = fmaxf(bm0, fmaxf(sacc[j][0], sacc[j][1]));
            bm1 = fmaxf(bm1, fmaxf(sacc[j][2], sacc[j][3]));
        }
        bm0 = fmaxf(bm0, __shfl_xor_sync(0xffffffffu, bm0, 1));
        bm0 = fmaxf(bm0, __shfl_xor_sync(0xffffffffu, bm0, 2));
        bm1 = fmaxf(bm1, __shfl_xor_sync(0xffffffffu, bm1, 1));
        bm1 = fmaxf(bm1, __shfl_xor_sync(0xffffffffu, bm1, 2));
        const float mn0 = fmaxf(m0, bm0), mn1 = fmaxf(m1, bm1);
        const float c0 = exp2f(m0 - mn0), c1 = exp2f(m1 - mn1);      // first block: exp2(-inf) = 0
        m0 = mn0;
        m1 = mn1;
        l0 *= c0;
        l1 *= c1;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            o[j][0] *= c0; o[j][1] *= c0;
            o[j][2] *= c1; o[j][3] *= c1;
        }
        // ---- P = exp2(S - m) as A fragments (hi, lo): 4 k-slices of 16 keys ----
        uint32_t ph[4][4], pl[4][4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
#pragma unroll
            for (int hf = 0; hf < 2; ++hf) {          // n-tiles 2i and 2i + 1 -> a0..a3 / a4..a7
                const int j = 2 * i + hf;
                const float p00 = exp2f(sacc[j][0] - mn0), p01 = exp2f(sacc[j][1] - mn0);
                const float p10 = exp2f(sacc[j][2] - mn1), p11 = exp2f(sacc[j][3] - mn1);
                l0 += p00 + p01;
                l1 += p10 + p11;
                split2(p00, p01, ph[i][2 * hf], pl[i][2 * hf]);             // rows g
                split2(p10, p11, ph[i][2 * hf + 1], pl[i][2 * hf + 1]);     // rows g + 8
            }
        }
        // ---- O += P V: 4 k-slices (16 keys) x 4 n-tiles (8 channels) ----
#pragma unroll
        for (int i = 0; i < 4; ++i) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const __nv_bfloat16* vr_h = Vh + (size_t)(8 * j + g) * vpad + kb + 16 * i + 2 * t;
                const __nv_bfloat16* vr_l = Vl + (size_t)(8 * j + g) * vpad + kb + 16 * i + 2 * t;
                const uint32_t bh0 = *reinterpret_cast<const uint32_t*>(vr_h);
                const uint32_t bh1 = *reinterpret_cast<const uint32_t*>(vr_h + 8);
                const uint32_t bl0 = *reinterpret_cast<const uint32_t*>(vr_l);
                const uint32_t bl1 = *reinterpret_cast<const uint32_t*>(vr_l + 8);
                mma_bf16_16816(o[j], pl[i], bh0, bh1);
                mma_bf16_16816(o[j], ph[i], bl0, bl1);
                mma_bf16_16816(o[j], ph[i], bh0, bh1);
            }
        }
    }
    // ---- finish: row sums over the quad, normalise, write hi / lo planes ----
    l0 += __shfl_xor_sync(0xffffffffu, l0, 1);
    l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
    l1 += __shfl_xor_sync(0xffffffffu, l1, 1);
    l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
    const float i0 = 1.0f / l0, i1 = 1.0f / l1;
    const int64_t r0 = ((int64_t)n * L + q0 + g) * C + head * kHeadDim + 2 * t;
    const int64_t r1 = r0 + (int64_t)8 * C;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        uint32_t h, l;
        split2(o[j][0] * i0, o[j][1] * i0, h, l);
        *reinterpret_cast<uint32_t*>(out_hi + r0 + 8 * j) = h;
        *reinterpret_cast<uint32_t*>(out_lo + r0 + 8 * j) = l;
        split2(o[j][2] * i1, o[j][3] * i1, h, l);
        *reinterpret_cast<uint32_t*>(out_hi + r1 + 8 * j) = h;
        *reinterpret_cast<uint32_t*>(out_lo + r1 + 8 * j) = l;
    }
}

}  // namespace v2a

extern "C" int v2a_attention(const float* qkv, int N, int L, int heads, void* out_hi, void* out_lo,
                             void* stream) {
    using namespace v2a;
    V2A_REQUIRE(N >= 1 && L >= 1 && heads >= 1, "attention: bad shape");
    const int Lp = (L + 7) & ~7;
    const size_t smem = (size_t)Lp * kHeadDim * sizeof(float) * 2;
    V2A_REQUIRE(smem <= 200 * 1024, "attention: L %d too long for the shared-memory K/V tile", L);
    static bool attr_set = false;
    if (!attr_set) {
        V2A_CUDA_OK(cudaFuncSetAttribute(attention_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        V2A_CUDA_OK(cudaFuncSetAttribute(attention_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        attr_set = true;
    }
    const char* env = getenv("V2A_ATTN_MMA");
    if (L % 64 == 0 && L <= 256 && !(env && atoi(env) == 0)) {      // tensor-core version: one warp per 16 queries
        const size_t sm = ((size_t)2 * L * kKPad + (size_t)2 * kHeadDim * (L + 8)) * sizeof(__nv_bfloat16);
        static bool mma_attr = false;
        if (!mma_attr) {
            V2A_CUDA_OK(cudaFuncSetAttribute(attention_mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
            mma_attr = true;
        }
        attention_mma_kernel<<<(unsigned)(N * heads), (unsigned)(L / 16 * 32), sm, (cudaStream_t)stream>>>(
            qkv, L, heads, (__nv_bfloat16*)out_hi, (__nv_bfloat16*)out_lo);
    } else if (L >= 256) {      // 2 queries per thread: halves the shared-memory reads per FMA
        dim3 grid((unsigned)(N * heads), (unsigned)((L + 255) / 256));
        attention_kernel<2><<<grid, 128, smem, (cudaStream_t)stream>>>(qkv, L, heads, (__nv_bfloat16*)out_hi,
                                                                       (__nv_bfloat16*)out_lo);
    } else {
        const int threads = L >= 128 ? 128 : ((L + 31) / 32) * 32;
        dim3 grid((unsigned)(N * heads), (unsigned)((L + threads - 1) / threads));
        attention_kernel<1><<<grid, threads, smem, (cudaStream_t)stream>>>(qkv, L, heads, (__nv_bfloat16*)out_hi,
                                                                           (__nv_bfloat16*)out_lo);
    }
    V2A_CUDA_OK(cudaGetLastError());
    g_launches.fetch_add(1);
    return 0;
}
