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
    if (L >= 256) {      // 2 queries per thread: halves the shared-memory reads per FMA
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
