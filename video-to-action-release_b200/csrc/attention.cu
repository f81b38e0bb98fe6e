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

// grid (N*heads, ceil(L/blockDim)); smem: K[L][32], V[L][32]
__global__ void __launch_bounds__(256) attention_kernel(const float* __restrict__ qkv, int L, int heads,
                                                        __nv_bfloat16* __restrict__ out_hi,
                                                        __nv_bfloat16* __restrict__ out_lo) {
    extern __shared__ float4 kv_smem[];
    float4* Ks = kv_smem;                 // [L][8] float4
    float4* Vs = kv_smem + (size_t)L * 8;
    const int n = blockIdx.x / heads, head = blockIdx.x % heads;
    const int C = heads * kHeadDim;
    const int ld = 3 * C;
    const float scale = 0.42044820762685725f;  // 32^-0.25
    const float* base = qkv + (int64_t)n * L * ld + head * 3 * kHeadDim;
    for (int i = threadIdx.x; i < L * 8; i += blockDim.x) {
        const int s = i >> 3, part = i & 7;
        float4 k = __ldg(reinterpret_cast<const float4*>(base + (int64_t)s * ld + kHeadDim) + part);
        k.x *= scale; k.y *= scale; k.z *= scale; k.w *= scale;
        Ks[i] = k;
        Vs[i] = __ldg(reinterpret_cast<const float4*>(base + (int64_t)s * ld + 2 * kHeadDim) + part);
    }
    __syncthreads();
    const int t = blockIdx.y * blockDim.x + threadIdx.x;
    if (t >= L) return;
    float q[kHeadDim];
    {
        const float4* qp = reinterpret_cast<const float4*>(base + (int64_t)t * ld);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const float4 v = __ldg(qp + j);
            q[4 * j] = v.x * scale; q[4 * j + 1] = v.y * scale;
            q[4 * j + 2] = v.z * scale; q[4 * j + 3] = v.w * scale;
        }
    }
    float m = -INFINITY, l = 0.0f, acc[kHeadDim];
#pragma unroll
    for (int j = 0; j < kHeadDim; ++j) acc[j] = 0.0f;
    for (int s = 0; s < L; ++s) {
        float dot = 0.0f;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const float4 k = Ks[s * 8 + j];
            dot += q[4 * j] * k.x + q[4 * j + 1] * k.y + q[4 * j + 2] * k.z + q[4 * j + 3] * k.w;
        }
        const float mn = fmaxf(m, dot);
        const float corr = expf(m - mn);   // first iteration: exp(-inf) = 0
        const float pe = expf(dot - mn);
        l = l * corr + pe;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const float4 v = Vs[s * 8 + j];
            acc[4 * j] = acc[4 * j] * corr + pe * v.x;
            acc[4 * j + 1] = acc[4 * j + 1] * corr + pe * v.y;
            acc[4 * j + 2] = acc[4 * j + 2] * corr + pe * v.z;
            acc[4 * j + 3] = acc[4 * j + 3] * corr + pe * v.w;
        }
        m = mn;
    }
    const float inv = 1.0f / l;
#pragma unroll
    for (int j = 0; j < kHeadDim; ++j) acc[j] *= inv;
    const int64_t o = ((int64_t)n * L + t) * C + head * kHeadDim;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        uint4 h, lo;
        split8(acc + 8 * j, h, lo);
        *reinterpret_cast<uint4*>(out_hi + o + 8 * j) = h;
        *reinterpret_cast<uint4*>(out_lo + o + 8 * j) = lo;
    }
}

}  // namespace v2a

extern "C" int v2a_attention(const float* qkv, int N, int L, int heads, void* out_hi, void* out_lo,
                             void* stream) {
    using namespace v2a;
    V2A_REQUIRE(N >= 1 && L >= 1 && heads >= 1, "attention: bad shape");
    const size_t smem = (size_t)L * kHeadDim * sizeof(float) * 2;
    V2A_REQUIRE(smem <= 200 * 1024, "attention: L %d too long for the shared-memory K/V tile", L);
    static bool attr_set = false;
    if (!attr_set) {
        V2A_CUDA_OK(cudaFuncSetAttribute(attention_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         200 * 1024));
        attr_set = true;
    }
    int threads = L >= 256 ? 256 : ((L + 31) / 32) * 32;
    dim3 grid((unsigned)(N * heads), (unsigned)((L + threads - 1) / threads));
    attention_kernel<<<grid, threads, smem, (cudaStream_t)stream>>>(
        qkv, L, heads, (__nv_bfloat16*)out_hi, (__nv_bfloat16*)out_lo);
    V2A_CUDA_OK(cudaGetLastError());
    g_launches.fetch_add(1);
    return 0;
}
