// Task-token conditioning of the video UNet: `task_attnpool` =
// PerceiverResampler(dim 512, depth 2, 64 latents + 4 mean-pooled latents) -> Linear(512, 512) -> mean over latents
//   guided_diffusion/unet.py:491-494,671; guided_diffusion/imagen.py:197-211 (gain-only LayerNorm),
//   :254-319 (PerceiverAttention: LayerNorms, l2-normalised q/k with learned scales, logits x 8, softmax),
//   :321-372 (PerceiverResampler), :1009-1017 (FeedForward: LN, Linear x4, GELU, LN, Linear)
// Step-invariant: evaluated once per sample() call on B x (L + 68) tokens (0.92 GFLOP against 213 TFLOP for the
// denoising loop), so these are plain fp32 CUDA-core kernels — small, exact-fp32 and launch-latency bound.  The
// dense layers in between run on v2a_linear.
#include "common.cuh"
#include "../../include/v2a_b200.h"

#include <atomic>

namespace v2a {
extern std::atomic<int64_t> g_launches;

#define V2A_PR_LAUNCH_OK()                   \
    do {                                     \
        V2A_CUDA_OK(cudaGetLastError());     \
        v2a::g_launches.fetch_add(1);        \
    } while (0)

__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752f)); }

__device__ __forceinline__ float block_sum_128(float v, float* red) {
    // 128 threads; every thread returns the total
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    __syncthreads();                       // `red` may still be read from the previous reduction
    if (lane == 0) red[warp] = v;
    __syncthreads();
    return red[0] + red[1] + red[2] + red[3];
}

// one block (128 threads) per token row.  Row r = (b, i) with i < n_tok:
//   in  = x   + b * x_batch   + i * ldx      (+ pos[i * D + c] when pos, after act)
//   out = out + b * out_batch + i * ld_out
// y = (act(in) - mean) * rsqrt(var + eps) * gamma (+ beta), biased variance, fp32 (F.layer_norm / imagen LayerNorm)
__global__ void __launch_bounds__(128) pr_layernorm_kernel(const float* __restrict__ x, int64_t x_batch, int ldx,
                                                           const float* __restrict__ pos, int n_tok, int D, int act,
                                                           const float* __restrict__ gamma,
                                                           const float* __restrict__ beta, float eps,
                                                           float* __restrict__ out, int64_t out_batch, int ld_out) {
    __shared__ float red[4];
    const int r = blockIdx.x;
    const int b = r / n_tok, i = r - b * n_tok;
    const float* in = x + (int64_t)b * x_batch + (int64_t)i * ldx;
    const float* p = pos ? pos + (int64_t)i * D : nullptr;
    float* o = out + (int64_t)b * out_batch + (int64_t)i * ld_out;
    float s = 0.f;
    for (int c = threadIdx.x; c < D; c += 128) {
        float v = in[c];
        if (act == 3) v = gelu_erf(v);
        if (p) v += p[c];
        s += v;
    }
    const float mean = block_sum_128(s, red) / (float)D;
    float q = 0.f;
    for (int c = threadIdx.x; c < D; c += 128) {
        float v = in[c];
        if (act == 3) v = gelu_erf(v);
        if (p) v += p[c];
        const float d = v - mean;
        q += d * d;
    }
    const float rstd = rsqrtf(block_sum_128(q, red) / (float)D + eps);
    for (int c = threadIdx.x; c < D; c += 128) {
        float v = in[c];
        if (act == 3) v = gelu_erf(v);
        if (p) v += p[c];
        float y = (v - mean) * rstd * gamma[c];
        if (beta) y += beta[c];
        o[c] = y;
    }
}

// one warp per (row, head): out = x / max(||x||_2, 1e-12) * scale   (F.normalize(dim=-1) * q_scale / k_scale)
__global__ void pr_l2norm_scale_kernel(const float* __restrict__ x, int ldx, int rows, int heads, int dh,
                                       const float* __restrict__ scale, float* __restrict__ out, int ld_out) {
    const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (w >= rows * heads) return;
    const int r = w / heads, h = w - r * heads;
    const float* in = x + (int64_t)r * ldx + h * dh;
    float* o = out + (int64_t)r * ld_out + h * dh;
    float s = 0.f;
    for (int d = lane; d < dh; d += 32) s += in[d] * in[d];
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
    const float inv = 1.0f / fmaxf(sqrtf(s), 1e-12f);
    for (int d = lane; d < dh; d += 32) o[d] = in[d] * inv * scale[d];
}

// one warp per (b, head, query): softmax_j(scale * q_i . k_j) @ v, online softmax in fp32.
// q [B * nq][ldq], k / v [B * nk][ldk / ldv], features of head h at columns h * dh .. (dh <= 128)
__global__ void pr_attention_kernel(const float* __restrict__ q, int ldq, const float* __restrict__ k, int ldk,
                                    const float* __restrict__ v, int ldv, int B, int heads, int dh, int nq, int nk,
                                    float scale, float* __restrict__ out, int ld_out) {
    const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (w >= B * heads * nq) return;
    const int i = w % nq;
    const int h = (w / nq) % heads;
    const int b = w / (nq * heads);
    const float* qi = q + (int64_t)(b * nq + i) * ldq + h * dh;
    float qr[4], acc[4];
#pragma unroll
    for (int t = 0; t < 4; ++t) {
        const int d = lane + 32 * t;
        qr[t] = d < dh ? qi[d] : 0.f;
        acc[t] = 0.f;
    }
    float m = -INFINITY, l = 0.f;
    for (int j = 0; j < nk; ++j) {
        const float* kj = k + (int64_t)(b * nk + j) * ldk + h * dh;
        const float* vj = v + (int64_t)(b * nk + j) * ldv + h * dh;
        float s = 0.f;
#pragma unroll
        for (int t = 0; t < 4; ++t) {
            const int d = lane + 32 * t;
            if (d < dh) s += qr[t] * kj[d];
        }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
        s *= scale;
        const float mn = fmaxf(m, s);
        const float corr = expf(m - mn);          // exp(-inf) = 0 on the first key
        const float pj = expf(s - mn);
        l = l * corr + pj;
#pragma unroll
        for (int t = 0; t < 4; ++t) {
            const int d = lane + 32 * t;
            if (d < dh) acc[t] = acc[t] * corr + pj * vj[d];
        }
        m = mn;
    }
    float* o = out + (int64_t)(b * nq + i) * ld_out + h * dh;
    const float invl = 1.0f / l;
#pragma unroll
    for (int t = 0; t < 4; ++t) {
        const int d = lane + 32 * t;
        if (d < dh) o[d] = acc[t] * invl;
    }
}

// out[b][c] = mean_i x[b][i][c]
__global__ void pr_token_mean_kernel(const float* __restrict__ x, int64_t x_batch, int ldx, int B, int n, int D,
                                     float* __restrict__ out, int ld_out) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= B * D) return;
    const int b = idx / D, c = idx - b * D;
    const float* in = x + (int64_t)b * x_batch + c;
    float s = 0.f;
    for (int i = 0; i < n; ++i) s += in[(int64_t)i * ldx];
    out[(int64_t)b * ld_out + c] = s / (float)n;
}

// out[b][i][c] = src[i][c]   (the learned latents broadcast over the batch)
__global__ void pr_broadcast_rows_kernel(const float* __restrict__ src, int n, int D, int B, float* __restrict__ out,
                                         int64_t out_batch, int ld_out) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= B * n * D) return;
    const int c = idx % D;
    const int i = (idx / D) % n;
    const int b = idx / (D * n);
    out[(int64_t)b * out_batch + (int64_t)i * ld_out + c] = src[(int64_t)i * D + c];
}

}  // namespace v2a

extern "C" {

int v2a_pr_layernorm(const float* x, int64_t x_batch, int ldx, const float* pos, int B, int n_tok, int D, int act,
                     const float* gamma, const float* beta, float eps, float* out, int64_t out_batch, int ld_out,
                     void* stream) {
    V2A_REQUIRE(x && gamma && out, "pr_layernorm: missing pointers");
    V2A_REQUIRE(B >= 1 && n_tok >= 1 && D >= 1 && (act == 0 || act == 3), "pr_layernorm: bad arguments");
    v2a::pr_layernorm_kernel<<<(unsigned)(B * n_tok), 128, 0, (cudaStream_t)stream>>>(
        x, x_batch, ldx, pos, n_tok, D, act, gamma, beta, eps, out, out_batch, ld_out);
    V2A_PR_LAUNCH_OK();
    return 0;
}

int v2a_pr_l2norm_scale(const float* x, int ldx, int rows, int heads, int dh, const float* scale, float* out,
                        int ld_out, void* stream) {
    V2A_REQUIRE(x && scale && out && rows >= 1 && heads >= 1 && dh >= 1, "pr_l2norm_scale: bad arguments");
    const int64_t threads = (int64_t)rows * heads * 32;
    v2a::pr_l2norm_scale_kernel<<<(unsigned)((threads + 127) / 128), 128, 0, (cudaStream_t)stream>>>(
        x, ldx, rows, heads, dh, scale, out, ld_out);
    V2A_PR_LAUNCH_OK();
    return 0;
}

int v2a_pr_attention(const float* q, int ldq, const float* k, int ldk, const float* v, int ldv, int B, int heads,
                     int dh, int nq, int nk, float scale, float* out, int ld_out, void* stream) {
    V2A_REQUIRE(q && k && v && out, "pr_attention: missing pointers");
    V2A_REQUIRE(B >= 1 && heads >= 1 && dh >= 1 && dh <= 128 && nq >= 1 && nk >= 1, "pr_attention: bad shape");
    const int64_t threads = (int64_t)B * heads * nq * 32;
    v2a::pr_attention_kernel<<<(unsigned)((threads + 127) / 128), 128, 0, (cudaStream_t)stream>>>(
        q, ldq, k, ldk, v, ldv, B, heads, dh, nq, nk, scale, out, ld_out);
    V2A_PR_LAUNCH_OK();
    return 0;
}

int v2a_pr_token_mean(const float* x, int64_t x_batch, int ldx, int B, int n, int D, float* out, int ld_out,
                      void* stream) {
    V2A_REQUIRE(x && out && B >= 1 && n >= 1 && D >= 1, "pr_token_mean: bad arguments");
    v2a::pr_token_mean_kernel<<<(unsigned)v2a::ceil_div(B * D, 256), 256, 0, (cudaStream_t)stream>>>(
        x, x_batch, ldx, B, n, D, out, ld_out);
    V2A_PR_LAUNCH_OK();
    return 0;
}

int v2a_pr_broadcast_rows(const float* src, int n, int D, int B, float* out, int64_t out_batch, int ld_out,
                          void* stream) {
    V2A_REQUIRE(src && out && B >= 1 && n >= 1 && D >= 1, "pr_broadcast_rows: bad arguments");
    v2a::pr_broadcast_rows_kernel<<<(unsigned)v2a::ceil_div(B * n * D, 256), 256, 0, (cudaStream_t)stream>>>(
        src, n, D, B, out, out_batch, ld_out);
    V2A_PR_LAUNCH_OK();
    return 0;
}

}  // extern "C"
