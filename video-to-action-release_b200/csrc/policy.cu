// Policy-path kernels around the tensor-core contractions of ConditionalUnet1D
// (diffuser/diffusion_policy/model/conditional_unet1d.py, conv1d_components.py).
//
// One sample of the policy UNet is tiny (T*C = 16*256 = 8*512 = 4*1024 = 4096 values per
// conv output), so GroupNorm + Mish + FiLM forward AND backward each run as ONE CTA per
// batch sample with the whole sample in registers: exact two-pass statistics, no atomics
// on the activation path, and the backward emits everything the conv gradients need in
// one pass (dy as bf16 hi/lo planes for the data-gradient GEMM, dy^T for the
// weight-gradient GEMM, bias / gamma / beta / FiLM gradients).
#include "common.cuh"
#include "../../include/v2a_b200.h"

#include <atomic>

namespace v2a {
extern std::atomic<int64_t> g_launches;

constexpr int kPolThreads = 256;
constexpr int kMaxOct = 8;  // 8-channel octets per thread (T*C <= 256*8*8 = 16384)

struct GnActParams {
    const float* y;        // [B][T][C] conv output
    int T, C, groups;
    float eps;
    const float* gamma; const float* beta;
    const float* film; int ld_film;       // [B][2C] scale|bias or null
    const float* addend; int ld_add;      // [B][T][ld_add] identity residual or null
    float* out_f32; int ld_out;           // [B][T][ld_out] or null
    __nv_bfloat16* out_hi; __nv_bfloat16* out_lo; int ld_hl;
    float* mean_rstd;                     // [B][groups][2]
    // backward only
    const float* dout; int ld_dout;       // [B][T][ld_dout]
    __nv_bfloat16* dy_hi; __nv_bfloat16* dy_lo;     // [B][T][C]
    __nv_bfloat16* dyT_hi; __nv_bfloat16* dyT_lo;   // [C][ld_T] (ld_T >= B*T, zero padded)
    long long ld_T;
    float* dy_f32;                        // optional [B][T][C]
    float* dbias; float* dgamma; float* dbeta;      // [C] accumulated with atomics over samples
    float* dfilm; int ld_dfilm;           // [B][2C] written
    int B;
    int stage_words;                      // bwd: T*C when dy^T is staged through shared memory, else 0
    float* partials;                      // bwd: [B][3][C] per-sample bias/gamma/beta sums (no global atomics) or null
};

// thread -> (octet of 8 channels, lane over t); values of a thread share one group
__device__ __forceinline__ void thread_map(int C, int& octs, int& tl, int& oc, int& tlane) {
    octs = C >> 3;
    tl = kPolThreads / octs;
    if (tl < 1) tl = 1;
    oc = threadIdx.x % octs;
    tlane = threadIdx.x / octs;
}

__device__ __forceinline__ void load8(const float* p, float* v) {
    const float4 a = __ldg(reinterpret_cast<const float4*>(p));
    const float4 b = __ldg(reinterpret_cast<const float4*>(p) + 1);
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
__device__ __forceinline__ void store8(float* p, const float* v) {
    reinterpret_cast<float4*>(p)[0] = make_float4(v[0], v[1], v[2], v[3]);
    reinterpret_cast<float4*>(p)[1] = make_float4(v[4], v[5], v[6], v[7]);
}

// block-wide per-group reduction of one float per thread (thread's group = g); result broadcast
__device__ __forceinline__ float group_sum(float v, int g, float* sh, int groups) {
    __syncthreads();
    if (threadIdx.x < groups) sh[threadIdx.x] = 0.0f;
    __syncthreads();
    atomicAdd(&sh[g], v);
    __syncthreads();
    return sh[g];
}

__global__ void __launch_bounds__(kPolThreads) gn_act_fwd_kernel(const GnActParams p) {
    __shared__ float sh[64];
    const int b = blockIdx.x;
    int octs, tl, oc, tlane;
    thread_map(p.C, octs, tl, oc, tlane);
    const bool active = tlane < tl && oc < octs;
    const int c0 = oc * 8;
    const int cpg = p.C / p.groups;
    const int g = active ? c0 / cpg : 0;
    float v[kMaxOct][8];
    int nt = 0;
    float s = 0.0f;
    if (active)
        for (int t = tlane; t < p.T; t += tl, ++nt) {
            load8(p.y + ((int64_t)b * p.T + t) * p.C + c0, v[nt]);
#pragma unroll
            for (int j = 0; j < 8; ++j) s += v[nt][j];
        }
    const float cnt = (float)(cpg * p.T);
    const float mean = group_sum(s, g, sh, p.groups) / cnt;
    float q = 0.0f;
    for (int i = 0; i < nt; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) { const float d = v[i][j] - mean; q += d * d; }
    const float var = group_sum(q, g, sh, p.groups) / cnt;
    const float rstd = rsqrtf(var + p.eps);
    if (!active) return;
    if (tlane == 0 && (c0 % cpg) == 0) {
        p.mean_rstd[((int64_t)b * p.groups + g) * 2] = mean;
        p.mean_rstd[((int64_t)b * p.groups + g) * 2 + 1] = rstd;
    }
    float ga[8], be[8], fs[8], fb[8];
    load8(p.gamma + c0, ga);
    load8(p.beta + c0, be);
    if (p.film) {
        load8(p.film + (int64_t)b * p.ld_film + c0, fs);
        load8(p.film + (int64_t)b * p.ld_film + p.C + c0, fb);
    }
    int i = 0;
    for (int t = tlane; t < p.T; t += tl, ++i) {
        float o[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            float x = mish_f((v[i][j] - mean) * rstd * ga[j] + be[j]);
            if (p.film) x = fs[j] * x + fb[j];
            o[j] = x;
        }
        const int64_t row = (int64_t)b * p.T + t;
        if (p.addend) {
            float a[8];
            load8(p.addend + row * p.ld_add + c0, a);
#pragma unroll
            for (int j = 0; j < 8; ++j) o[j] += a[j];
        }
        if (p.out_f32) store8(p.out_f32 + row * p.ld_out + c0, o);
        if (p.out_hi) {
            uint4 h, l;
            split8(o, h, l);
            *reinterpret_cast<uint4*>(p.out_hi + row * p.ld_hl + c0) = h;
            *reinterpret_cast<uint4*>(p.out_lo + row * p.ld_hl + c0) = l;
        }
    }
}

// Small-batch form of the forward (`predict_action`: B = 1 .. 4 between simulator steps): one CTA per (sample, GROUP)
// instead of one per sample -- at B = 1 the per-sample kernel is a single block walking 16 k values through two
// shared-memory atomic reductions (~9 us, 200 launches per predict_action call).  Here a block holds T x C/groups
// values (<= 2 octets per thread), every parameter load is issued before the statistics, and the two reductions are
// warp shuffles + one barrier each.
constexpr int kGrpOct = 2;
__global__ void __launch_bounds__(kPolThreads) gn_act_fwd_group_kernel(const GnActParams p) {
    __shared__ float red[2][kPolThreads / 32];
    const int b = blockIdx.x, g = blockIdx.y;
    const int cpg = p.C / p.groups;
    const int octs = cpg >> 3;                    // octets per row inside the group; divides the block size
    const int n_oct = p.T * octs;
    const int oc = threadIdx.x % octs;
    const int c0 = g * cpg + oc * 8;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
    float v[kGrpOct][8], ga[8], be[8], fs[8], fb[8], ad[kGrpOct][8];
    bool on[kGrpOct];
    pdl_trigger();
    load8(p.gamma + c0, ga);       // parameters: not written by any kernel of the chain
    load8(p.beta + c0, be);
    pdl_wait();
#pragma unroll
    for (int k = 0; k < kGrpOct; ++k) {
        const int i = threadIdx.x + k * blockDim.x;
        on[k] = i < n_oct;
        const int64_t row = (int64_t)b * p.T + i / octs;
        if (on[k]) {
            load8(p.y + row * p.C + c0, v[k]);
            if (p.addend) load8(p.addend + row * p.ld_add + c0, ad[k]);
        } else {
#pragma unroll
            for (int j = 0; j < 8; ++j) v[k][j] = 0.0f;
        }
    }
    if (p.film) {
        load8(p.film + (int64_t)b * p.ld_film + c0, fs);
        load8(p.film + (int64_t)b * p.ld_film + p.C + c0, fb);
    }
    auto block_sum = [&](float x, float* sh) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
        if (lane == 0) sh[warp] = x;
        __syncthreads();
        float t = 0.0f;
        for (int w = 0; w < nw; ++w) t += sh[w];
        return t;
    };
    float s = 0.0f;
#pragma unroll
    for (int k = 0; k < kGrpOct; ++k)
#pragma unroll
        for (int j = 0; j < 8; ++j) s += v[k][j];
    const float cnt = (float)(cpg * p.T);
    const float mean = block_sum(s, red[0]) / cnt;
    float q = 0.0f;
#pragma unroll
    for (int k = 0; k < kGrpOct; ++k)
        if (on[k]) {
#pragma unroll
            for (int j = 0; j < 8; ++j) { const float d = v[k][j] - mean; q += d * d; }
        }
    const float rstd = rsqrtf(block_sum(q, red[1]) / cnt + p.eps);
    if (threadIdx.x == 0) {
        p.mean_rstd[((int64_t)b * p.groups + g) * 2] = mean;
        p.mean_rstd[((int64_t)b * p.groups + g) * 2 + 1] = rstd;
    }
#pragma unroll
    for (int k = 0; k < kGrpOct; ++k) {
        if (!on[k]) continue;
        const int i = threadIdx.x + k * blockDim.x;
        const int64_t row = (int64_t)b * p.T + i / octs;
        float o[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            float x = mish_f((v[k][j] - mean) * rstd * ga[j] + be[j]);
            if (p.film) x = fs[j] * x + fb[j];
            if (p.addend) x += ad[k][j];
            o[j] = x;
        }
        if (p.out_f32) store8(p.out_f32 + row * p.ld_out + c0, o);
        if (p.out_hi) {
            uint4 h, l;
            split8(o, h, l);
            *reinterpret_cast<uint4*>(p.out_hi + row * p.ld_hl + c0) = h;
            *reinterpret_cast<uint4*>(p.out_lo + row * p.ld_hl + c0) = l;
        }
    }
}

// backward of  out = film_s * mish(gamma * (y - mean) * rstd + beta) + film_b  (+ addend)
__global__ void __launch_bounds__(kPolThreads) gn_act_bwd_kernel(const GnActParams p) {
    __shared__ float sh[64];
    const int b = blockIdx.x;
    int octs, tl, oc, tlane;
    thread_map(p.C, octs, tl, oc, tlane);
    const bool active = tlane < tl && oc < octs;
    const int c0 = oc * 8;
    const int cpg = p.C / p.groups;
    const int g = active ? c0 / cpg : 0;
    const float mean = p.mean_rstd[((int64_t)b * p.groups + g) * 2];
    const float rstd = p.mean_rstd[((int64_t)b * p.groups + g) * 2 + 1];
    float ga[8], be[8], fs[8];
    if (active) {
        load8(p.gamma + c0, ga);
        load8(p.beta + c0, be);
        if (p.film) load8(p.film + (int64_t)b * p.ld_film + c0, fs);
    }
    float n[kMaxOct][8], dn[kMaxOct][8];  // normalised input, grad wrt normalised input
    float dsc[8] = {0, 0, 0, 0, 0, 0, 0, 0}, dsh[8] = {0, 0, 0, 0, 0, 0, 0, 0};   // FiLM grads
    float dga[8] = {0, 0, 0, 0, 0, 0, 0, 0}, dbe[8] = {0, 0, 0, 0, 0, 0, 0, 0};   // gamma / beta grads
    float s1 = 0.0f, s2 = 0.0f;
    int nt = 0;
    if (active)
        for (int t = tlane; t < p.T; t += tl, ++nt) {
            const int64_t row = (int64_t)b * p.T + t;
            float y[8], d[8];
            load8(p.y + row * p.C + c0, y);
            load8(p.dout + row * p.ld_dout + c0, d);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const float nn = (y[j] - mean) * rstd;
                const float gg = nn * ga[j] + be[j];
                float dm = d[j];
                if (p.film) {
                    dsc[j] += d[j] * mish_f(gg);
                    dsh[j] += d[j];
                    dm = d[j] * fs[j];
                }
                const float dg = dm * mish_grad_f(gg);
                dga[j] += dg * nn;
                dbe[j] += dg;
                const float dnn = dg * ga[j];
                n[nt][j] = nn;
                dn[nt][j] = dnn;
                s1 += dnn;
                s2 += dnn * nn;
            }
        }
    const float cnt = (float)(cpg * p.T);
    const float m1 = group_sum(s1, g, sh, p.groups) / cnt;
    const float m2 = group_sum(s2, g, sh, p.groups) / cnt;
    // dy^T staging: (hi, lo) pairs of the whole sample, [T][C] 32-bit words, when it fits (policy sizes: 16 KB)
    extern __shared__ uint32_t stage[];
    const bool staged = p.dyT_hi != nullptr && p.stage_words > 0;
    float dbi[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    int i = 0;
    const int64_t BT = p.ld_T;
    if (active)
        for (int t = tlane; t < p.T; t += tl, ++i) {
            const int64_t row = (int64_t)b * p.T + t;
            float dy[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                dy[j] = rstd * (dn[i][j] - m1 - n[i][j] * m2);
                dbi[j] += dy[j];
            }
            if (p.dy_f32) store8(p.dy_f32 + row * p.C + c0, dy);
            uint4 h, l;
            split8(dy, h, l);
            if (p.dy_hi) {
                *reinterpret_cast<uint4*>(p.dy_hi + row * p.C + c0) = h;
                *reinterpret_cast<uint4*>(p.dy_lo + row * p.C + c0) = l;
            }
            if (p.dyT_hi) {
                const uint16_t* hh = reinterpret_cast<const uint16_t*>(&h);
                const uint16_t* ll = reinterpret_cast<const uint16_t*>(&l);
                if (staged) {
#pragma unroll
                    for (int j = 0; j < 8; ++j)
                        stage[t * p.C + c0 + j] = (uint32_t)hh[j] | ((uint32_t)ll[j] << 16);
                } else {
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        reinterpret_cast<uint16_t*>(p.dyT_hi)[(int64_t)(c0 + j) * BT + row] = hh[j];
                        reinterpret_cast<uint16_t*>(p.dyT_lo)[(int64_t)(c0 + j) * BT + row] = ll[j];
                    }
                }
            }
        }
    if (staged) {
        __syncthreads();
        // one channel per thread: its T values are one contiguous run of dy^T (4 bf16 = 8 B per store)
        for (int c = threadIdx.x; c < p.C; c += blockDim.x) {
            uint16_t* dh = reinterpret_cast<uint16_t*>(p.dyT_hi) + (int64_t)c * BT + (int64_t)b * p.T;
            uint16_t* dl = reinterpret_cast<uint16_t*>(p.dyT_lo) + (int64_t)c * BT + (int64_t)b * p.T;
            for (int t = 0; t < p.T; t += 4) {
                const uint32_t w0 = stage[t * p.C + c], w1 = stage[(t + 1) * p.C + c];
                const uint32_t w2 = stage[(t + 2) * p.C + c], w3 = stage[(t + 3) * p.C + c];
                *reinterpret_cast<uint2*>(dh + t) = make_uint2((w0 & 0xffffu) | (w1 << 16), (w2 & 0xffffu) | (w3 << 16));
                *reinterpret_cast<uint2*>(dl + t) = make_uint2((w0 >> 16) | (w1 & 0xffff0000u), (w2 >> 16) | (w3 & 0xffff0000u));
            }
        }
    }
    // per-channel reductions.  Global atomics from every CTA onto the same 3C addresses were the whole cost of
    // this kernel (L2 serialises them: ~5*C*B atomics per launch); instead the CTA reduces its T lanes in shared
    // memory, stores FiLM gradients plainly (one owner per (sample, channel)) and writes its bias / gamma / beta
    // partial sums to partials[b][3][C] for colsum_partials_kernel.
    if (p.partials) {
        float* red = reinterpret_cast<float*>(stage);   // [5][C], after the dy^T staging is consumed
        __syncthreads();
        for (int i = threadIdx.x; i < 5 * p.C; i += blockDim.x) red[i] = 0.0f;
        __syncthreads();
        if (active) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                atomicAdd(&red[c0 + j], dbi[j]);
                atomicAdd(&red[p.C + c0 + j], dga[j]);
                atomicAdd(&red[2 * p.C + c0 + j], dbe[j]);
                if (p.film) {
                    atomicAdd(&red[3 * p.C + c0 + j], dsc[j]);
                    atomicAdd(&red[4 * p.C + c0 + j], dsh[j]);
                }
            }
        }
        __syncthreads();
        float* part = p.partials + (int64_t)b * 3 * p.C;
        for (int i = threadIdx.x; i < 3 * p.C; i += blockDim.x) part[i] = red[i];
        if (p.film)
            for (int i = threadIdx.x; i < 2 * p.C; i += blockDim.x)
                p.dfilm[(int64_t)b * p.ld_dfilm + i] += red[3 * p.C + i];
        return;
    }
    if (!active) return;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        if (p.dbias) atomicAdd(&p.dbias[c0 + j], dbi[j]);
        atomicAdd(&p.dgamma[c0 + j], dga[j]);
        atomicAdd(&p.dbeta[c0 + j], dbe[j]);
        if (p.film) {
            atomicAdd(&p.dfilm[(int64_t)b * p.ld_dfilm + c0 + j], dsc[j]);
            atomicAdd(&p.dfilm[(int64_t)b * p.ld_dfilm + p.C + c0 + j], dsh[j]);
        }
    }
}

// out_k[c] += sum_b partials[b][k][c], k = bias | gamma | beta.  grid = ceil(3C / 32), block 32 x 8
__global__ void __launch_bounds__(256) colsum_partials_kernel(const float* __restrict__ partials, int B, int C,
                                                              float* dbias, float* dgamma, float* dbeta) {
    __shared__ float sh[8][33];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int col = blockIdx.x * 32 + tx;
    float s = 0.0f;
    if (col < 3 * C)
        for (int b = ty; b < B; b += 8) s += __ldg(&partials[(int64_t)b * 3 * C + col]);
    sh[ty][tx] = s;
    __syncthreads();
    if (ty == 0 && col < 3 * C) {
        float t = 0.0f;
#pragma unroll
        for (int i = 0; i < 8; ++i) t += sh[i][tx];
        const int k = col / C, c = col - k * C;
        float* dst = k == 0 ? dbias : (k == 1 ? dgamma : dbeta);
        if (dst) dst[c] += t;
    }
}

// ---------------------------------------------------------------------------
// transposed im2col of a bf16 hi/lo activation: x [B][Tin][C] -> [C*ntaps][B*Tout],
// row (c*ntaps + k), column (b*Tout + o) = x[b][stride*o + off[k]][c] (zero outside)
// ---------------------------------------------------------------------------
struct Im2colTParams {
    const __nv_bfloat16* x_hi; const __nv_bfloat16* x_lo;
    int ld_x, c_off;      // source row pitch and first channel (slice of a wider tensor)
    __nv_bfloat16* o_hi; __nv_bfloat16* o_lo;
    long long ld_out;
    int B, Tin, Tout, C, ntaps, stride;
    int off[8];
};
__global__ void __launch_bounds__(256) im2col_t_kernel(const Im2colTParams p) {
    // tile = 64 channels x 64 output columns (b, o); per tap: coalesced 128-byte channel rows in,
    // transposed through shared memory, 128-byte column runs out.  grid (C/64, cols/64).
    __shared__ uint32_t tile[64][65];   // (hi, lo) pairs
    const int c_base = blockIdx.x * 64;
    const int64_t col_base = (int64_t)blockIdx.y * 64;
    const int64_t ncols = (int64_t)p.B * p.Tout;
    const int64_t BT = p.ld_out;
    const uint16_t* xh = reinterpret_cast<const uint16_t*>(p.x_hi);
    const uint16_t* xl = reinterpret_cast<const uint16_t*>(p.x_lo);
    uint16_t* oh = reinterpret_cast<uint16_t*>(p.o_hi);
    uint16_t* ol = reinterpret_cast<uint16_t*>(p.o_lo);
    for (int k = 0; k < p.ntaps; ++k) {
        __syncthreads();
        for (int idx = threadIdx.x; idx < 64 * 64; idx += 256) {
            const int col = idx >> 6, ch = idx & 63;
            const int64_t gc = col_base + col;
            uint32_t w = 0;
            if (gc < ncols && c_base + ch < p.C) {
                const int bb = (int)(gc / p.Tout), o = (int)(gc % p.Tout);
                const int t = p.stride * o + p.off[k];
                if (t >= 0 && t < p.Tin) {
                    const int64_t src = ((int64_t)bb * p.Tin + t) * p.ld_x + p.c_off + c_base + ch;
                    w = (uint32_t)xh[src] | ((uint32_t)xl[src] << 16);
                }
            }
            tile[col][ch] = w;
        }
        __syncthreads();
        for (int idx = threadIdx.x; idx < 64 * 64; idx += 256) {
            const int ch = idx >> 6, col = idx & 63;
            const int64_t gc = col_base + col;
            if (gc < ncols && c_base + ch < p.C) {
                const uint32_t w = tile[col][ch];
                const int64_t r = (int64_t)(c_base + ch) * p.ntaps + k;
                oh[r * BT + gc] = (uint16_t)(w & 0xffffu);
                ol[r * BT + gc] = (uint16_t)(w >> 16);
            }
        }
    }
}

// dst[dst_off[r] + c] = src[r * ld + c]: the concatenated FiLM weight / bias gradients back into the
// per-block parameter-gradient windows (one launch instead of one copy per block)
__global__ void scatter_rows_kernel(const float* __restrict__ src, int ld, int64_t rows, int cols,
                                    const int64_t* __restrict__ dst_off, float* __restrict__ dst) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= rows * cols) return;
    const int64_t r = i / cols;
    const int c = (int)(i % cols);
    dst[dst_off[r] + c] = src[r * ld + c];
}

// ---------------------------------------------------------------------------
// gradient prep for convs whose output gradient arrives as plain fp32:
// dy [rows][ld] (C used) -> dy_hl [rows][C], dy^T_hl [C][rows], colsum[C] += sum_rows
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256) grad_prep_kernel(const float* __restrict__ dy, int64_t rows, int C,
                                                        int ld, __nv_bfloat16* hi, __nv_bfloat16* lo,
                                                        int ld_hl, __nv_bfloat16* t_hi, __nv_bfloat16* t_lo,
                                                        long long ld_T, float* colsum) {
    __shared__ float tile[32][33];
    const int c_base = blockIdx.x * 32, r_base = blockIdx.y * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
    for (int i = ty; i < 32; i += 8) {
        const int64_t r = r_base + i;
        const int c = c_base + tx;
        float v = 0.0f;
        if (r < rows && c < C) v = dy[r * ld + c];
        tile[i][tx] = v;
        if (hi && r < rows && c < ld_hl) {
            __nv_bfloat16 h, l;
            split_bf16(c < C ? v : 0.0f, h, l);
            hi[r * ld_hl + c] = h;
            lo[r * ld_hl + c] = l;
        }
    }
    __syncthreads();
    for (int i = ty; i < 32; i += 8) {
        const int c = c_base + i;
        const int64_t r = r_base + tx;
        if (c < C && r < rows && t_hi) {
            __nv_bfloat16 h, l;
            split_bf16(tile[tx][i], h, l);
            t_hi[(int64_t)c * ld_T + r] = h;
            t_lo[(int64_t)c * ld_T + r] = l;
        }
    }
    if (colsum && ty == 0) {
        float s = 0.0f;
        for (int i = 0; i < 32; ++i) s += tile[i][tx];
        if (c_base + tx < C) atomicAdd(&colsum[c_base + tx], s);
    }
}

// dx = dy * act'(x)  (act: 1 SiLU, 2 Mish), optional hi/lo split of the result
__global__ void act_bwd_kernel(const float* __restrict__ x, const float* __restrict__ dy, float* dx,
                               __nv_bfloat16* hi, __nv_bfloat16* lo, int64_t n, int act) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float xv = x[i];
    float g;
    if (act == 2) {
        g = mish_grad_f(xv);
    } else {
        const float sg = 1.0f / (1.0f + expf(-xv));
        g = sg * (1.0f + xv * (1.0f - sg));
    }
    const float r = dy[i] * g;
    if (dx) dx[i] = r;
    if (hi) {
        __nv_bfloat16 h, l;
        split_bf16(r, h, l);
        hi[i] = h;
        lo[i] = l;
    }
}

// dst[r][c] (+)= src[r][c] over a column window of two row-pitched fp32 matrices
__global__ void add_strided_kernel(float* dst, int ld_dst, const float* __restrict__ src, int ld_src,
                                   int64_t rows, int C, int accumulate) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= rows * C) return;
    const int64_t r = i / C;
    const int c = (int)(i % C);
    const float v = src[r * ld_src + c];
    float* d = dst + r * ld_dst + c;
    *d = accumulate ? *d + v : v;
}

// fused clip-by-global-norm + AdamW + EMA over one flat parameter slab
// (trainer: clip_grad_norm_(1.0); AdamW(lr, betas, eps, wd).step; EMA.update —
//  diffuser/libero/lb_online_trainer_v7.py:608-624)
__global__ void sumsq_kernel(const float* __restrict__ g, int64_t n, double* out) {
    float s = 0.0f;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const int64_t n4 = ((reinterpret_cast<uintptr_t>(g) & 15) == 0) ? (n >> 2) : 0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
        const float4 q = reinterpret_cast<const float4*>(g)[i];
        s += q.x * q.x + q.y * q.y + q.z * q.z + q.w * q.w;
    }
    for (int64_t i = (n4 << 2) + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
        s += g[i] * g[i];
    for (int off = 16; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
    __shared__ float w[32];
    if ((threadIdx.x & 31) == 0) w[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int i = 0; i < (int)(blockDim.x >> 5); ++i) t += (double)w[i];
        atomicAdd(out, t);
    }
}
__device__ __forceinline__ void adamw_ema_one(float& pi, const float gi_raw, float& mi, float& vi, float* ei,
                                              float clip, float lr, float beta1, float beta2, float eps, float wd,
                                              float bc1, float rsqrt_bc2, float ema_decay) {
    const float gi = gi_raw * clip;
    pi *= (1.0f - lr * wd);                            // decoupled weight decay
    mi = beta1 * mi + (1.0f - beta1) * gi;
    vi = beta2 * vi + (1.0f - beta2) * gi * gi;
    const float denom = sqrtf(vi) * rsqrt_bc2 + eps;
    pi -= (lr / bc1) * (mi / denom);
    if (ei) *ei = *ei - (1.0f - ema_decay) * (*ei - pi);  // ema.lerp_(p, 1 - decay)
}
// 16-byte accesses over the slab (7 HBM streams: read p g m v ema, write p m v ema); scalar tail
__global__ void adamw_ema_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                 float* __restrict__ v, float* __restrict__ ema, int64_t n,
                                 const double* __restrict__ sumsq, float max_norm, float lr, float beta1,
                                 float beta2, float eps, float wd, float bc1, float bc2, float ema_decay) {
    float clip = 1.0f;
    if (max_norm > 0.0f) {
        const float total = (float)sqrt(*sumsq);
        const float c = max_norm / (total + 1e-6f);   // torch clip_grad_norm_
        clip = c < 1.0f ? c : 1.0f;
    }
    const float rs2 = 1.0f / sqrtf(bc2);
    const int64_t n4 = n >> 2;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
        float4 pp = reinterpret_cast<float4*>(p)[i];
        const float4 gg = reinterpret_cast<const float4*>(g)[i];
        float4 mm = reinterpret_cast<float4*>(m)[i];
        float4 vv = reinterpret_cast<float4*>(v)[i];
        float4 ee = ema ? reinterpret_cast<float4*>(ema)[i] : make_float4(0.f, 0.f, 0.f, 0.f);
        adamw_ema_one(pp.x, gg.x, mm.x, vv.x, ema ? &ee.x : nullptr, clip, lr, beta1, beta2, eps, wd, bc1, rs2, ema_decay);
        adamw_ema_one(pp.y, gg.y, mm.y, vv.y, ema ? &ee.y : nullptr, clip, lr, beta1, beta2, eps, wd, bc1, rs2, ema_decay);
        adamw_ema_one(pp.z, gg.z, mm.z, vv.z, ema ? &ee.z : nullptr, clip, lr, beta1, beta2, eps, wd, bc1, rs2, ema_decay);
        adamw_ema_one(pp.w, gg.w, mm.w, vv.w, ema ? &ee.w : nullptr, clip, lr, beta1, beta2, eps, wd, bc1, rs2, ema_decay);
        reinterpret_cast<float4*>(p)[i] = pp;
        reinterpret_cast<float4*>(m)[i] = mm;
        reinterpret_cast<float4*>(v)[i] = vv;
        if (ema) reinterpret_cast<float4*>(ema)[i] = ee;
    }
    for (int64_t i = (n4 << 2) + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
        adamw_ema_one(p[i], g[i], m[i], v[i], ema ? &ema[i] : nullptr, clip, lr, beta1, beta2, eps, wd, bc1, rs2,
                      ema_decay);
}

// DDIM update of the action trajectory, in place (diffusers DDIMScheduler.step with eta = 0, restated in
// diffusion_policy.DDIMScheduler.step; one launch instead of ~10 elementwise torch kernels per denoise step).
// Operation order follows that torch chain (separate roundings, no FMA contraction).
__global__ void policy_ddim_step_kernel(float* __restrict__ x, int ldx, const float* __restrict__ mo, int ldm,
                                        int64_t rows, int C, float c1, float c2, float c3, float c4,
                                        int pred_sample, int clip) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= rows * C) return;
    const int64_t r = i / C;
    const int c = (int)(i - r * C);
    const float xt = x[r * ldx + c], m = mo[r * ldm + c];
    float x0, eps;
    if (pred_sample) {
        x0 = m;
        if (clip) x0 = x0 != x0 ? x0 : fminf(fmaxf(x0, -1.0f), 1.0f);
        eps = __fdiv_rn(__fsub_rn(xt, __fmul_rn(c2, x0)), c1);     // (sample - sqrt(a_t) x0) / sqrt(1 - a_t)
    } else {
        eps = m;
        x0 = __fdiv_rn(__fsub_rn(xt, __fmul_rn(c1, m)), c2);        // (sample - sqrt(1 - a_t) eps) / sqrt(a_t)
        if (clip) x0 = x0 != x0 ? x0 : fminf(fmaxf(x0, -1.0f), 1.0f);
    }
    x[r * ldx + c] = __fadd_rn(__fmul_rn(c3, x0), __fmul_rn(c4, eps));   // sqrt(a_prev) x0 + sqrt(1 - a_prev) eps
}

}  // namespace v2a

using namespace v2a;

#define POL_LAUNCH_OK()                  \
    do {                                 \
        V2A_CUDA_OK(cudaGetLastError()); \
        g_launches.fetch_add(1);         \
    } while (0)

static int check_gn(const v2a_policy_gn_desc* d) {
    V2A_REQUIRE(d->C % 8 == 0 && d->C / 8 <= kPolThreads, "policy_gn: C %d must be a multiple of 8, <= 2048", d->C);
    V2A_REQUIRE(d->groups >= 1 && d->groups <= 64 && d->C % d->groups == 0 && (d->C / d->groups) % 8 == 0,
                "policy_gn: channels per group must be a multiple of 8");
    const int octs = d->C / 8;
    const int tl = kPolThreads / octs > 0 ? kPolThreads / octs : 1;
    V2A_REQUIRE((d->T + tl - 1) / tl <= kMaxOct, "policy_gn: T*C = %d*%d too large for one CTA", d->T, d->C);
    return 0;
}

static GnActParams to_params(const v2a_policy_gn_desc* d) {
    GnActParams p;
    p.y = d->y; p.T = d->T; p.C = d->C; p.groups = d->groups; p.eps = d->eps;
    p.gamma = d->gamma; p.beta = d->beta;
    p.film = d->film; p.ld_film = d->ld_film;
    p.addend = d->addend; p.ld_add = d->ld_add;
    p.out_f32 = d->out_f32; p.ld_out = d->ld_out;
    p.out_hi = (__nv_bfloat16*)d->out_hi; p.out_lo = (__nv_bfloat16*)d->out_lo; p.ld_hl = d->ld_hl;
    p.mean_rstd = d->mean_rstd;
    p.dout = d->dout; p.ld_dout = d->ld_dout;
    p.dy_hi = (__nv_bfloat16*)d->dy_hi; p.dy_lo = (__nv_bfloat16*)d->dy_lo;
    p.dyT_hi = (__nv_bfloat16*)d->dyT_hi; p.dyT_lo = (__nv_bfloat16*)d->dyT_lo;
    p.ld_T = d->ld_T > 0 ? d->ld_T : (long long)d->B * d->T;
    p.dy_f32 = d->dy_f32;
    p.dbias = d->dbias; p.dgamma = d->dgamma; p.dbeta = d->dbeta;
    p.dfilm = d->dfilm; p.ld_dfilm = d->ld_dfilm;
    p.B = d->B;
    p.stage_words = 0;
    p.partials = d->partials;
    return p;
}

extern "C" {

int v2a_policy_gn_act_fwd(const v2a_policy_gn_desc* d, void* stream) {
    if (int rc = check_gn(d)) return rc;
    V2A_REQUIRE(d->y && d->gamma && d->beta && d->mean_rstd, "policy_gn_fwd: missing pointers");
    {
        // small batches: one CTA per (sample, group)  (V2A_POLICY_GN_GROUPS=0: the per-sample kernel, A/B probe)
        static const bool allow = [] { const char* e = getenv("V2A_POLICY_GN_GROUPS"); return !(e && atoi(e) == 0); }();
        const int cpg = d->C / d->groups, octs = cpg / 8;
        if (allow && d->B <= 16 && cpg % 8 == 0 && octs >= 1 && kPolThreads % octs == 0) {
            const int n_oct = d->T * octs;
            int threads = ((n_oct + kGrpOct - 1) / kGrpOct + 31) / 32 * 32;
            while (threads % octs) threads += 32;                     // every thread keeps ONE channel octet
            if (threads <= kPolThreads && threads * kGrpOct >= n_oct) {
                V2A_CUDA_OK(launch_maybe_pdl(gn_act_fwd_group_kernel, dim3((unsigned)d->B, (unsigned)d->groups),
                                             dim3((unsigned)threads), 0, (cudaStream_t)stream, to_params(d)));
                POL_LAUNCH_OK();
                return 0;
            }
        }
    }
    gn_act_fwd_kernel<<<d->B, kPolThreads, 0, (cudaStream_t)stream>>>(to_params(d));
    POL_LAUNCH_OK();
    return 0;
}

int v2a_policy_gn_act_bwd(const v2a_policy_gn_desc* d, void* stream) {
    if (int rc = check_gn(d)) return rc;
    V2A_REQUIRE(d->y && d->dout && d->mean_rstd && d->dgamma && d->dbeta, "policy_gn_bwd: missing pointers");
    V2A_REQUIRE(!d->film || d->dfilm, "policy_gn_bwd: FiLM needs dfilm");
    GnActParams p = to_params(d);
    // dy^T goes through shared memory (vector stores of T-long runs) when the sample fits 48 KB
    if (d->dyT_hi && d->T % 4 == 0 && (size_t)d->T * d->C * 4 <= 48 * 1024 && p.ld_T % 4 == 0)
        p.stage_words = d->T * d->C;
    size_t smem_words = (size_t)p.stage_words;
    if (p.partials && (size_t)5 * d->C > smem_words) smem_words = (size_t)5 * d->C;
    V2A_REQUIRE(smem_words * 4 <= 48 * 1024, "policy_gn_bwd: C %d too large for the shared-memory reductions", d->C);
    gn_act_bwd_kernel<<<d->B, kPolThreads, smem_words * 4, (cudaStream_t)stream>>>(p);
    if (p.partials) {
        POL_LAUNCH_OK();
        colsum_partials_kernel<<<(unsigned)((3 * d->C + 31) / 32), 256, 0, (cudaStream_t)stream>>>(
            p.partials, d->B, d->C, d->dbias, d->dgamma, d->dbeta);
    }
    POL_LAUNCH_OK();
    return 0;
}

int v2a_policy_im2col_t(const void* x_hi, const void* x_lo, int ld_x, int c_off, int B, int Tin, int Tout,
                        int C, int ntaps, int stride, const int* offsets, void* out_hi, void* out_lo,
                        int64_t ld_out, void* stream) {
    V2A_REQUIRE(ntaps >= 1 && ntaps <= 8, "im2col_t: ntaps out of range");
    Im2colTParams p;
    p.x_hi = (const __nv_bfloat16*)x_hi; p.x_lo = (const __nv_bfloat16*)x_lo;
    p.ld_x = ld_x; p.c_off = c_off;
    p.o_hi = (__nv_bfloat16*)out_hi; p.o_lo = (__nv_bfloat16*)out_lo;
    p.B = B; p.Tin = Tin; p.Tout = Tout; p.C = C; p.ntaps = ntaps; p.stride = stride;
    p.ld_out = ld_out > 0 ? ld_out : (long long)B * Tout;
    for (int i = 0; i < ntaps; ++i) p.off[i] = offsets[i];
    dim3 grid((unsigned)((C + 63) / 64), (unsigned)(((int64_t)B * Tout + 63) / 64));
    im2col_t_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(p);
    POL_LAUNCH_OK();
    return 0;
}

int v2a_scatter_rows(const float* src, int ld, int64_t rows, int cols, const int64_t* dst_off, float* dst,
                     void* stream) {
    V2A_REQUIRE(src && dst_off && dst && cols >= 1 && ld >= cols, "scatter_rows: bad arguments");
    const int64_t total = rows * cols;
    if (total == 0) return 0;
    scatter_rows_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(src, ld, rows, cols,
                                                                                        dst_off, dst);
    POL_LAUNCH_OK();
    return 0;
}

int v2a_grad_prep(const float* dy, int64_t rows, int C, int ld, void* hi, void* lo, int ld_hl, void* t_hi,
                  void* t_lo, int64_t ld_T, float* colsum, void* stream) {
    const int cw = ld_hl > C ? ld_hl : C;
    dim3 grid((unsigned)((cw + 31) / 32), (unsigned)((rows + 31) / 32));
    grad_prep_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(dy, rows, C, ld, (__nv_bfloat16*)hi,
                                                             (__nv_bfloat16*)lo, ld_hl, (__nv_bfloat16*)t_hi,
                                                             (__nv_bfloat16*)t_lo, ld_T > 0 ? ld_T : rows, colsum);
    POL_LAUNCH_OK();
    return 0;
}

int v2a_policy_ddim_step(float* x, int ldx, const float* model_out, int ldm, int64_t rows, int C, float sqrt_1m_at,
                         float sqrt_at, float sqrt_aprev, float coef_eps, int pred_sample, int clip, void* stream) {
    V2A_REQUIRE(x && model_out && rows >= 0 && C >= 1 && ldx >= C && ldm >= C, "policy_ddim_step: bad arguments");
    const int64_t n = rows * C;
    if (n == 0) return 0;
    policy_ddim_step_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        x, ldx, model_out, ldm, rows, C, sqrt_1m_at, sqrt_at, sqrt_aprev, coef_eps, pred_sample, clip);
    POL_LAUNCH_OK();
    return 0;
}

int v2a_act_bwd(const float* x, const float* dy, float* dx, void* hi, void* lo, int64_t n, int act,
                void* stream) {
    act_bwd_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        x, dy, dx, (__nv_bfloat16*)hi, (__nv_bfloat16*)lo, n, act);
    POL_LAUNCH_OK();
    return 0;
}

int v2a_add_strided(float* dst, int ld_dst, const float* src, int ld_src, int64_t rows, int C, int accumulate,
                    void* stream) {
    const int64_t n = rows * C;
    add_strided_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(dst, ld_dst, src, ld_src, rows,
                                                                                      C, accumulate);
    POL_LAUNCH_OK();
    return 0;
}

int v2a_grad_sumsq(const float* g, int64_t n, double* out, void* stream) {
    int blocks = (int)((n + 256 * 8 - 1) / (256 * 8));
    if (blocks > 148 * 8) blocks = 148 * 8;
    if (blocks < 1) blocks = 1;
    sumsq_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(g, n, out);
    POL_LAUNCH_OK();
    return 0;
}

int v2a_adamw_ema_step(float* p, const float* g, float* m, float* v, float* ema, int64_t n,
                       const double* grad_sumsq, float max_norm, float lr, float beta1, float beta2, float eps,
                       float weight_decay, int step, float ema_decay, void* stream) {
    const float bc1 = 1.0f - powf(beta1, (float)step);
    const float bc2 = 1.0f - powf(beta2, (float)step);
    V2A_REQUIRE(((uintptr_t)p | (uintptr_t)g | (uintptr_t)m | (uintptr_t)v | (uintptr_t)ema) % 16 == 0,
                "adamw_ema_step: slabs must be 16-byte aligned");
    int64_t blocks = ((n >> 2) + 255) / 256;
    if (blocks > 148 * 16) blocks = 148 * 16;   // grid-stride: a multiple of the SM count
    if (blocks < 1) blocks = 1;
    adamw_ema_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(
        p, g, m, v, ema, n, grad_sumsq, max_norm, lr, beta1, beta2, eps, weight_decay, bc1, bc2, ema_decay);
    POL_LAUNCH_OK();
    return 0;
}

}  // extern "C"
