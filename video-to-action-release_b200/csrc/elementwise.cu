// HBM-bound kernels around the tensor-core contractions: GroupNorm statistics
// and apply (+SiLU/Mish, concat, nearest upsample, stride-2 phase split, FiLM,
// bf16 hi/lo split), boundary layout packs, small dense layers, sampler steps.
// All loads/stores are 16-byte vectors over the contiguous channel axis.
#include "common.cuh"
#include "../../include/v2a_b200.h"

#include <atomic>
#include <cstdlib>

namespace v2a {
extern std::atomic<int64_t> g_launches;

#define V2A_LAUNCH_OK()                      \
    do {                                     \
        V2A_CUDA_OK(cudaGetLastError());     \
        v2a::g_launches.fetch_add(1);        \
    } while (0)

// ---------------------------------------------------------------------------
// per-(instance, channel) sum / sum of squares
// ---------------------------------------------------------------------------
// grid (chunks, instances); block = (C/4 quads) x pixel lanes
__global__ void channel_stats_kernel(const float* __restrict__ x, int64_t ppi, int C, int pix_per_block,
                                     double* __restrict__ stats) {
    extern __shared__ float sh[];  // [2*C]
    const int quads = C >> 2;
    const int lanes = blockDim.x / quads;
    const int q = threadIdx.x % quads;
    const int pl = threadIdx.x / quads;
    const int64_t inst = blockIdx.y;
    for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) sh[i] = 0.0f;
    __syncthreads();
    float s[4] = {0, 0, 0, 0}, ss[4] = {0, 0, 0, 0};
    if (pl < lanes) {
        const int64_t p0 = (int64_t)blockIdx.x * pix_per_block;
        int64_t p1 = p0 + pix_per_block;
        if (p1 > ppi) p1 = ppi;
        const float4* base = reinterpret_cast<const float4*>(x + (inst * ppi) * C) + q;
        for (int64_t p = p0 + pl; p < p1; p += lanes) {
            float4 v = __ldg(base + p * quads);
            s[0] += v.x; s[1] += v.y; s[2] += v.z; s[3] += v.w;
            ss[0] += v.x * v.x; ss[1] += v.y * v.y; ss[2] += v.z * v.z; ss[3] += v.w * v.w;
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            atomicAdd(&sh[(4 * q + j) * 2], s[j]);
            atomicAdd(&sh[(4 * q + j) * 2 + 1], ss[j]);
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 2 * C; i += blockDim.x)
        atomicAdd(&stats[inst * 2 * C + i], (double)sh[i]);
}

// (sample, group) -> mean, rstd from per-channel sums of up to two concat sources
// one warp per (sample, group): lanes stride over the (instance, channel) sums
__global__ void gn_finalize_kernel(const double* __restrict__ st0, const double* __restrict__ st1,
                                   int C0, int C1, int groups, int inst_per_group, double count,
                                   float eps, float2* __restrict__ mr, int samples, int rep0, long long rs0,
                                   int rep1, long long rs1) {
    const int idx = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (idx >= samples * groups) return;
    const int g = idx % groups, smp = idx / groups;
    const int cpg = (C0 + C1) / groups;
    double s = 0.0, ss = 0.0;
    for (int e = lane; e < inst_per_group * cpg; e += 32) {
        const int64_t inst = (int64_t)smp * inst_per_group + e / cpg;
        const int c = g * cpg + e % cpg;
        const bool first = c < C0;
        const double* base = first ? st0 + (inst * C0 + c) * 2 : st1 + (inst * C1 + (c - C0)) * 2;
        const int reps = first ? rep0 : rep1;
        const long long rs = first ? rs0 : rs1;
        for (int r = 0; r < reps; ++r) {
            const double2 v = *reinterpret_cast<const double2*>(base + r * rs);
            s += v.x;
            ss += v.y;
        }
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        s += __shfl_xor_sync(0xffffffffu, s, off);
        ss += __shfl_xor_sync(0xffffffffu, ss, off);
    }
    if (lane == 0) {
        const double mean = s / count;
        double var = ss / count - mean * mean;
        if (var < 0.0) var = 0.0;
        mr[idx] = make_float2((float)mean, (float)(1.0 / sqrt(var + (double)eps)));
    }
}

struct PrepParams {
    const float* x0; const float* x1;
    int C0, C1;
    const float2* mr;          // [samples][groups] or null
    int64_t pixels_per_sample; // pixels sharing one GroupNorm sample
    int groups;
    const float* gamma; const float* beta;
    int act;
    const float* film; int64_t pixels_per_film;
    int mode, H, W;
    int64_t P;
    __nv_bfloat16* out_hi; __nv_bfloat16* out_lo; float* out_f32;
    __nv_bfloat16* raw_hi; __nv_bfloat16* raw_lo;
};

// One thread = 8 channels x `iters` OUTPUT pixels.  blockDim.x = oct * lanes (oct = C / 8): thread
// (o8, pl) walks pixels base + pl, base + pl + lanes, ...  Everything that depends only on the channel octet
// (and the GroupNorm sample) -- gamma/beta, mean/rstd, the group walk -- is folded ONCE into a per-channel
// (mean, scale, shift) triple and reused for every pixel of the walk: the one-pixel-per-thread version spent ~45
// instructions per element (ncu: issue slots 77 % busy, DRAM 56 %) and was issue-bound, not HBM-bound.
__global__ void __launch_bounds__(256, 4) prep_kernel(const PrepParams p, int64_t total_out_pix, int lanes, int iters) {
    const int C = p.C0 + p.C1;
    const int oct = C >> 3;
    const int pl = threadIdx.x / oct;
    const int o8 = threadIdx.x - pl * oct;
    if (pl >= lanes) return;
    const int c = o8 * 8;
    const uint32_t base = blockIdx.x * (uint32_t)(lanes * iters) + pl;
    const bool from0 = c < p.C0;
    const float* const src_base = from0 ? p.x0 + c : p.x1 + (c - p.C0);
    const int src_ld = from0 ? p.C0 : p.C1;
    float sc[8], mn[8], be[8];
    uint32_t cur_smp = 0xffffffffu;
    const int cpg = p.mr ? C / p.groups : 1;
    const uint32_t total = (uint32_t)total_out_pix;
    // output pixel -> (input pixel, output index)
    auto map_pixel = [&](uint32_t op, uint32_t& ip, uint32_t& out_index) {
        ip = op;
        out_index = op;
        if (p.mode == 1) {
            const uint32_t W2 = p.W * 2, H2 = p.H * 2;
            const uint32_t row = op / W2;
            const uint32_t w = op - row * W2;
            const uint32_t img = row / H2;
            const uint32_t h = row - img * H2;
            ip = (img * p.H + (h >> 1)) * p.W + (w >> 1);
        } else if (p.mode == 2) {
            // iterate input pixels, scatter to [img][ph*2+pw][H/2][W/2]
            const uint32_t row = op / (uint32_t)p.W;
            const uint32_t w = op - row * p.W;
            const uint32_t img = row / (uint32_t)p.H;
            const uint32_t h = row - img * p.H;
            const uint32_t Hh = p.H >> 1, Wh = p.W >> 1;
            out_index = ((img * 4 + (h & 1) * 2 + (w & 1)) * Hh + (h >> 1)) * Wh + (w >> 1);
        }
    };
    // the loads of pixel it+1 are issued before pixel it is processed (two 16-byte loads always in flight)
    float4 na = make_float4(0.f, 0.f, 0.f, 0.f), nb = na;
    uint32_t nip = 0, nout = 0;
    if (base < total) {
        map_pixel(base, nip, nout);
        const float* src = src_base + (int64_t)nip * src_ld;
        na = ld_nc_f4(src);
        nb = ld_nc_f4(src + 4);
    }
    for (int it = 0; it < iters; ++it) {
        const uint32_t op = base + (uint32_t)(it * lanes);
        if (op >= total) break;
        const uint32_t ip = nip, out_index = nout;
        float v[8] = {na.x, na.y, na.z, na.w, nb.x, nb.y, nb.z, nb.w};
        if (it + 1 < iters && op + (uint32_t)lanes < total) {
            map_pixel(op + (uint32_t)lanes, nip, nout);
            const float* src = src_base + (int64_t)nip * src_ld;
            na = ld_nc_f4(src);
            nb = ld_nc_f4(src + 4);
        }
        if (p.raw_hi) {
            uint4 h, l;
            split8(v, h, l);
            *reinterpret_cast<uint4*>(p.raw_hi + (int64_t)out_index * C + c) = h;
            *reinterpret_cast<uint4*>(p.raw_lo + (int64_t)out_index * C + c) = l;
        }
        if (p.mr) {
            const uint32_t smp = ip / (uint32_t)p.pixels_per_sample;
            if (smp != cur_smp) {   // once per thread except where a pixel walk straddles two samples
                cur_smp = smp;
                int g = c / cpg, rem = c - g * cpg;   // walk the (at most 8) groups without divisions
                const float2* mrp = p.mr + (int64_t)smp * p.groups;
                float2 m = __ldg(&mrp[g]);
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    sc[j] = m.y * __ldg(&p.gamma[c + j]);
                    be[j] = __ldg(&p.beta[c + j]);
                    mn[j] = m.x;
                    if (++rem == cpg && j < 7) {
                        rem = 0;
                        m = __ldg(&mrp[++g]);
                    }
                }
            }
#pragma unroll
            for (int j = 0; j < 8; ++j) v[j] = fmaf(v[j] - mn[j], sc[j], be[j]);
        }
        if (p.act == 1) {
            // SiLU with the fast exp / divide intrinsics (rel. error ~2^-21, far below the 2^-17 of the
            // bf16 hi/lo split the result is stored in)
#pragma unroll
            for (int j = 0; j < 8; ++j) v[j] = __fdividef(v[j], 1.0f + __expf(-v[j]));
        } else if (p.act == 2) {
#pragma unroll
            for (int j = 0; j < 8; ++j) v[j] = mish_f(v[j]);
        }
        if (p.film) {
            const float* f = p.film + (int64_t)(ip / (uint32_t)p.pixels_per_film) * (2 * C);
#pragma unroll
            for (int j = 0; j < 8; ++j) v[j] = __ldg(&f[c + j]) * v[j] + __ldg(&f[C + c + j]);
        }
        if (p.out_hi) {
            uint4 h, l;
            split8(v, h, l);
            *reinterpret_cast<uint4*>(p.out_hi + (int64_t)out_index * C + c) = h;
            *reinterpret_cast<uint4*>(p.out_lo + (int64_t)out_index * C + c) = l;
        }
        if (p.out_f32) {
            float4* o = reinterpret_cast<float4*>(p.out_f32 + (int64_t)out_index * C + c);
            o[0] = make_float4(v[0], v[1], v[2], v[3]);
            o[1] = make_float4(v[4], v[5], v[6], v[7]);
        }
    }
}

// ---------------------------------------------------------------------------
// small dense layer: one warp per output feature, batch tiled by 8
// ---------------------------------------------------------------------------
__device__ __forceinline__ float act_f(float x, int a) {
    return a == 1 ? silu_f(x) : (a == 2 ? mish_f(x) : x);
}
__global__ void linear_kernel(const float* __restrict__ x, int ldx, const float* __restrict__ W,
                              const float* __restrict__ bias, const float* __restrict__ add, int ld_add,
                              float* __restrict__ y, int ldy, int B, int IN, int OUT, int act_in,
                              int act_out) {
    const int o = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (o >= OUT) return;
    const float* w = W + (int64_t)o * IN;
    for (int b0 = blockIdx.y * 8; b0 < B; b0 += gridDim.y * 8) {
        float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        for (int i = lane; i < IN; i += 32) {
            const float wv = __ldg(&w[i]);
#pragma unroll
            for (int j = 0; j < 8; ++j)
                if (b0 + j < B) acc[j] += act_f(__ldg(&x[(int64_t)(b0 + j) * ldx + i]), act_in) * wv;
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) {
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) acc[j] += __shfl_xor_sync(0xffffffffu, acc[j], off);
        }
        if (lane < 8 && b0 + lane < B) {
            float r = 0.f;
#pragma unroll
            for (int j = 0; j < 8; ++j)
                if (lane == j) r = acc[j];
            if (bias) r += bias[o];
            r = act_f(r, act_out);
            if (add) r += add[(int64_t)(b0 + lane) * ld_add + o];
            y[(int64_t)(b0 + lane) * ldy + o] = r;
        }
    }
}

__global__ void timestep_embedding_kernel(const int64_t* __restrict__ t, int B, int dim, int mode,
                                          float* __restrict__ out) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    const int half = dim / 2;
    if (idx >= B * half) return;
    const int b = idx / half, i = idx % half;
    const float tv = (float)t[b];
    // reference computes freqs in fp32: exp(-ln(1e4) * i / denom)
    float freq;
    if (mode == 0)  // th.exp(-log(1e4) * arange(half, fp32) / half)
        freq = expf(-9.210340371976184f * (float)i / (float)half);
    else            // th.exp(arange(half) * -(log(1e4) / (half - 1))), scalar rounded from double
        freq = expf((float)i * (float)(-(9.210340371976184 / (double)(half - 1))));
    const float a = tv * freq;
    if (mode == 0) {
        out[(int64_t)b * dim + i] = cosf(a);
        out[(int64_t)b * dim + half + i] = sinf(a);
    } else {
        out[(int64_t)b * dim + i] = sinf(a);
        out[(int64_t)b * dim + half + i] = cosf(a);
    }
}

// ---------------------------------------------------------------------------
// UNet boundary packs
// ---------------------------------------------------------------------------
// one thread = 8 of the 64 im2col channels of one pixel.  Sources are addressed
// through (batch, frame, channel) element strides so both the packed
// [B][(f c)][H][W] (+ broadcast cond frame) and the 5-D [B][6][F][H][W] layouts work.
struct PackStrides { int64_t b, f, c; };
__global__ void unet_input_pack_kernel(const float* __restrict__ x, PackStrides xs,
                                       const float* __restrict__ cond, PackStrides cs,
                                       int B, int F, int H, int W, __nv_bfloat16* __restrict__ ohi,
                                       __nv_bfloat16* __restrict__ olo) {
    const int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t npix = (int64_t)B * F * H * W;
    if (gid >= npix * 8) return;
    const int oc = (int)(gid & 7);  // which 8-channel octet of the 64
    const int64_t pix = gid >> 3;
    // 32-bit index arithmetic (the host checks npix < 2^31): the four 64-bit divisions per thread this replaced
    // made the kernel issue-bound at 1 TB/s
    const uint32_t p32 = (uint32_t)pix;
    const uint32_t row = p32 / (uint32_t)W;
    const int w = (int)(p32 - row * (uint32_t)W);
    const uint32_t img = row / (uint32_t)H;
    const int h = (int)(row - img * (uint32_t)H);
    const int b = (int)(img / (uint32_t)F);
    const int f = (int)(img - (uint32_t)b * (uint32_t)F);
    const float* const xb = x + b * xs.b + f * xs.f;
    const float* const cb = cond + b * cs.b + f * cs.f;
    float v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const int k = oc * 8 + j;
        float val = 0.0f;
        if (k < 54) {
            const int tap = k / 6, c = k - tap * 6;
            const int th = tap / 3;
            const int hh = h + th - 1, ww = w + (tap - th * 3) - 1;
            if (hh >= 0 && hh < H && ww >= 0 && ww < W) {
                const int sp = hh * W + ww;
                val = c < 3 ? __ldg(&xb[c * xs.c + sp]) : __ldg(&cb[(c - 3) * cs.c + sp]);
            }
        }
        v[j] = val;
    }
    uint4 hi, lo;
    split8(v, hi, lo);
    *reinterpret_cast<uint4*>(ohi + pix * 64 + oc * 8) = hi;
    *reinterpret_cast<uint4*>(olo + pix * 64 + oc * 8) = lo;
}

// temporal Conv1d(3,3,k=3, zero pad) over frames + NHWC -> strided (b, f, c) output
__global__ void unet_output_head_kernel(const float* __restrict__ y, int ldy, const float* __restrict__ wt,
                                        const float* __restrict__ bt, int B, int F, int H, int W,
                                        float* __restrict__ out, PackStrides os) {
    const int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t HW = (int64_t)H * W;
    if (gid >= (int64_t)B * F * HW) return;
    const int64_t hw = gid % HW;
    const int f = (int)((gid / HW) % F);
    const int b = (int)(gid / (HW * F));
    float acc[3] = {bt[0], bt[1], bt[2]};
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const int ff = f + k - 1;
        if (ff < 0 || ff >= F) continue;
        const float* yp = y + (((int64_t)b * F + ff) * HW + hw) * ldy;
        const float y0 = yp[0], y1 = yp[1], y2 = yp[2];
#pragma unroll
        for (int co = 0; co < 3; ++co)  // weight [co][ci][k]
            acc[co] += wt[(co * 3 + 0) * 3 + k] * y0 + wt[(co * 3 + 1) * 3 + k] * y1 +
                       wt[(co * 3 + 2) * 3 + k] * y2;
    }
#pragma unroll
    for (int co = 0; co < 3; ++co) out[b * os.b + f * os.f + co * os.c + hw] = acc[co];
}

// Second half of a 3x3 conv with very few output channels (the UNet's out head, 128 -> 3): the first half is a 1x1
// conv to 9*cout columns, P[pix][tap*cout + co] = sum_c a[pix][c] W[co][c][tap] (ONE pass over the activation on
// the tensor cores instead of nine TMA taps feeding a 16-wide N tile), this kernel gathers the nine taps:
//   y[n,h,w][co] = bias[co] + sum_{kh,kw} P[n, h+kh-1, w+kw-1][(kh*3+kw)*cout + co]      (zero outside the image)
template <int kCout>
__global__ void stencil9_kernel(const float* __restrict__ P, int ldp, const float* __restrict__ bias, int N, int H,
                                int W, float* __restrict__ y, int ldy) {
    const int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= (int64_t)N * H * W) return;
    const int w = (int)(gid % W);
    const int h = (int)((gid / W) % H);
    float acc[kCout];
#pragma unroll
    for (int co = 0; co < kCout; ++co) acc[co] = bias ? __ldg(bias + co) : 0.f;
#pragma unroll
    for (int kh = 0; kh < 3; ++kh) {
        const int hh = h + kh - 1;
        if (hh < 0 || hh >= H) continue;
#pragma unroll
        for (int kw = 0; kw < 3; ++kw) {
            const int ww = w + kw - 1;
            if (ww < 0 || ww >= W) continue;
            const float* pp = P + (gid + (int64_t)(kh - 1) * W + (kw - 1)) * ldp + (kh * 3 + kw) * kCout;
#pragma unroll
            for (int co = 0; co < kCout; ++co) acc[co] += __ldg(pp + co);
        }
    }
#pragma unroll
    for (int co = 0; co < kCout; ++co) y[gid * ldy + co] = acc[co];
}

// Tiled form of the same gather (the one that runs when the row stride allows 16-byte loads): a block stages the
// P rows of a (32 + 2) x (kTH + 2) pixel window in shared memory with coalesced loads (row stride odd -> the
// gather below is bank-conflict free), then each thread sums its nine taps from there.  The direct kernel above
// issues 9 * cout scalar loads per pixel, each warp instruction touching 32 different 128-byte rows: LSU-bound at
// 0.27 ms for 1.8 M pixels where the 235 MB of P take 0.05 ms to stream.
template <int kCout, int kTH>
__global__ void __launch_bounds__(32 * kTH) stencil9_tiled_kernel(const float* __restrict__ P, int ldp,
                                                                   const float* __restrict__ bias, int N, int H, int W,
                                                                   float* __restrict__ y, int ldy) {
    constexpr int TW = 32, RW = TW + 2, RH = kTH + 2, NV = 9 * kCout, Q = (NV + 3) / 4, RS = (4 * Q) | 1;
    __shared__ float tile[RH * RW * RS];
    const int tiles_w = (W + TW - 1) / TW, tiles_h = (H + kTH - 1) / kTH;
    int b = blockIdx.x;
    const int tw = b % tiles_w; b /= tiles_w;
    const int th = b % tiles_h;
    const int n = b / tiles_h;
    const int h0 = th * kTH, w0 = tw * TW;
    const float* const Pn = P + (int64_t)n * H * W * ldp;
    for (int idx = threadIdx.x; idx < RH * RW * Q; idx += 32 * kTH) {
        const int r = idx / Q, q = idx - r * Q;
        const int rh = r / RW, rw = r - rh * RW;
        const int hh = h0 + rh - 1, ww = w0 + rw - 1;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (hh >= 0 && hh < H && ww >= 0 && ww < W)
            v = __ldg(reinterpret_cast<const float4*>(Pn + (int64_t)(hh * W + ww) * ldp) + q);
        float* t = tile + r * RS + 4 * q;
        t[0] = v.x; t[1] = v.y; t[2] = v.z; t[3] = v.w;
    }
    __syncthreads();
    const int ty = threadIdx.x >> 5, tx = threadIdx.x & 31;
    const int h = h0 + ty, w = w0 + tx;
    if (h >= H || w >= W) return;
    float acc[kCout];
#pragma unroll
    for (int co = 0; co < kCout; ++co) acc[co] = bias ? __ldg(bias + co) : 0.f;
#pragma unroll
    for (int kh = 0; kh < 3; ++kh)
#pragma unroll
        for (int kw = 0; kw < 3; ++kw) {
            const float* t = tile + ((ty + kh) * RW + tx + kw) * RS + (kh * 3 + kw) * kCout;
#pragma unroll
            for (int co = 0; co < kCout; ++co) acc[co] += t[co];
        }
    float* yo = y + ((int64_t)n * H * W + (int64_t)h * W + w) * ldy;
#pragma unroll
    for (int co = 0; co < kCout; ++co) yo[co] = acc[co];
}

// ---------------------------------------------------------------------------
// sampler steps (operation order follows the reference's fp32 op chain)
// ---------------------------------------------------------------------------
// torch.clamp propagates NaN; fminf / fmaxf return the other operand, which would turn a diverged UNet output
// into a plausible-looking finite pixel.  Keep the reference's behaviour: NaN in, NaN out.
__device__ __forceinline__ float clamp_nan(float x, float lo, float hi) {
    return x != x ? x : fminf(fmaxf(x, lo), hi);
}

__global__ void ddpm_step_kernel(float* __restrict__ x, const float* __restrict__ v,
                                 const float* __restrict__ noise, const float* __restrict__ coef,
                                 int64_t n) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float sa = coef[0], s1 = coef[1], c1 = coef[2], c2 = coef[3], sigma = coef[4], vt = coef[5];
    const float xt = x[i];
    float x0 = __fsub_rn(__fmul_rn(sa, xt), __fmul_rn(s1, v[i]));  // predict_start_from_v
    x0 = clamp_nan(x0, -1.0f, 1.0f);                               // clamp_(-1, 1)
    const float mean = __fadd_rn(__fmul_rn(c1, x0), __fmul_rn(c2, xt));  // q_posterior mean
    const float nz = noise ? __fmul_rn(noise[i], vt) : 0.0f;
    x[i] = __fadd_rn(mean, __fmul_rn(sigma, nz));
}

__global__ void ddim_step_kernel(float* __restrict__ x, const float* __restrict__ v,
                                 const float* __restrict__ noise, const float* __restrict__ coef,
                                 int64_t n) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float sa = coef[0], s1 = coef[1], sr = coef[2], srm1 = coef[3];
    const float san = coef[4], cc = coef[5], sigma = coef[6], last = coef[7];
    const float xt = x[i];
    const float x0 = __fsub_rn(__fmul_rn(sa, xt), __fmul_rn(s1, v[i]));
    if (last != 0.0f) {
        x[i] = x0;
        return;
    }
    const float eps = __fdiv_rn(__fsub_rn(__fmul_rn(sr, xt), x0), srm1);  // predict_noise_from_start
    const float nz = noise ? noise[i] : 0.0f;
    // x0 * sqrt(a_next) + c * eps + sigma * noise, left to right
    x[i] = __fadd_rn(__fadd_rn(__fmul_rn(x0, san), __fmul_rn(cc, eps)), __fmul_rn(sigma, nz));
}

// Classifier-free guidance (guidance_weight > 0, objective pred_v; goal_diffusion.py:503-514,536-548): the UNet ran
// on the doubled batch [conditional | unconditional]; v holds both halves.  Mixing happens in NOISE space:
//   x0_c = sa x - s1 v_c,  x0_u = sa x - s1 v_u,  eps_* = (sr x - x0_*) / srm1,  eps = (1 + w) eps_c - w eps_u,
//   x0 = sr x - srm1 eps, then the ancestral (DDPM) or the eta-DDIM update; the new x is written to BOTH halves
// (the next UNet call reads the same image twice).  coef[8] = guidance weight; coef[6], coef[7] of the DDPM table
// carry sqrt_recip / sqrt_recipm1 (the DDIM table has them at [2], [3]).  n = elements of ONE half.
template <bool kDdim>
__global__ void cfg_step_kernel(float* __restrict__ x, const float* __restrict__ v, const float* __restrict__ noise,
                                const float* __restrict__ coef, int64_t n) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float sa = coef[0], s1 = coef[1], gw = coef[8];
    const float sr = kDdim ? coef[2] : coef[6], srm1 = kDdim ? coef[3] : coef[7];
    const float xt = x[i];
    const float x0c = __fsub_rn(__fmul_rn(sa, xt), __fmul_rn(s1, v[i]));
    const float x0u = __fsub_rn(__fmul_rn(sa, xt), __fmul_rn(s1, v[n + i]));
    const float srx = __fmul_rn(sr, xt);
    const float ec = __fdiv_rn(__fsub_rn(srx, x0c), srm1), eu = __fdiv_rn(__fsub_rn(srx, x0u), srm1);
    const float eps = __fsub_rn(__fmul_rn(1.0f + gw, ec), __fmul_rn(gw, eu));
    float x0 = __fsub_rn(srx, __fmul_rn(srm1, eps));
    const float nz = noise ? noise[i] : 0.0f;
    float out;
    if (kDdim) {
        const float san = coef[4], cc = coef[5], sigma = coef[6], last = coef[7];
        out = last != 0.0f ? x0 : __fadd_rn(__fadd_rn(__fmul_rn(x0, san), __fmul_rn(cc, eps)), __fmul_rn(sigma, nz));
    } else {
        const float c1 = coef[2], c2 = coef[3], sigma = coef[4], vt = coef[5];
        x0 = clamp_nan(x0, -1.0f, 1.0f);
        out = __fadd_rn(__fadd_rn(__fmul_rn(c1, x0), __fmul_rn(c2, xt)), __fmul_rn(sigma, __fmul_rn(nz, vt)));
    }
    x[i] = out;
    x[n + i] = out;
}

__global__ void unnormalize_clamp_kernel(const float* __restrict__ x, float* __restrict__ out, int64_t n) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float u = __fmul_rn(__fadd_rn(x[i], 1.0f), 0.5f);
    out[i] = clamp_nan(u, 0.0f, 1.0f);
}

__global__ void split_hl_kernel(const float* __restrict__ x, int64_t rows, int cols, int ld,
                                __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= rows * ld) return;
    const int c = (int)(i % ld);
    const int64_t r = i / ld;
    const float v = c < cols ? x[r * cols + c] : 0.0f;
    __nv_bfloat16 h, l;
    split_bf16(v, h, l);
    hi[i] = h;
    lo[i] = l;
}

// ---------------------------------------------------------------------------
// 64-bit content fingerprint of a list of fp32 tensors: sum over elements of bits(x_i) * (2 * global_index + 1)
// mod 2^64 (integer addition: exact and order independent, so the grid layout does not matter).  The engines use
// it to notice parameter updates that bypass autograd's version counter (`p.data.copy_()`, what
// ema_pytorch.EMA.update() does to the EMA model).
// ---------------------------------------------------------------------------
struct FpItem {
    const uint32_t* ptr;
    int64_t n;
    int64_t off;
};
__global__ void __launch_bounds__(256) fingerprint_kernel(const FpItem* __restrict__ items, int nitems,
                                                          unsigned long long* out) {
    // 16-byte loads, four in flight per thread (the scalar version read at 0.9 TB/s: 0.37 ms per predict_action call)
    unsigned long long h = 0;
    for (int it = blockIdx.x; it < nitems; it += gridDim.x) {
        const FpItem item = items[it];
        const int64_t head = min(item.n, (int64_t)(((16 - ((uintptr_t)item.ptr & 15)) & 15) >> 2));
        if (threadIdx.x < head)
            h += (unsigned long long)item.ptr[threadIdx.x] * (unsigned long long)(2 * (item.off + threadIdx.x) + 1);
        const uint4* v = reinterpret_cast<const uint4*>(item.ptr + head);
        const int64_t nv = (item.n - head) >> 2;
        const unsigned long long base = 2ull * (unsigned long long)(item.off + head) + 1ull;
#pragma unroll 4
        for (int64_t i = threadIdx.x; i < nv; i += blockDim.x) {
            const uint4 x = __ldg(v + i);
            const unsigned long long k = base + 8ull * (unsigned long long)i;
            h += (unsigned long long)x.x * k + (unsigned long long)x.y * (k + 2) + (unsigned long long)x.z * (k + 4) +
                 (unsigned long long)x.w * (k + 6);
        }
        const int64_t t = head + 4 * nv + threadIdx.x;
        if (t < item.n) h += (unsigned long long)item.ptr[t] * (unsigned long long)(2 * (item.off + t) + 1);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) h += __shfl_xor_sync(0xffffffffu, h, o);
    if ((threadIdx.x & 31) == 0 && h) atomicAdd(out, h);
}

// deterministic reduction of split-K slices: out[r][c] = x[0][r][c] + x[1][r][c] + ... (slice order) -> bf16 planes
__global__ void sum_slices_hl_kernel(const float* __restrict__ x, int slices, int64_t stride, int64_t n8,
                                     __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n8) return;
    float v[8];
    {
        const float4 a = __ldg(reinterpret_cast<const float4*>(x + i * 8));
        const float4 b = __ldg(reinterpret_cast<const float4*>(x + i * 8) + 1);
        v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
    }
    for (int s = 1; s < slices; ++s) {
        const float4 a = __ldg(reinterpret_cast<const float4*>(x + s * stride + i * 8));
        const float4 b = __ldg(reinterpret_cast<const float4*>(x + s * stride + i * 8) + 1);
        v[0] += a.x; v[1] += a.y; v[2] += a.z; v[3] += a.w; v[4] += b.x; v[5] += b.y; v[6] += b.z; v[7] += b.w;
    }
    uint4 h, l;
    split8(v, h, l);
    *reinterpret_cast<uint4*>(hi + i * 8) = h;
    *reinterpret_cast<uint4*>(lo + i * 8) = l;
}

}  // namespace v2a

using namespace v2a;

// ---------------------------------------------------------------------------
// table-driven weight repack: out[i] = map[i] ? src[map[i] - 1] : 0, written as bf16 hi/lo planes
// (tensor-core operand layout) or fp32 (bias / gain vectors).  The map is traced once per engine from
// the host-side packers (pure gathers), so re-packing after an optimiser step is ONE launch per arena.
// ---------------------------------------------------------------------------
__global__ void gather_split_kernel(const float* __restrict__ src, const int32_t* __restrict__ map, int64_t n,
                                    __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo,
                                    float* __restrict__ out_f32, int fmt) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x * 8;
    for (int64_t i0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 8; i0 < n; i0 += stride) {
        float v[8];
        if (i0 + 8 <= n) {
            const int4 m0 = *reinterpret_cast<const int4*>(map + i0);
            const int4 m1 = *reinterpret_cast<const int4*>(map + i0 + 4);
            const int32_t m[8] = {m0.x, m0.y, m0.z, m0.w, m1.x, m1.y, m1.z, m1.w};
#pragma unroll
            for (int j = 0; j < 8; ++j) v[j] = m[j] ? __ldg(src + (m[j] - 1)) : 0.0f;
            if (hi) {
                uint4 h, l;
                split8_fmt(v, h, l, fmt);
                *reinterpret_cast<uint4*>(hi + i0) = h;
                *reinterpret_cast<uint4*>(lo + i0) = l;
            }
            if (out_f32) {
                *reinterpret_cast<float4*>(out_f32 + i0) = make_float4(v[0], v[1], v[2], v[3]);
                *reinterpret_cast<float4*>(out_f32 + i0 + 4) = make_float4(v[4], v[5], v[6], v[7]);
            }
        } else {
            for (int64_t i = i0; i < n; ++i) {
                const int32_t m = map[i];
                const float x = m ? src[m - 1] : 0.0f;
                if (hi) {
                    if (fmt) {
                        const __half h = __float2half_rn(x);
                        const __half l = __float2half_rn(x - __half2float(h));
                        hi[i] = *reinterpret_cast<const __nv_bfloat16*>(&h);
                        lo[i] = *reinterpret_cast<const __nv_bfloat16*>(&l);
                    } else {
                        split_bf16(x, hi[i], lo[i]);
                    }
                }
                if (out_f32) out_f32[i] = x;
            }
        }
    }
}

extern "C" {

int v2a_channel_stats(const float* x, int64_t instances, int64_t ppi, int C, double* stats,
                      void* stream) {
    V2A_REQUIRE(C % 4 == 0 && C >= 4 && C / 4 <= 1024, "channel_stats: C %d unsupported", C);
    const int quads = C / 4;
    int threads = quads >= 256 ? quads : (256 / quads) * quads;
    if (threads > 1024) threads = quads;
    const int lanes = threads / quads;
    // ~512 pixels per lane-row keeps fp32 partial sums short
    int64_t pix_per_block = (int64_t)lanes * 128;
    if (pix_per_block > ppi) pix_per_block = ppi;
    const int64_t chunks = (ppi + pix_per_block - 1) / pix_per_block;
    V2A_REQUIRE(instances <= 65535, "channel_stats: too many instances");
    dim3 grid((unsigned)chunks, (unsigned)instances);
    channel_stats_kernel<<<grid, threads, 2 * C * sizeof(float), (cudaStream_t)stream>>>(
        x, ppi, C, (int)pix_per_block, stats);
    V2A_LAUNCH_OK();
    return 0;
}

int v2a_gn_finalize(const double* stats, int replicas, int64_t rep_stride, int instances, int C, int groups,
                    int64_t pixels_per_inst, float eps, float* mean_rstd, void* stream) {
    V2A_REQUIRE(C % groups == 0 && instances >= 1, "gn_finalize: bad shape");
    const int n = instances * groups;
    const double count = (double)pixels_per_inst * (double)(C / groups);
    gn_finalize_kernel<<<ceil_div(n * 32, 128), 128, 0, (cudaStream_t)stream>>>(
        stats, nullptr, C, 0, groups, 1, count, eps, reinterpret_cast<float2*>(mean_rstd), instances,
        replicas > 0 ? replicas : 1, rep_stride, 1, 0);
    V2A_LAUNCH_OK();
    return 0;
}

int v2a_prep(const v2a_prep_desc* d, void* stream) {
    const int C = d->C0 + d->C1;
    V2A_REQUIRE(C % 8 == 0 && d->C0 % 8 == 0, "prep: channel counts must be multiples of 8");
    V2A_REQUIRE(d->out_hi || d->out_f32 || d->raw_hi, "prep: no output");
    PrepParams p;
    p.x0 = d->x0; p.x1 = d->x1; p.C0 = d->C0; p.C1 = d->C1;
    p.mr = nullptr;
    p.groups = d->groups > 0 ? d->groups : 1;
    p.pixels_per_sample = d->pixels_per_inst * (d->inst_per_group > 0 ? d->inst_per_group : 1);
    p.gamma = d->gamma; p.beta = d->beta; p.act = d->act;
    p.film = d->film; p.pixels_per_film = d->pixels_per_film > 0 ? d->pixels_per_film : 1;
    p.mode = d->mode; p.H = d->H; p.W = d->W; p.P = d->P;
    p.out_hi = (__nv_bfloat16*)d->out_hi; p.out_lo = (__nv_bfloat16*)d->out_lo;
    p.out_f32 = d->out_f32;
    p.raw_hi = (__nv_bfloat16*)d->raw_hi; p.raw_lo = (__nv_bfloat16*)d->raw_lo;
    if (d->stats0) {
        V2A_REQUIRE(d->gamma && d->beta, "prep: GroupNorm needs gamma/beta");
        V2A_REQUIRE(C % d->groups == 0, "prep: C %d not divisible by groups %d", C, d->groups);
        V2A_REQUIRE(d->C1 == 0 || d->stats1, "prep: second source needs stats");
        V2A_REQUIRE(d->P % p.pixels_per_sample == 0, "prep: P not a multiple of the GroupNorm sample size");
        V2A_REQUIRE(d->gn_scratch != nullptr, "prep: gn_scratch missing");
        const int64_t inst = d->P / d->pixels_per_inst;
        const int samples = (int)(inst / d->inst_per_group);
        float2* mr = reinterpret_cast<float2*>(d->gn_scratch);
        const double count = (double)p.pixels_per_sample * (double)(C / d->groups);
        const int n = samples * d->groups;
        gn_finalize_kernel<<<ceil_div(n * 32, 128), 128, 0, (cudaStream_t)stream>>>(
            d->stats0, d->stats1, d->C0, d->C1, d->groups, d->inst_per_group, count, d->eps, mr, samples,
            d->stats_rep0 > 0 ? d->stats_rep0 : 1, d->stats_rep_stride0, d->stats_rep1 > 0 ? d->stats_rep1 : 1,
            d->stats_rep_stride1);
        V2A_LAUNCH_OK();
        p.mr = mr;
    }
    int64_t out_pix = d->P;
    if (d->mode == 1) out_pix = d->P * 4;
    if (d->mode == 1 || d->mode == 2)
        V2A_REQUIRE(d->H > 0 && d->W > 0 && d->P % ((int64_t)d->H * d->W) == 0 &&
                        (d->mode == 1 || (d->H % 2 == 0 && d->W % 2 == 0)),
                    "prep: bad H/W for resampling mode");
    const int64_t total = out_pix * (C / 8);
    V2A_REQUIRE(total < (int64_t)1 << 31 && p.pixels_per_sample < (int64_t)1 << 31 &&
                    (!d->film || d->pixels_per_film < (int64_t)1 << 31),
                "prep: %lld 8-channel chunks exceed the 32-bit index range of the kernel", (long long)total);
    // block = oct x lanes threads (<= 256); each thread walks `iters` pixels when the launch is large enough to
    // still fill the machine several times over
    const int oct = C / 8;
    V2A_REQUIRE(oct <= 256, "prep: %d channels exceed the 2048-channel block of the kernel", C);
    const int lanes = 256 / oct;
    int iters = 16;    // B200 sweep of the B=16 UNet's 74 prep launches (tools/ab_prep.py): 2: 16.2 ms, 4: 14.4, 8: 14.0, 16: 13.5, 32: 14.2
    {
        static const int env_iters = getenv("V2A_PREP_ITERS") ? atoi(getenv("V2A_PREP_ITERS")) : 0;   // tuning probe
        if (env_iters > 0) iters = env_iters;
    }
    while (iters > 1 && out_pix / ((int64_t)lanes * iters) < 4 * 148) iters >>= 1;
    const int64_t pix_per_block = (int64_t)lanes * iters;
    prep_kernel<<<(unsigned)((out_pix + pix_per_block - 1) / pix_per_block), oct * lanes, 0, (cudaStream_t)stream>>>(
        p, out_pix, lanes, iters);
    V2A_LAUNCH_OK();
    return 0;
}

int v2a_linear(const float* x, int ldx, const float* W, const float* bias, const float* add, int ld_add,
               float* y, int ldy, int B, int IN, int OUT, int act_in, int act_out, void* stream) {
    V2A_REQUIRE(B >= 1 && IN >= 1 && OUT >= 1, "linear: bad shape");
    const int warps = 4;
    int by = ceil_div(B, 8);
    if (by > 64) by = 64;
    dim3 grid(ceil_div(OUT, warps), by);
    linear_kernel<<<grid, warps * 32, 0, (cudaStream_t)stream>>>(x, ldx, W, bias, add, ld_add, y, ldy, B,
                                                                 IN, OUT, act_in, act_out);
    V2A_LAUNCH_OK();
    return 0;
}

int v2a_timestep_embedding(const int64_t* t, int B, int dim, int mode, float* out, void* stream) {
    V2A_REQUIRE(dim % 2 == 0 && dim >= 4, "timestep_embedding: dim must be even");
    const int n = B * dim / 2;
    timestep_embedding_kernel<<<ceil_div(n, 128), 128, 0, (cudaStream_t)stream>>>(t, B, dim, mode, out);
    V2A_LAUNCH_OK();
    return 0;
}

int v2a_unet_input_pack(const float* x, const int64_t* x_strides, const float* cond,
                        const int64_t* cond_strides, int B, int F, int H, int W, void* out_hi,
                        void* out_lo, void* stream) {
    const int64_t total = (int64_t)B * F * H * W * 8;
    V2A_REQUIRE(total / 8 < ((int64_t)1 << 31), "unet_input_pack: more than 2^31 pixels");
    PackStrides xs{x_strides[0], x_strides[1], x_strides[2]};
    PackStrides cs{cond_strides[0], cond_strides[1], cond_strides[2]};
    unet_input_pack_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        x, xs, cond, cs, B, F, H, W, (__nv_bfloat16*)out_hi, (__nv_bfloat16*)out_lo);
    V2A_LAUNCH_OK();
    return 0;
}

int v2a_unet_output_head(const float* y, int ldy, const float* wt, const float* bt, int B, int F, int H,
                         int W, float* out, const int64_t* out_strides, void* stream) {
    const int64_t total = (int64_t)B * F * H * W;
    PackStrides os{out_strides[0], out_strides[1], out_strides[2]};
    unet_output_head_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        y, ldy, wt, bt, B, F, H, W, out, os);
    V2A_LAUNCH_OK();
    return 0;
}

int v2a_stencil9(const float* P, int ldp, const float* bias, int N, int H, int W, int cout, float* y, int ldy,
                 void* stream) {
    V2A_REQUIRE(P && y && cout >= 1 && cout <= 4 && ldp >= 9 * cout && ldy >= cout, "stencil9: bad arguments");
    const int64_t total = (int64_t)N * H * W;
    const unsigned blocks = (unsigned)((total + 255) / 256);
    cudaStream_t st = (cudaStream_t)stream;
    if (cout == 3 && ldp % 4 == 0 && ldp >= 28 && ((uintptr_t)P & 15) == 0 && (int64_t)H * W < ((int64_t)1 << 30)) {
        const unsigned tiles = (unsigned)N * (unsigned)((H + 7) / 8) * (unsigned)((W + 31) / 32);
        stencil9_tiled_kernel<3, 8><<<tiles, 256, 0, st>>>(P, ldp, bias, N, H, W, y, ldy);
        V2A_LAUNCH_OK();
        return 0;
    }
    switch (cout) {
        case 1: stencil9_kernel<1><<<blocks, 256, 0, st>>>(P, ldp, bias, N, H, W, y, ldy); break;
        case 2: stencil9_kernel<2><<<blocks, 256, 0, st>>>(P, ldp, bias, N, H, W, y, ldy); break;
        case 3: stencil9_kernel<3><<<blocks, 256, 0, st>>>(P, ldp, bias, N, H, W, y, ldy); break;
        default: stencil9_kernel<4><<<blocks, 256, 0, st>>>(P, ldp, bias, N, H, W, y, ldy); break;
    }
    V2A_LAUNCH_OK();
    return 0;
}

int v2a_ddpm_step(float* x, const float* v, const float* noise, const float* coef, int64_t n,
                  void* stream) {
    ddpm_step_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(x, v, noise, coef, n);
    V2A_LAUNCH_OK();
    return 0;
}

int v2a_ddim_step(float* x, const float* v, const float* noise, const float* coef, int64_t n,
                  void* stream) {
    ddim_step_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(x, v, noise, coef, n);
    V2A_LAUNCH_OK();
    return 0;
}

int v2a_cfg_step(float* x, const float* v, const float* noise, const float* coef, int64_t n_half, int ddim,
                 void* stream) {
    const unsigned blocks = (unsigned)((n_half + 255) / 256);
    if (ddim) cfg_step_kernel<true><<<blocks, 256, 0, (cudaStream_t)stream>>>(x, v, noise, coef, n_half);
    else cfg_step_kernel<false><<<blocks, 256, 0, (cudaStream_t)stream>>>(x, v, noise, coef, n_half);
    V2A_LAUNCH_OK();
    return 0;
}

int v2a_unnormalize_clamp(const float* x, float* out, int64_t n, void* stream) {
    unnormalize_clamp_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(x, out, n);
    V2A_LAUNCH_OK();
    return 0;
}

int v2a_gather_split(const float* src, const int32_t* map, int64_t n, void* out_hi, void* out_lo,
                     float* out_f32, void* stream) {
    return v2a_gather_split_fmt(src, map, n, out_hi, out_lo, out_f32, 0, stream);
}

int v2a_gather_split_fmt(const float* src, const int32_t* map, int64_t n, void* out_hi, void* out_lo,
                         float* out_f32, int plane_fmt, void* stream) {
    V2A_REQUIRE(src && map && n >= 0, "gather_split: missing pointers");
    V2A_REQUIRE((out_hi != nullptr) == (out_lo != nullptr) && (out_hi || out_f32), "gather_split: no output");
    V2A_REQUIRE(((uintptr_t)map | (uintptr_t)out_hi | (uintptr_t)out_lo | (uintptr_t)out_f32) % 16 == 0,
                "gather_split: map / outputs must be 16-byte aligned");
    if (n == 0) return 0;
    int64_t blocks = (n / 8 + 255) / 256;
    if (blocks > 148 * 16) blocks = 148 * 16;
    if (blocks < 1) blocks = 1;
    gather_split_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(
        src, map, n, (__nv_bfloat16*)out_hi, (__nv_bfloat16*)out_lo, out_f32, plane_fmt);
    V2A_LAUNCH_OK();
    return 0;
}

int v2a_sum_slices_hl(const float* x, int slices, int64_t stride, int64_t rows, int cols, void* out_hi, void* out_lo,
                      void* stream) {
    V2A_REQUIRE(x && out_hi && out_lo && slices >= 1 && rows >= 0 && cols % 8 == 0 && stride % 4 == 0 &&
                    (slices == 1 || stride >= rows * cols),
                "sum_slices_hl: cols must be a multiple of 8, slices at least rows * cols apart");
    const int64_t n8 = rows * cols / 8;
    if (n8 == 0) return 0;
    sum_slices_hl_kernel<<<(unsigned)((n8 + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        x, slices, stride, n8, (__nv_bfloat16*)out_hi, (__nv_bfloat16*)out_lo);
    V2A_LAUNCH_OK();
    return 0;
}

int v2a_split_hl(const float* x, int64_t rows, int cols, int ld_out, void* out_hi, void* out_lo,
                 void* stream) {
    V2A_REQUIRE(ld_out >= cols, "split_hl: ld_out < cols");
    const int64_t total = rows * ld_out;
    split_hl_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        x, rows, cols, ld_out, (__nv_bfloat16*)out_hi, (__nv_bfloat16*)out_lo);
    V2A_LAUNCH_OK();
    return 0;
}

int v2a_params_fingerprint(const void* items, int nitems, uint64_t* out, void* stream) {
    V2A_REQUIRE(items && out && nitems >= 0, "params_fingerprint: missing pointers");
    V2A_CUDA_OK(cudaMemsetAsync(out, 0, sizeof(uint64_t), (cudaStream_t)stream));
    if (nitems == 0) return 0;
    const int blocks = nitems < 148 * 8 ? nitems : 148 * 8;
    fingerprint_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>((const FpItem*)items, nitems,
                                                                 (unsigned long long*)out);
    V2A_LAUNCH_OK();
    return 0;
}

}  // extern "C"
