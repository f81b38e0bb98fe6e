// HBM-bound kernels of the observation encoder (2 x ResNet18-GroupNorm + SpatialSoftmax + Linear), forward
// and backward, around the tensor-core convolutions (igemm.cu forward / data gradient, wgrad.cu weight
// gradient).  Activations are channels-last fp32 [image][y][x][C] plus bf16 hi/lo operand planes.
//
// replaces (autograd included)
//   ResNet18Conv trunk, BatchNorm -> GroupNorm(C/16)   diffusion_policy/common/vision_nets.py:9-39
//                                                      diffusion_policy/model/multi_image_obs_encoder.py:67-74
//   SpatialSoftmax                                     diffusion_policy/common/base_nets.py:234-285
//   VisualCore Linear(64, 64)                          diffusion_policy/common/vision_nets.py:113-143
#include "common.cuh"
#include "../../include/v2a_b200.h"

#include <atomic>

namespace v2a {
extern std::atomic<int64_t> g_launches;

#define V2A_ENC_LAUNCH_OK()                  \
    do {                                     \
        V2A_CUDA_OK(cudaGetLastError());     \
        v2a::g_launches.fetch_add(1);        \
    } while (0)

// ---------------------------------------------------------------------------
// stem operand: im2col of the 7x7 stride-2 pad-3 conv over a [B, 3, H, W] image (NCHW, as the policy
// receives it), k = c * 49 + ky * 7 + kx (= conv1.weight.view(64, 147)), zero padded to 192, split hi/lo.
// `scale * x + shift` is the policy's image normaliser (2x - 1) folded in; padding is zero AFTER it.
// one thread = 8 consecutive k of one output pixel
// ---------------------------------------------------------------------------
// one block = one output row of one image: the 7 input rows x 3 channels it needs are staged in shared memory
// (coalesced, normalised, zero padded) and every thread assembles 8 consecutive k of one output pixel from there
// (the first version gathered 8 scalars per thread from global memory: 0.50 ms at B = 256, LSU-bound)
__global__ void __launch_bounds__(256) enc_stem_pack_kernel(const float* __restrict__ x, float scale, float shift,
                                                            int B, int H, int W, __nv_bfloat16* __restrict__ hi,
                                                            __nv_bfloat16* __restrict__ lo, int fmt,
                                                            __nv_bfloat16* __restrict__ hi2,
                                                            __nv_bfloat16* __restrict__ lo2) {
    extern __shared__ float tile[];   // [3][7][W + 6]
    __shared__ short lut[192];        // k -> offset of (c, ky, kx) in the tile (-1: zero padding of K)
    const int Ho = H >> 1, Wo = W >> 1, Wp = W + 6;
    const int b = blockIdx.x / Ho, oy = blockIdx.x - b * Ho;
    if (threadIdx.x < 192) {
        const int k = threadIdx.x;
        const int c = k / 49, t = k - c * 49;
        const int ky = t / 7, kx = t - ky * 7;
        lut[k] = k < 147 ? (short)((c * 7 + ky) * Wp + kx) : (short)-1;
    }
    for (int i = threadIdx.x; i < 21 * Wp; i += blockDim.x) {
        const int xx = i % Wp, r = (i / Wp) % 7, c = i / (7 * Wp);
        const int iy = 2 * oy + r - 3, ix = xx - 3;
        float v = 0.0f;
        if (iy >= 0 && iy < H && ix >= 0 && ix < W) v = fmaf(__ldg(&x[(((int64_t)b * 3 + c) * H + iy) * W + ix]), scale, shift);
        tile[i] = v;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < Wo * 24; i += blockDim.x) {
        const int ox = i / 24, o8 = i - ox * 24;
        float v[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int o = lut[o8 * 8 + j];
            v[j] = o >= 0 ? tile[o + 2 * ox] : 0.0f;
        }
        const int64_t pix = ((int64_t)b * Ho + oy) * Wo + ox;
        uint4 h, l;
        split8_fmt(v, h, l, fmt);
        *reinterpret_cast<uint4*>(hi + pix * 192 + o8 * 8) = h;
        *reinterpret_cast<uint4*>(lo + pix * 192 + o8 * 8) = l;
        if (hi2) {   // bf16 twin: the weight-gradient GEMM pairs x with bf16 gradient planes (one format per MMA)
            split8(v, h, l);
            *reinterpret_cast<uint4*>(hi2 + pix * 192 + o8 * 8) = h;
            *reinterpret_cast<uint4*>(lo2 + pix * 192 + o8 * 8) = l;
        }
    }
}

// ---------------------------------------------------------------------------
// GroupNorm apply helpers: thread = 8 channels (one octet), constants per (image, octet)
// ---------------------------------------------------------------------------
struct Oct {
    float sc[8], mn[8], be[8];
};
__device__ __forceinline__ void load_oct(Oct& o, const float2* __restrict__ mr, int img, int groups, int C, int c,
                                         const float* __restrict__ gamma, const float* __restrict__ beta) {
    const int cpg = C / groups;
    int g = c / cpg, rem = c - g * cpg;
    const float2* mrp = mr + (int64_t)img * groups;
    float2 m = __ldg(&mrp[g]);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        o.sc[j] = m.y * __ldg(&gamma[c + j]);
        o.be[j] = __ldg(&beta[c + j]);
        o.mn[j] = m.x;
        if (++rem == cpg && j < 7) {
            rem = 0;
            m = __ldg(&mrp[++g]);
        }
    }
}
__device__ __forceinline__ void load8(const float* p, float* v) {
    const float4 a = ld_nc_f4(p), b = ld_nc_f4(p + 4);
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
__device__ __forceinline__ void store8(float* p, const float* v) {
    reinterpret_cast<float4*>(p)[0] = make_float4(v[0], v[1], v[2], v[3]);
    reinterpret_cast<float4*>(p)[1] = make_float4(v[4], v[5], v[6], v[7]);
}

// ---------------------------------------------------------------------------
// stem tail: GroupNorm -> ReLU -> MaxPool2d(3, stride 2, pad 1); [B, H, W, C] -> [B, H/2, W/2, C]
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256) enc_gn_relu_maxpool_kernel(const float* __restrict__ raw,
                                                                  const float2* __restrict__ mr, int groups,
                                                                  const float* __restrict__ gamma,
                                                                  const float* __restrict__ beta, int B, int H, int W,
                                                                  int C, float* __restrict__ out,
                                                                  __nv_bfloat16* __restrict__ hi,
                                                                  __nv_bfloat16* __restrict__ lo, int fmt,
                                                                  __nv_bfloat16* __restrict__ hi2,
                                                                  __nv_bfloat16* __restrict__ lo2) {
    const int oct = C >> 3;
    const int Ho = H >> 1, Wo = W >> 1;
    const int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= (int64_t)B * Ho * Wo * oct) return;
    const int c = (int)(gid % oct) * 8;
    const int64_t pix = gid / oct;
    const int ox = (int)(pix % Wo), oy = (int)((pix / Wo) % Ho), b = (int)(pix / ((int64_t)Wo * Ho));
    Oct o;
    load_oct(o, mr, b, groups, C, c, gamma, beta);
    float m[8] = {0, 0, 0, 0, 0, 0, 0, 0};   // ReLU output >= 0 and the centre tap is always inside
    for (int ky = 0; ky < 3; ++ky) {
        const int iy = 2 * oy + ky - 1;
        if (iy < 0 || iy >= H) continue;
        for (int kx = 0; kx < 3; ++kx) {
            const int ix = 2 * ox + kx - 1;
            if (ix < 0 || ix >= W) continue;
            float v[8];
            load8(raw + (((int64_t)b * H + iy) * W + ix) * C + c, v);
#pragma unroll
            for (int j = 0; j < 8; ++j) m[j] = fmaxf(m[j], fmaf(v[j] - o.mn[j], o.sc[j], o.be[j]));
        }
    }
    store8(out + pix * C + c, m);
    uint4 h, l;
    split8_fmt(m, h, l, fmt);
    *reinterpret_cast<uint4*>(hi + pix * C + c) = h;
    *reinterpret_cast<uint4*>(lo + pix * C + c) = l;
    if (hi2) {
        split8(m, h, l);
        *reinterpret_cast<uint4*>(hi2 + pix * C + c) = h;
        *reinterpret_cast<uint4*>(lo2 + pix * C + c) = l;
    }
}

// backward of the stem tail up to (not including) the GroupNorm: g[in pixel] = sum over the <= 4 pooling
// windows that contain the pixel and whose maximum it is (a == pooled value, a > 0) of dP[window]
__global__ void __launch_bounds__(256) enc_maxpool_relu_bwd_kernel(const float* __restrict__ raw,
                                                                   const float2* __restrict__ mr, int groups,
                                                                   const float* __restrict__ gamma,
                                                                   const float* __restrict__ beta,
                                                                   const float* __restrict__ pooled,
                                                                   const float* __restrict__ dpooled, int B, int H,
                                                                   int W, int C, float* __restrict__ g) {
    const int oct = C >> 3;
    const int Ho = H >> 1, Wo = W >> 1;
    const int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= (int64_t)B * H * W * oct) return;
    const int c = (int)(gid % oct) * 8;
    const int64_t pix = gid / oct;
    const int ix = (int)(pix % W), iy = (int)((pix / W) % H), b = (int)(pix / ((int64_t)W * H));
    Oct o;
    load_oct(o, mr, b, groups, C, c, gamma, beta);
    float v[8], a[8], r[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    load8(raw + pix * C + c, v);
#pragma unroll
    for (int j = 0; j < 8; ++j) a[j] = fmaxf(fmaf(v[j] - o.mn[j], o.sc[j], o.be[j]), 0.0f);
    // windows oy with 2 oy - 1 <= iy <= 2 oy + 1
    const int oy0 = iy >> 1, oy1 = (iy + 1) >> 1, ox0 = ix >> 1, ox1 = (ix + 1) >> 1;
    for (int oy = oy0; oy <= oy1; ++oy) {
        if (oy >= Ho) continue;
        for (int ox = ox0; ox <= ox1; ++ox) {
            if (ox >= Wo) continue;
            const int64_t q = (((int64_t)b * Ho + oy) * Wo + ox) * C + c;
            float pv[8], dv[8];
            load8(pooled + q, pv);
            load8(dpooled + q, dv);
#pragma unroll
            for (int j = 0; j < 8; ++j)
                if (a[j] > 0.0f && a[j] == pv[j]) r[j] += dv[j];
        }
    }
    store8(g + pix * C + c, r);
}

// ---------------------------------------------------------------------------
// block prep: out = relu( GN_a(xa) [+ idn | + GN_b(xb)] ) -> fp32 and/or hi/lo planes (normal or
// stride-2 phase-split layout [img][py*2+px][H/2][W/2][C])
// ---------------------------------------------------------------------------
struct EncPrepParams {
    const float* xa; const float2* mra; const float* gamma_a; const float* beta_a;
    const float* xb; const float2* mrb; const float* gamma_b; const float* beta_b;
    const float* idn;
    int groups, C, H, W;
    int64_t pixels;           // images * H * W
    int relu, phase_split, fmt;
    float* out_f32; __nv_bfloat16* out_hi; __nv_bfloat16* out_lo;
    __nv_bfloat16* out2_hi; __nv_bfloat16* out2_lo;   // optional bf16 twin of the planes (weight-gradient operand)
};
__global__ void __launch_bounds__(256, 3) enc_prep_kernel(const EncPrepParams p, int lanes, int iters) {
    const int oct = p.C >> 3;
    const int pl = threadIdx.x / oct;
    const int o8 = threadIdx.x - pl * oct;
    if (pl >= lanes) return;
    const int c = o8 * 8;
    const int HW = p.H * p.W;
    const int64_t base = (int64_t)blockIdx.x * (lanes * iters) + pl;
    int cur_img = -1;
    Oct oa, ob;
    for (int it = 0; it < iters; ++it) {
        const int64_t pix = base + (int64_t)it * lanes;
        if (pix >= p.pixels) break;
        const int img = (int)(pix / HW);
        if (img != cur_img) {
            cur_img = img;
            load_oct(oa, p.mra, img, p.groups, p.C, c, p.gamma_a, p.beta_a);
            if (p.xb) load_oct(ob, p.mrb, img, p.groups, p.C, c, p.gamma_b, p.beta_b);
        }
        float v[8];
        load8(p.xa + pix * p.C + c, v);
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = fmaf(v[j] - oa.mn[j], oa.sc[j], oa.be[j]);
        if (p.xb) {
            float w[8];
            load8(p.xb + pix * p.C + c, w);
#pragma unroll
            for (int j = 0; j < 8; ++j) v[j] += fmaf(w[j] - ob.mn[j], ob.sc[j], ob.be[j]);
        } else if (p.idn) {
            float w[8];
            load8(p.idn + pix * p.C + c, w);
#pragma unroll
            for (int j = 0; j < 8; ++j) v[j] += w[j];
        }
        if (p.relu) {
#pragma unroll
            for (int j = 0; j < 8; ++j) v[j] = fmaxf(v[j], 0.0f);
        }
        if (p.out_f32) store8(p.out_f32 + pix * p.C + c, v);
        if (p.out_hi) {
            int64_t oi = pix;
            if (p.phase_split) {
                const int rem = (int)(pix - (int64_t)img * HW);
                const int y = rem / p.W, x = rem - y * p.W;
                const int Hh = p.H >> 1, Wh = p.W >> 1;
                oi = (((int64_t)img * 4 + (y & 1) * 2 + (x & 1)) * Hh + (y >> 1)) * Wh + (x >> 1);
            }
            uint4 h, l;
            split8_fmt(v, h, l, p.fmt);
            *reinterpret_cast<uint4*>(p.out_hi + oi * p.C + c) = h;
            *reinterpret_cast<uint4*>(p.out_lo + oi * p.C + c) = l;
            if (p.out2_hi) {
                split8(v, h, l);
                *reinterpret_cast<uint4*>(p.out2_hi + oi * p.C + c) = h;
                *reinterpret_cast<uint4*>(p.out2_lo + oi * p.C + c) = l;
            }
        }
    }
}

// ---------------------------------------------------------------------------
// GroupNorm backward, two passes over the activation.
//   g   = dout * mask            mask: out > 0 (saved post-ReLU output), or GN(raw) > 0 (recomputed), or none
//   xh  = (raw - mean) * rstd
//   pass 1: sums[img][c] = (sum g, sum g * xh)
//   fin   : coef[img][grp] = (sum_c gamma_c sum_g, sum_c gamma_c sum_gxh) / m;  dgamma += sum_img sum_gxh; dbeta += ..
//   pass 2: draw = rstd * (gamma * g - coef.x - xh * coef.y)   -> hi/lo planes (+ optional fp32 g)
// ---------------------------------------------------------------------------
struct EncGnBwdParams {
    const float* dout; const float* outv;      // outv: saved post-ReLU output (mask) or null
    const float* raw; const float2* mr; const float* gamma; const float* beta;
    int mask_mode;                              // 0 none, 1 outv > 0, 2 GN(raw) > 0
    int groups, C, HW;
    int64_t pixels;
    float* sums;                                // [images][C][2]           (pass 1 out, pass 2 in)
    float inv_m;                                // 1 / (HW * C / groups)
    float* dgamma; float* dbeta;                // parameter gradients, accumulated by block 0 of pass 2
    __nv_bfloat16* d_hi; __nv_bfloat16* d_lo;   // draw planes              (pass 2 out)
    float* g_out;                               // optional fp32 g          (pass 2 out)
};

template <int PASS>
__global__ void __launch_bounds__(256, 3) enc_gn_bwd_kernel(const EncGnBwdParams p, int lanes, int iters) {
    __shared__ float red[2][256][9];   // pass 1 block reduction (padded); pass 2: the image's group coefficients
    const int oct = p.C >> 3;
    const int pl = threadIdx.x / oct;
    const int o8 = threadIdx.x - pl * oct;
    const bool active = pl < lanes;
    const int c = o8 * 8;
    // a block never straddles two images: blocks per image = ceil(HW / (lanes * iters))
    const int bpi = (p.HW + lanes * iters - 1) / (lanes * iters);
    const int img = blockIdx.x / bpi;
    const int first = (blockIdx.x - img * bpi) * (lanes * iters);
    float s[8] = {0, 0, 0, 0, 0, 0, 0, 0}, sx[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    if (PASS == 2) {
        // (the former finalize launch) group coefficients of this image from the pass-1 sums:
        // coef[g] = (sum_c gamma_c sum_g, sum_c gamma_c sum_gxh) / m
        float* cf = &red[0][0][0];       // [groups][2]
        const int cpg = p.C / p.groups;
        for (int g = threadIdx.x; g < p.groups; g += blockDim.x) {
            float a = 0.f, b2 = 0.f;
            for (int j = 0; j < cpg; ++j) {
                const int ch = g * cpg + j;
                const float gam = __ldg(&p.gamma[ch]);
                a += gam * p.sums[((int64_t)img * p.C + ch) * 2];
                b2 += gam * p.sums[((int64_t)img * p.C + ch) * 2 + 1];
            }
            cf[2 * g] = a * p.inv_m;
            cf[2 * g + 1] = b2 * p.inv_m;
        }
        if (blockIdx.x == img * bpi) {
            // the first block of every image adds the image's sums into the parameter gradients (a single block
            // walking all images serially was this launch's critical path: 37 us at 4x4x512)
            for (int ch = threadIdx.x; ch < p.C; ch += blockDim.x) {
                atomicAdd(&p.dbeta[ch], p.sums[((int64_t)img * p.C + ch) * 2]);
                atomicAdd(&p.dgamma[ch], p.sums[((int64_t)img * p.C + ch) * 2 + 1]);
            }
        }
        __syncthreads();
    }
    if (active) {
        const int cpg = p.C / p.groups;
        float mean[8], rstd[8], ga[8], be[8], cx[8], cy[8];
        {
            int g = c / cpg, rem = c - g * cpg;
            float2 m = __ldg(&p.mr[(int64_t)img * p.groups + g]);
            const float* cfs = &red[0][0][0];
            float2 cf = PASS == 2 ? make_float2(cfs[2 * g], cfs[2 * g + 1]) : make_float2(0.f, 0.f);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                mean[j] = m.x; rstd[j] = m.y; cx[j] = cf.x; cy[j] = cf.y;
                ga[j] = __ldg(&p.gamma[c + j]);
                be[j] = __ldg(&p.beta[c + j]);
                if (++rem == cpg && j < 7) {
                    rem = 0;
                    ++g;
                    m = __ldg(&p.mr[(int64_t)img * p.groups + g]);
                    if (PASS == 2) cf = make_float2(cfs[2 * g], cfs[2 * g + 1]);
                }
            }
        }
        for (int it = 0; it < iters; ++it) {
            const int q = first + it * lanes + pl;
            if (q >= p.HW) break;
            const int64_t e = ((int64_t)img * p.HW + q) * p.C + c;
            float dv[8], rv[8], xh[8];
            load8(p.dout + e, dv);
            load8(p.raw + e, rv);
#pragma unroll
            for (int j = 0; j < 8; ++j) xh[j] = (rv[j] - mean[j]) * rstd[j];
            if (p.mask_mode == 1) {
                float ov[8];
                load8(p.outv + e, ov);
#pragma unroll
                for (int j = 0; j < 8; ++j) dv[j] = ov[j] > 0.0f ? dv[j] : 0.0f;
            } else if (p.mask_mode == 2) {
#pragma unroll
                for (int j = 0; j < 8; ++j) dv[j] = fmaf(xh[j], ga[j], be[j]) > 0.0f ? dv[j] : 0.0f;
            }
            if (PASS == 1) {
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    s[j] += dv[j];
                    sx[j] += dv[j] * xh[j];
                }
            } else {
                float r[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) r[j] = rstd[j] * (ga[j] * dv[j] - cx[j] - xh[j] * cy[j]);
                uint4 h, l;
                split8(r, h, l);
                *reinterpret_cast<uint4*>(p.d_hi + e) = h;
                *reinterpret_cast<uint4*>(p.d_lo + e) = l;
                if (p.g_out) store8(p.g_out + e, dv);
            }
        }
    }
    if (PASS == 1) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            red[0][threadIdx.x][j] = s[j];
            red[1][threadIdx.x][j] = sx[j];
        }
        __syncthreads();
        // thread (o8, pl == 0) sums the lanes of its octet
        if (active && pl == 0) {
            for (int l2 = 1; l2 < lanes; ++l2) {
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    s[j] += red[0][l2 * oct + o8][j];
                    sx[j] += red[1][l2 * oct + o8][j];
                }
            }
            float* dst = p.sums + ((int64_t)img * p.C + c) * 2;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                atomicAdd(dst + 2 * j, s[j]);
                atomicAdd(dst + 2 * j + 1, sx[j]);
            }
        }
    }
}

// ---------------------------------------------------------------------------
// data gradient of a stride-2 3x3 conv arrives phase-blocked [img][H/2][W/2][(py, px)][C]; the 1x1 stride-2
// downsample path adds at phase (0, 0).  -> dx [img][H][W][C]
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256) enc_unblock_add_kernel(const float* __restrict__ blocked,
                                                              const float* __restrict__ ds, int images, int H, int W,
                                                              int C, float* __restrict__ dx) {
    const int q4 = C >> 2;
    const int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= (int64_t)images * H * W * q4) return;
    const int c = (int)(gid % q4) * 4;
    const int64_t pix = gid / q4;
    const int x = (int)(pix % W), y = (int)((pix / W) % H), img = (int)(pix / ((int64_t)W * H));
    const int Hh = H >> 1, Wh = W >> 1;
    const int64_t cell = ((int64_t)img * Hh + (y >> 1)) * Wh + (x >> 1);
    float4 v = *reinterpret_cast<const float4*>(blocked + (cell * 4 + (y & 1) * 2 + (x & 1)) * C + c);
    if (ds && !(y & 1) && !(x & 1)) {
        const float4 d = *reinterpret_cast<const float4*>(ds + cell * C + c);
        v.x += d.x; v.y += d.y; v.z += d.z; v.w += d.w;
    }
    *reinterpret_cast<float4*>(dx + pix * C + c) = v;
}

// ---------------------------------------------------------------------------
// SpatialSoftmax: logits [B][P][K] (P pixels, K keypoints) -> kp [B][K][2] = E_softmax[(pos_x, pos_y)]
// ---------------------------------------------------------------------------
__global__ void enc_spatial_softmax_fwd_kernel(const float* __restrict__ logits, int ld, int B, int P, int K,
                                               float inv_temp, const float* __restrict__ pos_x,
                                               const float* __restrict__ pos_y, float* __restrict__ att,
                                               float* __restrict__ kp) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B * K) return;
    const int b = i / K, k = i - b * K;
    const float* lp = logits + (int64_t)b * P * ld + k;
    float m = -INFINITY;
    for (int q = 0; q < P; ++q) m = fmaxf(m, lp[(int64_t)q * ld] * inv_temp);
    float den = 0.f, ex = 0.f, ey = 0.f;
    for (int q = 0; q < P; ++q) {
        const float e = expf(lp[(int64_t)q * ld] * inv_temp - m);
        den += e;
        ex += e * pos_x[q];
        ey += e * pos_y[q];
    }
    const float inv = 1.0f / den;
    for (int q = 0; q < P; ++q)
        att[((int64_t)b * P + q) * K + k] = expf(lp[(int64_t)q * ld] * inv_temp - m) * inv;
    kp[((int64_t)b * K + k) * 2] = ex * inv;
    kp[((int64_t)b * K + k) * 2 + 1] = ey * inv;
}

// dlogit[b][q][k] = att * ((px[q] - ex) dex + (py[q] - ey) dey) / T  -> hi/lo planes; dbias[k] += sum
__global__ void enc_spatial_softmax_bwd_kernel(const float* __restrict__ att, const float* __restrict__ kp,
                                               const float* __restrict__ dkp, int B, int P, int K, float inv_temp,
                                               const float* __restrict__ pos_x, const float* __restrict__ pos_y,
                                               __nv_bfloat16* __restrict__ d_hi, __nv_bfloat16* __restrict__ d_lo,
                                               float* __restrict__ dbias) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B * K) return;
    const int b = i / K, k = i - b * K;
    const float ex = kp[((int64_t)b * K + k) * 2], ey = kp[((int64_t)b * K + k) * 2 + 1];
    const float dex = dkp[((int64_t)b * K + k) * 2], dey = dkp[((int64_t)b * K + k) * 2 + 1];
    float tot = 0.f;
    for (int q = 0; q < P; ++q) {
        const int64_t e = ((int64_t)b * P + q) * K + k;
        const float d = att[e] * ((pos_x[q] - ex) * dex + (pos_y[q] - ey) * dey) * inv_temp;
        __nv_bfloat16 h, l;
        split_bf16(d, h, l);
        d_hi[e] = h;
        d_lo[e] = l;
        tot += d;
    }
    if (dbias) atomicAdd(&dbias[k], tot);
}

// Linear(IN -> OUT) backward for a small layer: dx = dy W;  dW += dy^T x;  db += sum dy
// grid: B blocks (dx) followed by OUT blocks (dW row o, db[o]); IN threads
__global__ void enc_linear_bwd_kernel(const float* __restrict__ x, const float* __restrict__ dy,
                                      const float* __restrict__ Wt, int B, int IN, int OUT, float* __restrict__ dx,
                                      float* __restrict__ dW, float* __restrict__ db) {
    const int j = threadIdx.x;
    if (j >= IN) return;
    if ((int)blockIdx.x < B) {
        const int b = blockIdx.x;
        float a = 0.f;
        for (int o = 0; o < OUT; ++o) a += dy[(int64_t)b * OUT + o] * Wt[(int64_t)o * IN + j];
        dx[(int64_t)b * IN + j] = a;
    } else {
        const int o = blockIdx.x - B;
        float a = 0.f, s = 0.f;
        for (int b = 0; b < B; ++b) {
            const float d = dy[(int64_t)b * OUT + o];
            a += d * x[(int64_t)b * IN + j];
            s += d;
        }
        dW[(int64_t)o * IN + j] += a;
        if (j == 0) db[o] += s;
    }
}

static void walk_shape(int C, int64_t pixels, int& lanes, int& iters) {
    const int oct = C / 8;
    lanes = 256 / oct;
    iters = 8;
    while (iters > 1 && pixels / ((int64_t)lanes * iters) < 4 * 148) iters >>= 1;
}

}  // namespace v2a

extern "C" {
using namespace v2a;

int v2a_enc_stem_pack(const float* x, float scale, float shift, int B, int H, int W, void* out_hi, void* out_lo,
                      int plane_fmt, void* twin_hi, void* twin_lo, void* stream) {
    V2A_REQUIRE(B >= 1 && H % 2 == 0 && W % 2 == 0, "enc_stem_pack: bad shape");
    V2A_REQUIRE(W <= 1024, "enc_stem_pack: image width %d exceeds the shared-memory row tile", W);   // lut offsets < 32768
    enc_stem_pack_kernel<<<(unsigned)(B * (H / 2)), 256, 21 * (W + 6) * sizeof(float), (cudaStream_t)stream>>>(
        x, scale, shift, B, H, W, (__nv_bfloat16*)out_hi, (__nv_bfloat16*)out_lo, plane_fmt,
        (__nv_bfloat16*)twin_hi, (__nv_bfloat16*)twin_lo);
    V2A_ENC_LAUNCH_OK();
    return 0;
}

int v2a_enc_gn_relu_maxpool(const float* raw, const float* mean_rstd, int groups, const float* gamma,
                            const float* beta, int B, int H, int W, int C, float* out, void* out_hi, void* out_lo,
                            int plane_fmt, void* twin_hi, void* twin_lo, void* stream) {
    V2A_REQUIRE(C % 8 == 0 && C % groups == 0 && H % 2 == 0 && W % 2 == 0, "enc_gn_relu_maxpool: bad shape");
    const int64_t total = (int64_t)B * (H / 2) * (W / 2) * (C / 8);
    enc_gn_relu_maxpool_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        raw, reinterpret_cast<const float2*>(mean_rstd), groups, gamma, beta, B, H, W, C, out,
        (__nv_bfloat16*)out_hi, (__nv_bfloat16*)out_lo, plane_fmt, (__nv_bfloat16*)twin_hi, (__nv_bfloat16*)twin_lo);
    V2A_ENC_LAUNCH_OK();
    return 0;
}

int v2a_enc_maxpool_relu_bwd(const float* raw, const float* mean_rstd, int groups, const float* gamma,
                             const float* beta, const float* pooled, const float* dpooled, int B, int H, int W, int C,
                             float* g, void* stream) {
    V2A_REQUIRE(C % 8 == 0 && C % groups == 0 && H % 2 == 0 && W % 2 == 0, "enc_maxpool_relu_bwd: bad shape");
    const int64_t total = (int64_t)B * H * W * (C / 8);
    enc_maxpool_relu_bwd_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        raw, reinterpret_cast<const float2*>(mean_rstd), groups, gamma, beta, pooled, dpooled, B, H, W, C, g);
    V2A_ENC_LAUNCH_OK();
    return 0;
}

int v2a_enc_prep(const v2a_enc_prep_desc* d, void* stream) {
    V2A_REQUIRE(d->C % 8 == 0 && d->C / 8 <= 256 && d->C % d->groups == 0, "enc_prep: bad channel count %d", d->C);
    V2A_REQUIRE(d->images >= 1 && d->H >= 1 && d->W >= 1, "enc_prep: bad shape");
    V2A_REQUIRE(!d->phase_split || (d->H % 2 == 0 && d->W % 2 == 0), "enc_prep: phase split needs even H, W");
    V2A_REQUIRE(!(d->xb && d->idn), "enc_prep: idn and a second normalised source are exclusive");
    EncPrepParams p;
    p.xa = d->xa; p.mra = reinterpret_cast<const float2*>(d->mean_rstd_a); p.gamma_a = d->gamma_a; p.beta_a = d->beta_a;
    p.xb = d->xb; p.mrb = reinterpret_cast<const float2*>(d->mean_rstd_b); p.gamma_b = d->gamma_b; p.beta_b = d->beta_b;
    p.idn = d->idn;
    p.groups = d->groups; p.C = d->C; p.H = d->H; p.W = d->W;
    p.pixels = (int64_t)d->images * d->H * d->W;
    p.relu = d->relu; p.phase_split = d->phase_split; p.fmt = d->plane_fmt;
    p.out_f32 = d->out_f32; p.out_hi = (__nv_bfloat16*)d->out_hi; p.out_lo = (__nv_bfloat16*)d->out_lo;
    p.out2_hi = (__nv_bfloat16*)d->out2_hi; p.out2_lo = (__nv_bfloat16*)d->out2_lo;
    V2A_REQUIRE(!p.out2_hi || p.out_hi, "enc_prep: twin planes need the primary planes");
    int lanes, iters;
    walk_shape(d->C, p.pixels, lanes, iters);
    const int64_t ppb = (int64_t)lanes * iters;
    enc_prep_kernel<<<(unsigned)((p.pixels + ppb - 1) / ppb), (d->C / 8) * lanes, 0, (cudaStream_t)stream>>>(p, lanes,
                                                                                                          iters);
    V2A_ENC_LAUNCH_OK();
    return 0;
}

int v2a_enc_gn_bwd(const v2a_enc_gn_bwd_desc* d, void* stream) {
    V2A_REQUIRE(d->C % 8 == 0 && d->C / 8 <= 256 && d->C <= 1024 && d->C % d->groups == 0,
                "enc_gn_bwd: bad channel count %d", d->C);
    V2A_REQUIRE(d->mask_mode != 1 || d->outv, "enc_gn_bwd: mask_mode 1 needs the saved output");
    EncGnBwdParams p;
    p.dout = d->dout; p.outv = d->outv; p.raw = d->raw; p.mr = reinterpret_cast<const float2*>(d->mean_rstd);
    p.gamma = d->gamma; p.beta = d->beta; p.mask_mode = d->mask_mode; p.groups = d->groups; p.C = d->C;
    p.HW = d->HW; p.pixels = (int64_t)d->images * d->HW;
    p.sums = d->sums;
    p.inv_m = 1.0f / ((float)d->HW * (float)(d->C / d->groups));
    p.dgamma = d->dgamma; p.dbeta = d->dbeta;
    p.d_hi = (__nv_bfloat16*)d->d_hi; p.d_lo = (__nv_bfloat16*)d->d_lo; p.g_out = d->g_out;
    V2A_REQUIRE(d->groups * 2 <= 2 * 256 * 9, "enc_gn_bwd: too many groups");
    int lanes, iters;
    walk_shape(d->C, p.pixels, lanes, iters);
    while (iters > 1 && lanes * iters > d->HW) iters >>= 1;
    const int bpi = ceil_div(d->HW, lanes * iters);
    const unsigned grid = (unsigned)(bpi * d->images);
    const int threads = (d->C / 8) * lanes;
    cudaStream_t st = (cudaStream_t)stream;
    // pass 1 (sums; the caller zeroes them once per backward) -> pass 2 (coefficients, parameter gradients, dx)
    enc_gn_bwd_kernel<1><<<grid, threads, 0, st>>>(p, lanes, iters);
    V2A_ENC_LAUNCH_OK();
    enc_gn_bwd_kernel<2><<<grid, threads, 0, st>>>(p, lanes, iters);
    V2A_ENC_LAUNCH_OK();
    return 0;
}

int v2a_enc_unblock_add(const float* blocked, const float* ds, int images, int H, int W, int C, float* dx,
                        void* stream) {
    V2A_REQUIRE(C % 4 == 0 && H % 2 == 0 && W % 2 == 0, "enc_unblock_add: bad shape");
    const int64_t total = (int64_t)images * H * W * (C / 4);
    enc_unblock_add_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(blocked, ds, images, H, W,
                                                                                            C, dx);
    V2A_ENC_LAUNCH_OK();
    return 0;
}

int v2a_enc_spatial_softmax_fwd(const float* logits, int ld, int B, int P, int K, float temperature,
                                const float* pos_x, const float* pos_y, float* att, float* kp, void* stream) {
    V2A_REQUIRE(B >= 1 && P >= 1 && K >= 1 && temperature > 0.f, "enc_spatial_softmax_fwd: bad shape");
    enc_spatial_softmax_fwd_kernel<<<ceil_div(B * K, 128), 128, 0, (cudaStream_t)stream>>>(
        logits, ld, B, P, K, 1.0f / temperature, pos_x, pos_y, att, kp);
    V2A_ENC_LAUNCH_OK();
    return 0;
}

int v2a_enc_spatial_softmax_bwd(const float* att, const float* kp, const float* dkp, int B, int P, int K,
                                float temperature, const float* pos_x, const float* pos_y, void* d_hi, void* d_lo,
                                float* dbias, void* stream) {
    V2A_REQUIRE(B >= 1 && P >= 1 && K >= 1 && temperature > 0.f, "enc_spatial_softmax_bwd: bad shape");
    enc_spatial_softmax_bwd_kernel<<<ceil_div(B * K, 128), 128, 0, (cudaStream_t)stream>>>(
        att, kp, dkp, B, P, K, 1.0f / temperature, pos_x, pos_y, (__nv_bfloat16*)d_hi, (__nv_bfloat16*)d_lo, dbias);
    V2A_ENC_LAUNCH_OK();
    return 0;
}

int v2a_enc_linear_bwd(const float* x, const float* dy, const float* W, int B, int IN, int OUT, float* dx, float* dW,
                       float* db, void* stream) {
    V2A_REQUIRE(IN >= 1 && IN <= 1024 && OUT >= 1 && B >= 1, "enc_linear_bwd: bad shape");
    enc_linear_bwd_kernel<<<B + OUT, IN, 0, (cudaStream_t)stream>>>(x, dy, W, B, IN, OUT, dx, dW, db);
    V2A_ENC_LAUNCH_OK();
    return 0;
}

}  // extern "C"
