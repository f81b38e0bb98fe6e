/* v2a_b200 — C ABI of the B200-native hot paths of video-to-action-release.
 *
 * The reference (pure Python/PyTorch) has no FFI of its own; its "operator
 * interface" for the two hot paths is the nn.Module call surface listed in
 * SURVEY.md §8(b).  This header is the boundary our Python host modules bind
 * with ctypes.  Each entry point cites the reference code it replaces
 * (paths relative to the reference checkout).
 *
 * Conventions
 *  - plain C, raw device pointers, sizes as int / int64_t; no torch types.
 *  - the caller allocates every buffer; nothing here mallocs device memory
 *    or synchronises.  Work is enqueued on `stream` (a cudaStream_t passed as
 *    void*), so calls are CUDA-graph capturable.
 *  - return 0 on success; non-zero on failure, message via v2a_last_error().
 *  - activations are channels-last.  "hl" tensors are two bf16 planes
 *    (hi, lo) with x ~= hi + lo: the operand format of every tensor-core
 *    contraction (3-pass bf16 split product, fp32 accumulate in TMEM).
 */
#ifndef V2A_B200_H
#define V2A_B200_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

#define V2A_MAX_SRC 2
#define V2A_MAX_TAPS 16

const char* v2a_last_error(void);
int v2a_version(void);

/* ------------------------------------------------------------------------
 * Implicit-GEMM convolution / linear on tcgen05 tensor cores.
 *
 * out[row, n] = sum_taps sum_c A_src[pixel(row) + tap.d, c] * W[n, k(tap, c)]
 *
 * replaces, depending on the tap program:
 *   nn.Conv2d 3x3 (spatial_conv)         guided_diffusion/nn.py:45,66
 *   nn.Conv1d k3 over frames (temporal)  guided_diffusion/nn.py:46,76-85
 *   1x1 skip_connection                  guided_diffusion/unet.py:225
 *   AttentionBlock qkv / proj_out        guided_diffusion/unet.py:290,298
 *   Conv1d k5/k3/k1, Linear (policy)     diffusion_policy/model/conv1d_components.py:7-40
 * ---------------------------------------------------------------------- */
typedef struct v2a_igemm_src {
    const void* hi;   /* bf16 [X3][X2][X1][X0][C] */
    const void* lo;   /* bf16 same shape (ignored when passes == 1) */
    int channels;     /* C, multiple of 8 */
    int dims[4];      /* X0..X3, X0 fastest */
} v2a_igemm_src;

typedef struct v2a_igemm_tap {
    int src;      /* index into desc.src */
    int d[4];     /* coordinate offset of this tap in X0..X3 (may be negative: zero padding) */
    int nchunks;  /* ceil(C / 64) K-chunks taken from this tap */
} v2a_igemm_tap;

typedef struct v2a_igemm_desc {
    v2a_igemm_src src[V2A_MAX_SRC];
    int nsrc;
    v2a_igemm_tap taps[V2A_MAX_TAPS];
    int ntaps;
    const void* w_hi;  /* bf16 [wrows][ktot], K-major; k runs tap-major, 64 per chunk */
    const void* w_lo;
    int wrows;         /* rows present in the weight matrix (>= cout) */
    int ktot;          /* 64 * sum(nchunks) */
    int out_dims[4];   /* output pixel grid D0..D3 (D0 fastest); rows = prod */
    int tile_log2[4];  /* log2 of the 128-row tile box along D0..D3 (sums to 7) */
    int block_n;       /* N tile: multiple of 16, <= 256 */
    int passes;        /* 3 = hi*hi + hi*lo + lo*hi (fp32-class); 1 = bf16 only */
    int cout;          /* valid output channels */
    int ldc;           /* output row pitch in elements, multiple of 16 and >= cout rounded up to 16 */
    float* out_f32;    /* [rows][ldc] or NULL */
    void* out_hi;      /* bf16 [rows][ldc] or NULL */
    void* out_lo;
    const float* bias;      /* [cout] or NULL */
    const float* rowvec;    /* per-row-group additive vector [groups][ld_rowvec] or NULL */
    int ld_rowvec;
    int rowvec_mul[4];      /* group = sum coord[d] * rowvec_mul[d] */
    const float* residual;  /* fp32 [rows][ld_res] added to the output, or NULL */
    int ld_res;
    double* stats;          /* per (instance, channel) {sum, sumsq} of the written fp32 value, or NULL */
    int stats_mul[4];       /* instance = sum coord[d] * stats_mul[d] */
    int stats_ld;           /* channels per instance in the stats buffer */
    int stats_replicas;     /* >= 1: CTAs spread their atomics over this many copies of the buffer */
    int64_t stats_rep_stride;   /* doubles between consecutive copies */
    int a_fp16;             /* source planes are fp16 (hi, lo) pairs instead of bf16 ones (22+ bit operands) */
    int b_fp16;             /* same for the weight planes */
    int64_t out_pix_mul[4]; /* all zero: output rows follow the grid densely (D0 fastest).  Otherwise the output row */
    int64_t out_pix_off;    /* (and residual row) of grid point c is out_pix_off + sum c[d] * out_pix_mul[d]: the four
                               sub-pixel phases of `Upsample` (F.interpolate(nearest, x2) -> 3x3 conv,
                               guided_diffusion/unet.py:107-114) are 2x2-tap convs over the LOW-resolution grid whose
                               results interleave on the fine grid */
    int64_t split_stride;   /* > 0: when the plan splits K (v2a_igemm_plan_k_splits > 1), split s writes its partial sums
                               with plain stores to out_f32 + s * split_stride (elements) instead of adding them
                               atomically into out_f32 -- the caller reduces the slices in a fixed order
                               (v2a_sum_slices_hl), so the result does not depend on the arrival order of the CTAs;
                               such plans use at most 16 splits */
} v2a_igemm_desc;

int v2a_igemm_plan_create(const v2a_igemm_desc* desc, void** plan_out);
int v2a_igemm_plan_run(void* plan, void* stream);
void v2a_igemm_plan_destroy(void* plan);
/* how many CTAs share one output tile's K loop (1 = no split-K); introspection for tests / probes */
int v2a_igemm_plan_k_splits(void* plan);
/* Conv3d as ONE launch (guided_diffusion/nn.py:53-87: Conv2d on every frame, then the zero-padded Conv1d(k = 3) over
 * frames): the spatial program and the temporal program that consumes its output planes, interleaved tile by tile
 * in one persistent kernel so the temporal conv's epilogue runs under the spatial conv's MMAs.  Both descriptors
 * must be pair launches with the fused split product (passes 3, block_n <= 128, one N tile) over the SAME dense
 * row space cut into the same 128-row tiles; `tiles_per_frame` = tiles of one (sample, frame) = the distance between
 * the tiles the temporal taps read; `flags` = caller-allocated uint32 [tiles] scratch (cleared by every run). */
int v2a_igemm_dual_plan_create(const v2a_igemm_desc* spatial, const v2a_igemm_desc* temporal, int frames,
                               int tiles_per_frame, void* flags, void** plan_out);
int v2a_igemm_dual_plan_run(void* plan, void* stream);
void v2a_igemm_dual_plan_destroy(void* plan);
/* kernel launches performed by v2a_* calls since process start */
int64_t v2a_launch_count(void);

/* ------------------------------------------------------------------------
 * Weight-gradient GEMM straight from channels-last operands (MN-major UMMA):
 *
 *   out[(unit, ci), co] += sum_pixels x_src(unit)[pixel + d(unit), 64*chunk(unit) + ci] * dy[pixel, co]
 *
 * replaces autograd's conv weight gradient for the observation encoder's
 * ResNet18 / keypoint convs   diffusion_policy/common/vision_nets.py:29-39
 *                             diffusion_policy/common/base_nets.py:183
 * `out` is accumulated with fp32 REDs (the pixel reduction is split between
 * CTAs): the caller zeroes it.  v2a_wgrad_scatter writes / adds the [(tap, ci)][co]
 * scratch into a parameter-gradient tensor laid out [co][ci][tap] (rows ld_dw apart).
 * ---------------------------------------------------------------------- */
#define V2A_WGRAD_MAX_UNITS 80
typedef struct v2a_wgrad_unit {
    int src;     /* index into desc.src */
    int d[4];    /* pixel offset of this tap in X0..X3 (zero padding outside) */
    int chunk;   /* 64-channel chunk of the source */
} v2a_wgrad_unit;

typedef struct v2a_wgrad_desc {
    v2a_igemm_src src[V2A_MAX_SRC];   /* x operands (bf16 hi/lo planes, channels-last) */
    int nsrc;
    v2a_wgrad_unit units[V2A_WGRAD_MAX_UNITS];   /* output rows = 64 * nunits, unit-major */
    int nunits;
    v2a_igemm_src dy;       /* output-gradient planes over the OUTPUT pixel grid dy.dims */
    int box_log2[4];        /* log2 of the 64-pixel reduction box along D0..D3 (sums to 6) */
    int cout;               /* multiple of 4 */
    int passes;             /* 3 = hi*hi + hi*lo + lo*hi; 1 = bf16 only */
    float* out;             /* fp32 [64*nunits][ld_out], accumulated */
    int ld_out;
    int x_fp16;             /* x AND dy planes are fp16 (hi, lo) pairs (one format per MMA) */
} v2a_wgrad_desc;
int v2a_wgrad_plan_create(const v2a_wgrad_desc* desc, void** plan_out);
int v2a_wgrad_plan_run(void* plan, void* stream);
int v2a_wgrad_plan_k_splits(void* plan);
void v2a_wgrad_plan_destroy(void* plan);
int v2a_wgrad_scatter(const float* wt, int ld, int cout, int cin, int ntaps, float* dw, int64_t ld_dw,
                      int accumulate, void* stream);

/* ------------------------------------------------------------------------
 * GroupNorm statistics + apply (HBM-bound elementwise).
 * replaces GroupNorm32 / SiLU / th.cat / F.interpolate(nearest) / the two
 * rearranges of Conv3d            guided_diffusion/nn.py:26-28,76,85
 *                                 guided_diffusion/unet.py:107-114,239-260,681
 * and Conv1dBlock's GroupNorm+Mish + FiLM
 *                                 diffusion_policy/model/conv1d_components.py:23-40
 *                                 diffusion_policy/model/conditional_unet1d.py:55-61
 * ---------------------------------------------------------------------- */
/* stats[(inst*C + c)*2 + {0,1}] += {sum, sumsq} over the pixels of instance */
int v2a_channel_stats(const float* x, int64_t instances, int64_t pixels_per_instance, int C,
                      double* stats, void* stream);

typedef struct v2a_prep_desc {
    const float* x0;        /* fp32 [P][C0] */
    const float* x1;        /* fp32 [P][C1] or NULL (channel concat) */
    int C0, C1;
    const double* stats0;   /* per (inst, channel) sums for x0 (NULL: no normalisation) */
    const double* stats1;
    int64_t pixels_per_inst;     /* pixels in one stats instance */
    int inst_per_group;          /* consecutive instances pooled into one GroupNorm sample */
    int groups;                  /* GroupNorm groups over C0+C1 */
    float eps;
    float* gn_scratch;      /* fp32 [samples*groups*2] mean/rstd table written by this call */
    const float* gamma;     /* [C0+C1] */
    const float* beta;
    int act;                /* 0 none, 1 SiLU, 2 Mish */
    const float* film;      /* policy FiLM [B][2*C]: out = scale*act(gn) + bias, or NULL */
    int64_t pixels_per_film;     /* pixels sharing one FiLM row */
    int mode;               /* 0 same grid, 1 nearest x2 upsample (H,W), 2 stride-2 phase split */
    int H, W;               /* input grid (mode 1/2); images = P / (H*W) */
    int64_t P;              /* input pixels */
    void* out_hi;           /* bf16 [P'][C0+C1] */
    void* out_lo;
    float* out_f32;         /* optional fp32 copy of the result (policy autograd) */
    void* raw_hi;           /* optional: hi/lo split of the un-normalised concat (1x1 skip operand) */
    void* raw_lo;
    int stats_rep0, stats_rep1;                    /* replica counts of stats0 / stats1 (0 = 1) */
    int64_t stats_rep_stride0, stats_rep_stride1;  /* doubles between replicas */
} v2a_prep_desc;
int v2a_prep(const v2a_prep_desc* d, void* stream);

/* ------------------------------------------------------------------------
 * Per-frame spatial self-attention, legacy head order.
 * replaces QKVAttentionLegacy.forward      guided_diffusion/unet.py:341-358
 * qkv fp32 [N][L][3*C] with channel = head*96 + {q:0..31, k:32..63, v:64..95};
 * writes a = softmax(q k^T / sqrt(32)) v as hl planes [N][L][C].
 * ---------------------------------------------------------------------- */
int v2a_attention(const float* qkv, int N, int L, int heads, void* out_hi, void* out_lo,
                  void* stream);

/* ------------------------------------------------------------------------
 * Small dense layers (batch <= a few hundred, fp32 CUDA cores).
 * y[b][o] = act_out( sum_i act_in(x[b][i]) * W[o][i] + bias[o] ) (+ add[b][o])
 * replaces time_embed / emb_layers / diffusion_step_encoder linears
 *   guided_diffusion/unet.py:204-210,482-486; conditional_unet1d.py:87-93
 * act codes: 0 none, 1 SiLU, 2 Mish
 * ---------------------------------------------------------------------- */
int v2a_linear(const float* x, int ldx, const float* W, const float* bias, const float* add,
               int ld_add, float* y, int ldy, int B, int IN, int OUT, int act_in, int act_out,
               void* stream);
/* sinusoidal embeddings: mode 0 = video [cos|sin], freq exp(-ln(1e4) i/half)
 *   (guided_diffusion/nn.py:171-189); mode 1 = policy [sin|cos], freq
 *   exp(-ln(1e4) i/(half-1)) (diffusion_policy/model/positional_embedding.py:5-17) */
int v2a_timestep_embedding(const int64_t* t, int B, int dim, int mode, float* out, void* stream);

/* ------------------------------------------------------------------------
 * Video UNet boundary layout kernels.  Unet_Libero.forward rearranges
 *   flowdiffusion/unet.py:216-222
 * ---------------------------------------------------------------------- */
/* First-conv operand: hl [B*F][H][W][64], k = tap*6 + c (c<3 frame rgb from `x`, c>=3
 * conditioning rgb from `cond`), 54 of 64 used.  Each source is addressed as
 * base[b*s[0] + f*s[1] + c*s[2] + h*W + w] (element strides), which covers both the
 * packed [B][(f c)+3][H][W] tensor (cond frame stride 0 = broadcast over frames) and
 * UNetModel's 5-D [B][6][F][H][W] input. */
int v2a_unet_input_pack(const float* x, const int64_t* x_strides, const float* cond,
                        const int64_t* cond_strides, int B, int F, int H, int W, void* out_hi,
                        void* out_lo, void* stream);
/* Second half of the out head's 3x3 conv (GN -> SiLU -> Conv3d(128, 3), guided_diffusion/unet.py:628-632): the
 * spatial conv runs as a 1x1 conv to 9*cout columns (P[pix][tap*cout + co], one igemm pass over the activation)
 * and this kernel gathers the nine taps: y[n,h,w][co] = bias[co] + sum_{kh,kw} P[n,h+kh-1,w+kw-1][(kh*3+kw)*cout+co],
 * zero outside the image.  cout <= 4. */
int v2a_stencil9(const float* P, int ldp, const float* bias, int N, int H, int W, int cout, float* y, int ldy,
                 void* stream);
/* y fp32 [B][F][H][W][ldy] (3 used) -> temporal Conv1d(3,3,k3)+bias ->
 * out[b*s[0] + f*s[1] + c*s[2] + h*W + w] */
int v2a_unet_output_head(const float* y, int ldy, const float* wt, const float* bt, int B, int F,
                         int H, int W, float* out, const int64_t* out_strides, void* stream);

/* ------------------------------------------------------------------------
 * Sampler steps.  replaces p_sample / ddim_sample elementwise chains
 *   flowdiffusion/goal_diffusion.py:484-497,561-580,601-641
 * coef (device, fp32[8]) = {sqrt_ac, sqrt_1m_ac, c1, c2, sigma, -, -, -} for DDPM,
 *        {sqrt_ac, sqrt_1m_ac, sqrt_recip_ac, sqrt_recipm1_ac, sqrt_ac_next, c, sigma, last} for DDIM
 * ---------------------------------------------------------------------- */
int v2a_ddpm_step(float* x, const float* v, const float* noise, const float* coef, int64_t n,
                  void* stream);
int v2a_ddim_step(float* x, const float* v, const float* noise, const float* coef, int64_t n,
                  void* stream);
/* Sampler update under classifier-free guidance (guidance_weight > 0, pred_v; goal_diffusion.py:503-514,536-548):
 * x and v hold the doubled batch [conditional | unconditional] (n_half elements each); the guided noise estimate is
 * mixed in noise space, the DDPM (ddim = 0) or eta-DDIM (ddim = 1) update applied, the result written to both
 * halves of x.  coef = the 8 floats of the plain step + coef[8] = guidance weight; for DDPM coef[6], coef[7] =
 * sqrt_recip_alphas_cumprod, sqrt_recipm1_alphas_cumprod of the step. */
int v2a_cfg_step(float* x, const float* v, const float* noise, const float* coef, int64_t n_half, int ddim,
                 void* stream);
/* out = clamp((x + 1) / 2, 0, 1)   goal_diffusion.py:598,650 */
int v2a_unnormalize_clamp(const float* x, float* out, int64_t n, void* stream);

/* out = sum over s < slices of x[s * stride + r * cols + c], summed in slice order, as (hi, lo) bf16 planes
 * [rows][cols]: the deterministic reduction of a split-K launch made with v2a_igemm_desc.split_stride */
int v2a_sum_slices_hl(const float* x, int slices, int64_t stride, int64_t rows, int cols, void* out_hi, void* out_lo,
                      void* stream);

/* fp32 -> (hi, lo) bf16 planes, optional [rows][cols] row-padding to ld */
int v2a_split_hl(const float* x, int64_t rows, int cols, int ld_out, void* out_hi, void* out_lo,
                 void* stream);

/* table-driven weight repack after an optimiser step: out[i] = map[i] ? src[map[i]-1] : 0 as bf16 hi/lo
 * planes and/or fp32.  Replaces the per-layer re-layout PyTorch does implicitly when cuDNN consumes
 * nn.Conv1d / nn.Linear weights (conditional_unet1d.py:36-44, conv1d_components.py:7-40). */
int v2a_gather_split(const float* src, const int32_t* map, int64_t n, void* out_hi, void* out_lo,
                     float* out_f32, void* stream);

/* ------------------------------------------------------------------------
 * Policy path (ConditionalUnet1D forward / backward).  One CTA per batch sample.
 * replaces Conv1dBlock's GroupNorm -> Mish (conv1d_components.py:23-40), the FiLM
 * modulation and residual add of ConditionalResidualBlock1D.forward
 * (conditional_unet1d.py:46-66) and their autograd backward.
 * ---------------------------------------------------------------------- */
typedef struct v2a_policy_gn_desc {
    int B, T, C, groups;
    float eps;
    const float* y;          /* conv output fp32 [B][T][C] */
    const float* gamma;      /* [C] */
    const float* beta;
    const float* film;       /* [B][ld_film]: scale = film[b][c], bias = film[b][C + c]; or NULL */
    int ld_film;
    const float* addend;     /* fwd: fp32 [B][T][ld_add] added to the result (identity residual) or NULL */
    int ld_add;
    float* out_f32;          /* fwd: [B][T][ld_out] or NULL */
    int ld_out;
    void* out_hi;            /* fwd: bf16 planes [B][T][ld_hl] or NULL */
    void* out_lo;
    int ld_hl;
    float* mean_rstd;        /* [B][groups][2]: written by fwd, read by bwd */
    /* backward */
    const float* dout;       /* grad wrt the block output, fp32 [B][T][ld_dout] */
    int ld_dout;
    void* dy_hi;             /* grad wrt y as bf16 planes [B][T][C] (operand of the data-gradient GEMM) */
    void* dy_lo;
    void* dyT_hi;            /* same, transposed [C][B*T] (operand of the weight-gradient GEMM) */
    void* dyT_lo;
    float* dy_f32;           /* optional fp32 copy */
    float* dbias;            /* [C] += sum_{b,t} dy   (conv bias grad) or NULL */
    float* dgamma;           /* [C] += ... */
    float* dbeta;
    float* dfilm;            /* [B][ld_dfilm] += (d scale | d bias) */
    int ld_dfilm;
    int64_t ld_T;            /* row pitch of dyT (>= B*T; the pad stays zero); 0 = B*T */
    float* partials;         /* bwd scratch [B][3][C] or NULL: per-sample bias/gamma/beta sums reduced by a second
                                launch instead of global atomics from every CTA (dfilm then needs no zeroing race:
                                one CTA owns each row) */
} v2a_policy_gn_desc;
int v2a_policy_gn_act_fwd(const v2a_policy_gn_desc* d, void* stream);
int v2a_policy_gn_act_bwd(const v2a_policy_gn_desc* d, void* stream);

/* transposed im2col of a bf16 hi/lo activation x [B][Tin][ld_x] (channels c_off..c_off+C):
 * out[(c*ntaps + k)][b*Tout + o] = x[b][stride*o + offsets[k]][c] (0 outside) — the K-major
 * operand of the weight-gradient GEMM dW[co][ci][k] = sum dy[b,o,co] x[b, s*o+off_k, ci] */
int v2a_policy_im2col_t(const void* x_hi, const void* x_lo, int ld_x, int c_off, int B, int Tin, int Tout,
                        int C, int ntaps, int stride, const int* offsets, void* out_hi, void* out_lo,
                        int64_t ld_out, void* stream);
/* dy fp32 [rows][ld] (C used) -> hl [rows][ld_hl] (zero padded), hl^T [C][rows], colsum[C] += */
int v2a_grad_prep(const float* dy, int64_t rows, int C, int ld, void* hi, void* lo, int ld_hl, void* t_hi,
                  void* t_lo, int64_t ld_T, float* colsum, void* stream);
/* In-place DDIM update (eta = 0) of the action trajectory x [rows][ldx] (C used) from the model output [rows][ldm]:
 *   epsilon prediction: x0 = clip((x - sqrt_1m_at * eps) / sqrt_at);  sample prediction: x0 = clip(model_out),
 *   eps = (x - sqrt_at * x0) / sqrt_1m_at;   x <- sqrt_aprev * x0 + coef_eps * eps
 * replaces diffusers' DDIMScheduler.step as predict_action drives it
 * (diffuser/diffusion_policy/diffusion_unet_image_policy.py:124-128) */
int v2a_policy_ddim_step(float* x, int ldx, const float* model_out, int ldm, int64_t rows, int C, float sqrt_1m_at,
                         float sqrt_at, float sqrt_aprev, float coef_eps, int pred_sample, int clip, void* stream);
/* dx = dy * act'(x), act 1 SiLU / 2 Mish; optional fp32 and hi/lo outputs */
int v2a_act_bwd(const float* x, const float* dy, float* dx, void* hi, void* lo, int64_t n, int act,
                void* stream);
/* dst[dst_off[r] + c] = src[r*ld + c], c < cols: rows of a concatenated gradient matrix back into the
 * per-module parameter-gradient windows (the FiLM cond_encoder linears run as ONE GEMM,
 * conditional_unet1d.py:36-40) */
int v2a_scatter_rows(const float* src, int ld, int64_t rows, int cols, const int64_t* dst_off, float* dst,
                     void* stream);
/* dst[r*ld_dst + c] (+)= src[r*ld_src + c], c < C (gradient fan-in of skip connections) */
int v2a_add_strided(float* dst, int ld_dst, const float* src, int ld_src, int64_t rows, int C, int accumulate,
                    void* stream);
/* optimiser tail of the train step (lb_online_trainer_v7.py:608-624): global grad-norm clip,
 * AdamW, EMA; out[0] += sum g^2 */
int v2a_grad_sumsq(const float* g, int64_t n, double* out, void* stream);
int v2a_adamw_ema_step(float* p, const float* g, float* m, float* v, float* ema, int64_t n,
                       const double* grad_sumsq, float max_norm, float lr, float beta1, float beta2, float eps,
                       float weight_decay, int step, float ema_decay, void* stream);

/* 64-bit content fingerprint of a list of fp32 tensors (no reference counterpart: it guards the packed-weight
 * caches of the engines against `param.data.copy_()` updates, which bypass autograd's version counter -- what
 * ema_pytorch.EMA.update() does to the EMA model the reference evaluates, lb_online_trainer_v7.py:624,1077).
 * items: device array of {const float* ptr; int64_t n; int64_t global_offset} chunks; out[0] = sum over all
 * elements of bits(x) * (2 * (global_offset + i) + 1) mod 2^64. */
int v2a_params_fingerprint(const void* items, int nitems, uint64_t* out, void* stream);

/* out[i] = src[map[i] - 1] (0 where map[i] == 0), split into planes of format plane_fmt (0 bf16, 1 fp16) */
int v2a_gather_split_fmt(const float* src, const int32_t* map, int64_t n, void* hi, void* lo, float* f32,
                         int plane_fmt, void* stream);

/* ------------------------------------------------------------------------
 * Observation encoder (2 x ResNet18-GroupNorm + SpatialSoftmax + Linear), forward and backward:
 * the HBM-bound kernels around the igemm / wgrad tensor-core contractions.
 * replaces  ResNet18Conv (BatchNorm -> GroupNorm(C/16))   diffusion_policy/common/vision_nets.py:9-39
 *                                                         diffusion_policy/model/multi_image_obs_encoder.py:67-74
 *           SpatialSoftmax.forward                        diffusion_policy/common/base_nets.py:234-285
 *           VisualCore's Linear                           diffusion_policy/common/vision_nets.py:113-143
 *           and their autograd backward.
 * ---------------------------------------------------------------------- */
/* (mean, rstd) per (instance, group) from the {sum, sumsq} a producing igemm accumulated */
int v2a_gn_finalize(const double* stats, int replicas, int64_t rep_stride, int instances, int C, int groups,
                    int64_t pixels_per_inst, float eps, float* mean_rstd, void* stream);
/* im2col of the 7x7 stride-2 pad-3 stem conv over x [B,3,H,W] (scale*x+shift folded in):
 * planes [B*(H/2)*(W/2)][192], k = c*49 + ky*7 + kx, zero padded */
int v2a_enc_stem_pack(const float* x, float scale, float shift, int B, int H, int W, void* out_hi, void* out_lo,
                      int plane_fmt, void* twin_hi, void* twin_lo, void* stream);
/* GroupNorm -> ReLU -> MaxPool2d(3, 2, 1): raw [B,H,W,C] -> out fp32 + planes [B,H/2,W/2,C] */
int v2a_enc_gn_relu_maxpool(const float* raw, const float* mean_rstd, int groups, const float* gamma,
                            const float* beta, int B, int H, int W, int C, float* out, void* out_hi, void* out_lo,
                            int plane_fmt, void* twin_hi, void* twin_lo, void* stream);
/* its backward up to the GroupNorm: g [B,H,W,C] from dpooled */
int v2a_enc_maxpool_relu_bwd(const float* raw, const float* mean_rstd, int groups, const float* gamma,
                             const float* beta, const float* pooled, const float* dpooled, int B, int H, int W, int C,
                             float* g, void* stream);

typedef struct v2a_enc_prep_desc {
    const float* xa;            /* fp32 [images*H*W][C] conv output */
    const float* mean_rstd_a;   /* [images][groups][2] */
    const float* gamma_a;
    const float* beta_a;
    const float* xb;            /* optional second normalised source (downsample path) */
    const float* mean_rstd_b;
    const float* gamma_b;
    const float* beta_b;
    const float* idn;           /* optional fp32 identity added after the GroupNorm (exclusive with xb) */
    int groups, C, H, W, images;
    int relu;
    int phase_split;            /* planes written as [img][py*2+px][H/2][W/2][C] (operand of a stride-2 conv) */
    int plane_fmt;              /* 0: bf16 (hi, lo) planes, 1: fp16 (hi, lo) planes */
    float* out_f32;             /* optional */
    void* out_hi;               /* optional */
    void* out_lo;
    void* out2_hi;              /* optional bf16 twin of the planes: the weight-gradient GEMM pairs x with the bf16 */
    void* out2_lo;              /* gradient planes and one tcgen05 MMA takes a single operand format */
} v2a_enc_prep_desc;
int v2a_enc_prep(const v2a_enc_prep_desc* d, void* stream);

typedef struct v2a_enc_gn_bwd_desc {
    const float* dout;          /* gradient wrt the block's post-activation output, fp32 [images*HW][C] */
    const float* outv;          /* mask_mode 1: saved post-ReLU output */
    const float* raw;           /* the GroupNorm's input (conv output) */
    const float* mean_rstd;
    const float* gamma;
    const float* beta;
    int mask_mode;              /* 0 none, 1 outv > 0, 2 GroupNorm(raw) > 0 (recomputed) */
    int groups, C, HW, images;
    float* sums;                /* scratch [images][C][2]: ZERO on entry (the caller clears it once per backward) */
    float* coef;                /* unused (kept for ABI stability) */
    void* d_hi;                 /* out: gradient wrt raw, bf16 planes */
    void* d_lo;
    float* g_out;               /* optional out: masked dout (identity / downsample branch gradient) */
    float* dgamma;              /* accumulated */
    float* dbeta;
} v2a_enc_gn_bwd_desc;
int v2a_enc_gn_bwd(const v2a_enc_gn_bwd_desc* d, void* stream);

/* phase-blocked data gradient of a stride-2 conv [img][H/2][W/2][4][C] (+ ds at phase 0) -> dx [img][H][W][C] */
int v2a_enc_unblock_add(const float* blocked, const float* ds, int images, int H, int W, int C, float* dx,
                        void* stream);
int v2a_enc_spatial_softmax_fwd(const float* logits, int ld, int B, int P, int K, float temperature,
                                const float* pos_x, const float* pos_y, float* att, float* kp, void* stream);
int v2a_enc_spatial_softmax_bwd(const float* att, const float* kp, const float* dkp, int B, int P, int K,
                                float temperature, const float* pos_x, const float* pos_y, void* d_hi, void* d_lo,
                                float* dbias, void* stream);
int v2a_enc_linear_bwd(const float* x, const float* dy, const float* W, int B, int IN, int OUT, float* dx, float* dW,
                       float* db, void* stream);

/* ------------------------------------------------------------------------
 * Replay-batch assembly on the device.
 * Replaces the per-step Python loop + torch.stack + host->device copy of
 * diffuser/datasets/env_img_replay_buffer.py:68-116 (`sample_random_batch_seq`),
 * diffuser/datasets/img_utils.py:27-37 (`img_np_toTensor`: uint8 HWC -> float CHW / 255)
 * and diffuser/libero/lb_online_trainer_v7.py:586 (`to_device_tp`).
 * ------------------------------------------------------------------------ */
/* frame_ptrs: DEVICE array of n device addresses, each a uint8 frame [H][W][3] (episodes own their
 * allocations); out fp32 [n][3][H][W] = float(u8) / 255, IEEE division = the reference's tensors bit for bit */
int v2a_replay_gather_images(const void* const* frame_ptrs, int n, int H, int W, float* out, void* stream);
/* row_ptrs: DEVICE array of B device addresses, each the first of T consecutive fp32 action rows [A];
 * out fp32 [B][T][A] */
int v2a_replay_gather_actions(const void* const* row_ptrs, int B, int T, int A, float* out, void* stream);

/* ------------------------------------------------------------------------
 * Task-token conditioning (row V13): task_attnpool = PerceiverResampler -> Linear -> mean over latents
 *   guided_diffusion/unet.py:491-494,671; guided_diffusion/imagen.py:197-211,254-372,1009-1017
 * Step-invariant (once per sample() call): fp32 CUDA-core kernels; the dense layers between them are v2a_linear.
 * Token rows are addressed as base + b * batch_stride + i * ld (i < n_tok), so a kernel can write straight into
 * a slice of a concatenated [B][tokens][D] buffer.
 * ------------------------------------------------------------------------ */
/* out = (act(x) (+ pos[i]) - mean) * rsqrt(var + eps) * gamma (+ beta); biased variance over D; act 0 none, 3 GELU(erf);
 * beta may be NULL (imagen's gain-only LayerNorm), pos [n_tok][D] may be NULL */
int v2a_pr_layernorm(const float* x, int64_t x_batch, int ldx, const float* pos, int B, int n_tok, int D, int act,
                     const float* gamma, const float* beta, float eps, float* out, int64_t out_batch, int ld_out,
                     void* stream);
/* per (row, head): out = x / max(||x||, 1e-12) * scale[dh]   (F.normalize * q_scale / k_scale, imagen.py:299-303) */
int v2a_pr_l2norm_scale(const float* x, int ldx, int rows, int heads, int dh, const float* scale, float* out,
                        int ld_out, void* stream);
/* out[b][i][h*dh..] = softmax_j(scale * q[b][i][h] . k[b][j][h]) @ v[b][j][h]; rows of q [B*nq], k / v [B*nk] */
int v2a_pr_attention(const float* q, int ldq, const float* k, int ldk, const float* v, int ldv, int B, int heads,
                     int dh, int nq, int nk, float scale, float* out, int ld_out, void* stream);
/* out[b][c] = mean_i x[b][i][c] */
int v2a_pr_token_mean(const float* x, int64_t x_batch, int ldx, int B, int n, int D, float* out, int ld_out,
                      void* stream);
/* out[b][i][c] = src[i][c] (learned latents broadcast over the batch) */
int v2a_pr_broadcast_rows(const float* src, int n, int D, int B, float* out, int64_t out_batch, int ld_out,
                          void* stream);

#ifdef __cplusplus
}
#endif
#endif
