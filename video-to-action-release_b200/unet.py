"""B200-native video UNet behind the reference's module surface.

Mirrors (names, constructor arguments, ``state_dict`` keys and shapes):
  * ``UNetModel``    flowdiffusion/flowdiffusion/guided_diffusion/guided_diffusion/unet.py:400-684
  * ``Unet_Libero``  flowdiffusion/flowdiffusion/unet.py:195-222
so reference checkpoints load with ``strict=True`` and ``ema_pytorch.EMA`` can
deep-copy the module.  The ``nn.Module`` tree below only HOLDS parameters (the
stock torch layer classes give the reference's initialisation for free); none
of their ``forward`` methods is ever called.  ``forward`` runs a planned list
of hand-written CUDA launches (``_UNetEngine``) through the C ABI:

  prep (GroupNorm-apply + SiLU + concat/upsample/phase-split + bf16 hi/lo split)
    -> tcgen05 implicit-GEMM 3x3 conv -> tcgen05 temporal conv (+1x1 skip as extra K,
       + bias + timestep-embedding add + residual + GroupNorm partial sums in the epilogue)
  per-frame attention: prep -> qkv GEMM -> fused softmax(QK)V -> proj GEMM (+residual, +sums)

Engines (TMA descriptors, packed weights, activation arena, CUDA graph) live in a
process-global registry keyed by module, never on the module itself.
"""
from __future__ import annotations

import math
import os
import weakref
from typing import Dict, List, Optional, Tuple

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import _lib, convs, ops, packing
from .ops import HL

# ---------------------------------------------------------------------------
# parameter holders (reference-identical names)
# ---------------------------------------------------------------------------


class Conv3d(nn.Module):
    """Pseudo-3D conv parameters: spatial Conv2d + temporal Conv1d (gd/nn.py:30-51)."""

    def __init__(self, dim, dim_out=None, kernel_size=3, stride=(1, 1, 1), padding=None):
        super().__init__()
        dim_out = dim_out or dim
        self.spatial_conv = nn.Conv2d(dim, dim_out, kernel_size, padding=kernel_size // 2, stride=tuple(stride[1:]))
        self.temporal_conv = nn.Conv1d(dim_out, dim_out, kernel_size) if kernel_size > 1 else None
        self.kernel_size = kernel_size
        self.stride = tuple(stride)
        if self.temporal_conv is not None:
            nn.init.dirac_(self.temporal_conv.weight.data)
            nn.init.zeros_(self.temporal_conv.bias.data)


class ResBlock(nn.Module):
    def __init__(self, channels, emb_channels, dropout, out_channels=None):
        super().__init__()
        self.channels = channels
        self.out_channels = out_channels or channels
        self.in_layers = nn.Sequential(nn.GroupNorm(32, channels), nn.SiLU(), Conv3d(channels, self.out_channels, 3))
        self.h_upd = self.x_upd = nn.Identity()
        self.emb_layers = nn.Sequential(nn.SiLU(), nn.Linear(emb_channels, self.out_channels))
        self.out_layers = nn.Sequential(nn.GroupNorm(32, self.out_channels), nn.SiLU(), nn.Dropout(p=dropout),
                                        Conv3d(self.out_channels, self.out_channels, 3))
        if self.out_channels == channels:
            self.skip_connection = nn.Identity()
        else:
            self.skip_connection = Conv3d(channels, self.out_channels, 1)


class AttentionBlock(nn.Module):
    def __init__(self, channels, num_heads=1, num_head_channels=-1):
        super().__init__()
        self.channels = channels
        if num_head_channels == -1:
            self.num_heads = num_heads
        else:
            assert channels % num_head_channels == 0, (
                f"q,k,v channels {channels} is not divisible by num_head_channels {num_head_channels}")
            self.num_heads = channels // num_head_channels
        self.norm = nn.GroupNorm(32, channels)
        self.qkv = nn.Conv1d(channels, channels * 3, 1)
        self.attention = nn.Identity()  # QKVAttentionLegacy has no parameters
        self.proj_out = nn.Conv1d(channels, channels, 1)


class Downsample(nn.Module):
    def __init__(self, channels, out_channels=None):
        super().__init__()
        self.channels = channels
        self.out_channels = out_channels or channels
        self.op = Conv3d(channels, self.out_channels, 3, stride=(1, 2, 2))


class Upsample(nn.Module):
    def __init__(self, channels, out_channels=None):
        super().__init__()
        self.channels = channels
        self.out_channels = out_channels or channels
        self.conv = Conv3d(channels, self.out_channels, 3)


class _GainLayerNorm(nn.Module):
    def __init__(self, dim):
        super().__init__()
        self.g = nn.Parameter(torch.ones(dim))


class PerceiverAttention(nn.Module):
    def __init__(self, *, dim, dim_head=64, heads=8, scale=8):
        super().__init__()
        self.scale, self.heads = scale, heads
        inner = dim_head * heads
        self.norm = nn.LayerNorm(dim)
        self.norm_latents = nn.LayerNorm(dim)
        self.to_q = nn.Linear(dim, inner, bias=False)
        self.to_kv = nn.Linear(dim, inner * 2, bias=False)
        self.q_scale = nn.Parameter(torch.ones(dim_head))
        self.k_scale = nn.Parameter(torch.ones(dim_head))
        self.to_out = nn.Sequential(nn.Linear(inner, dim, bias=False), nn.LayerNorm(dim))


class PerceiverResampler(nn.Module):
    def __init__(self, *, dim, depth, dim_head=64, heads=8, num_latents=64, num_latents_mean_pooled=4,
                 max_seq_len=512, ff_mult=4):
        super().__init__()
        self.pos_emb = nn.Embedding(max_seq_len, dim)
        self.latents = nn.Parameter(torch.randn(num_latents, dim))
        self.to_latents_from_mean_pooled_seq = None
        if num_latents_mean_pooled > 0:
            self.to_latents_from_mean_pooled_seq = nn.Sequential(
                _GainLayerNorm(dim), nn.Linear(dim, dim * num_latents_mean_pooled), nn.Identity())
        self.layers = nn.ModuleList([])
        hidden = int(dim * ff_mult)
        for _ in range(depth):
            ff = nn.Sequential(_GainLayerNorm(dim), nn.Linear(dim, hidden, bias=False), nn.GELU(),
                               _GainLayerNorm(hidden), nn.Linear(hidden, dim, bias=False))
            self.layers.append(nn.ModuleList([PerceiverAttention(dim=dim, dim_head=dim_head, heads=heads), ff]))


class TimestepEmbedSequential(nn.Sequential):
    pass


# ---------------------------------------------------------------------------
# engine registry
# ---------------------------------------------------------------------------
_ENGINES: "weakref.WeakKeyDictionary[nn.Module, Dict[tuple, _UNetEngine]]" = weakref.WeakKeyDictionary()


PRECISIONS = {"strict": 3, "fast": 1}


def _engine_for(model: "UNetModel", B: int, Fr: int, H: int, W: int, device) -> "_UNetEngine":
    per_model = _ENGINES.setdefault(model, {})
    device = torch.device(device)
    if device.type == "cuda" and device.index is None:
        device = torch.device("cuda", torch.cuda.current_device())
    precision = getattr(model, "precision", "strict")
    if precision not in PRECISIONS:
        raise ValueError(f"precision must be one of {sorted(PRECISIONS)}, got {precision!r}")
    key = (B, Fr, H, W, str(device), precision)
    eng = per_model.get(key)
    if eng is None:
        if len(per_model) >= 4:  # bound memory: drop the oldest shape
            per_model.pop(next(iter(per_model)))
        eng = _UNetEngine(model, B, Fr, H, W, device, passes=PRECISIONS[precision])
        per_model[key] = eng
    return eng


class UNetModel(nn.Module):
    """Parameter layout of the reference UNetModel (dims=3 pseudo-3D variant)."""

    def __init__(self, image_size, in_channels, model_channels, out_channels, num_res_blocks,
                 attention_resolutions, dropout=0, channel_mult=(1, 2, 4, 8), conv_resample=True, dims=2,
                 num_classes=None, task_tokens=True, task_token_channels=512, use_checkpoint=False,
                 use_fp16=False, num_heads=1, num_head_channels=-1, num_heads_upsample=-1,
                 use_scale_shift_norm=False, resblock_updown=False, use_new_attention_order=False):
        super().__init__()
        if dims != 3 or num_classes is not None or not task_tokens or use_scale_shift_norm or \
                resblock_updown or use_new_attention_order or not conv_resample or use_fp16 or in_channels != 6 \
                or out_channels != 3:
            raise NotImplementedError(
                "v2a_b200.UNetModel implements the configuration family the Libero video path uses "
                "(dims=3, task tokens, 6->3 channels, conv resample, legacy attention order)")
        if num_heads_upsample == -1:
            num_heads_upsample = num_heads
        self.image_size, self.in_channels, self.model_channels = image_size, in_channels, model_channels
        self.out_channels, self.num_res_blocks = out_channels, num_res_blocks
        self.attention_resolutions, self.dropout, self.channel_mult = attention_resolutions, dropout, channel_mult
        self.conv_resample, self.num_classes, self.task_tokens = conv_resample, num_classes, task_tokens
        self.use_checkpoint, self.dtype = use_checkpoint, torch.float32
        self.num_heads, self.num_head_channels, self.num_heads_upsample = num_heads, num_head_channels, num_heads_upsample
        # numerics class of the tensor-core contractions (a plain attribute, poked like `use_checkpoint`):
        #   "strict" (default) 3-pass bf16 split product, fp32-class: the 1e-3 parity bar of north_star
        #   "fast"   one bf16 product per contraction (~1e-2 on a forward): the class the reference itself ships
        #            on the GPU (fp16 autocast, scripts/train_libero_dp.py:25-26); opt-in, never the default
        self.precision = "strict"

        ted = model_channels * 4
        self.time_embed = nn.Sequential(nn.Linear(model_channels, ted), nn.SiLU(), nn.Linear(ted, ted))
        self.task_attnpool = nn.Sequential(PerceiverResampler(dim=task_token_channels, depth=2),
                                           nn.Linear(task_token_channels, ted))
        ch = input_ch = int(channel_mult[0] * model_channels)
        self.input_blocks = nn.ModuleList([TimestepEmbedSequential(Conv3d(in_channels, ch, 3))])
        chans = [ch]
        ds = 1
        attn = lambda c, nh: AttentionBlock(c, num_heads=nh, num_head_channels=num_head_channels)
        for level, mult in enumerate(channel_mult):
            for _ in range(num_res_blocks):
                layers = [ResBlock(ch, ted, dropout, out_channels=int(mult * model_channels))]
                ch = int(mult * model_channels)
                if ds in attention_resolutions:
                    layers.append(attn(ch, num_heads))
                self.input_blocks.append(TimestepEmbedSequential(*layers))
                chans.append(ch)
            if level != len(channel_mult) - 1:
                self.input_blocks.append(TimestepEmbedSequential(Downsample(ch, out_channels=ch)))
                chans.append(ch)
                ds *= 2
        self.middle_block = TimestepEmbedSequential(ResBlock(ch, ted, dropout), attn(ch, num_heads),
                                                    ResBlock(ch, ted, dropout))
        self.output_blocks = nn.ModuleList([])
        for level, mult in list(enumerate(channel_mult))[::-1]:
            for i in range(num_res_blocks + 1):
                ich = chans.pop()
                layers = [ResBlock(ch + ich, ted, dropout, out_channels=int(model_channels * mult))]
                ch = int(model_channels * mult)
                if ds in attention_resolutions:
                    layers.append(attn(ch, num_heads_upsample))
                if level and i == num_res_blocks:
                    layers.append(Upsample(ch, out_channels=ch))
                    ds //= 2
                self.output_blocks.append(TimestepEmbedSequential(*layers))
        self.out = nn.Sequential(nn.GroupNorm(32, ch), nn.SiLU(), Conv3d(input_ch, out_channels, 3))

    # -- reference signature: x [B, 6, F, H, W] -> [B, 3, F, H, W] (gd/unet.py:650-684)
    def forward(self, x, timesteps, y=None):
        assert y is not None, "must specify y if and only if the model is class-conditional"
        B, C, Fr, H, W = x.shape
        assert C == 6
        x = x.contiguous().float()
        eng = _engine_for(self, B, Fr, H, W, x.device)
        out = torch.empty(B, 3, Fr, H, W, device=x.device, dtype=torch.float32)
        HW = H * W
        eng.forward(x.data_ptr(), (6 * Fr * HW, HW, Fr * HW), x.data_ptr() + 4 * 3 * Fr * HW,
                    (6 * Fr * HW, HW, Fr * HW), timesteps, y, out, (3 * Fr * HW, HW, Fr * HW), keep=[x])
        return out

    # -- packed layout used by Unet_Libero / the sampler: x [B, 3F, H, W] + cond [B, 3, H, W]
    def forward_packed(self, x, cond, timesteps, y, out=None):
        B, C3, H, W = x.shape
        Fr = C3 // 3
        HW = H * W
        eng = _engine_for(self, B, Fr, H, W, x.device)
        if out is None:
            out = torch.empty(B, 3 * Fr, H, W, device=x.device, dtype=torch.float32)
        eng.forward(x.data_ptr(), (x.stride(0), 3 * HW, HW), cond.data_ptr(), (cond.stride(0), 0, HW),
                    timesteps, y, out, (3 * Fr * HW, 3 * HW, HW), keep=[x, cond])
        return out

    def engine(self, B, Fr, H, W, device) -> "_UNetEngine":
        return _engine_for(self, B, Fr, H, W, torch.device(device))


class Unet_Libero(nn.Module):
    """flowdiffusion/flowdiffusion/unet.py:195-222 (same constructor, same state_dict)."""

    def __init__(self):
        super().__init__()
        self.unet = UNetModel(image_size=(128, 128), in_channels=6, model_channels=128, out_channels=3,
                              num_res_blocks=2, attention_resolutions=(8, 16), dropout=0,
                              channel_mult=(1, 2, 3, 4, 5), conv_resample=True, dims=3, num_classes=None,
                              task_tokens=True, task_token_channels=512, use_checkpoint=False, use_fp16=False,
                              num_head_channels=32)

    def forward(self, x, t, task_embed=None, **kwargs):
        # x: [B, 3F + 3, H, W]; last 3 channels = conditioning frame broadcast over frames
        x = x.contiguous().float()
        return self.unet.forward_packed(x[:, :-3], x[:, -3:], t, task_embed)


# ---------------------------------------------------------------------------
# the engine
# ---------------------------------------------------------------------------
class _Pool:
    """Exact-size free lists so per-block scratch is reused across the forward."""

    def __init__(self, device):
        self.device = device
        self.free: Dict[int, List[torch.Tensor]] = {}
        self.total = 0

    def get(self, nbytes: int) -> torch.Tensor:
        lst = self.free.get(nbytes)
        if lst:
            return lst.pop()
        self.total += nbytes
        return torch.empty(nbytes, dtype=torch.uint8, device=self.device)

    def put(self, t: torch.Tensor) -> None:
        self.free.setdefault(t.numel(), []).append(t)


class _Act:
    """fp32 channels-last activation [N*H*W, C] + its per-(image, channel) GroupNorm sums."""

    def __init__(self, eng: "_UNetEngine", N, H, W, C):
        self.N, self.H, self.W, self.C = N, H, W, C
        self.rows = N * H * W
        self._raw_store = eng.pool.get(self.rows * C * 4)
        self.raw = self._raw_store.view(torch.float32).view(self.rows, C)
        # spread same-address atomics of the producer epilogue when many tiles share an instance
        reps = max(1, min(8, (H * W) // 512))
        self.stats = eng.take_stats(N, C, reps)
        self.refs = 1


class _UNetEngine:
    def __init__(self, model: UNetModel, B, Fr, H, W, device, passes: int = 3):
        if device.type != "cuda":
            raise RuntimeError("v2a_b200 UNet runs on CUDA only (no CPU fallback)")
        self.model_ref = weakref.ref(model)
        self.B, self.Fr, self.H, self.W, self.device = B, Fr, H, W, device
        self.N = B * Fr
        self.passes = passes
        self.pool = _Pool(device)
        self.steps: List = []          # callables, in launch order
        self.tags: List[str] = []      # what each step is (developer timing probes)
        self.igemms: List[ops.Igemm] = []
        self.packers: List = []        # (fn(model) -> fp32 tensor, HL destination)
        self.bias_packers: List = []
        self._stats_chunks: List[Tuple[int, int]] = []
        self._stats_total = 0
        self._stats_views: List = []
        self.graph: Optional[torch.cuda.CUDAGraph] = None
        self._wkey = None
        self._task_key = None
        mc = model.model_channels
        self.mc, self.ted = mc, mc * 4
        f32 = dict(dtype=torch.float32, device=device)
        self.t_buf = torch.zeros(B, dtype=torch.int64, device=device)
        self.temb = torch.empty(B, mc, **f32)
        self.temb_h = torch.empty(B, self.ted, **f32)
        self.task_emb = torch.zeros(B, self.ted, **f32)
        self.emb = torch.empty(B, self.ted, **f32)
        # static I/O for graph replay
        self.io = None
        self._build(model)
        self.stats_arena = torch.zeros(self._stats_total, dtype=torch.float64, device=device)
        for (off, n, shape), holder in zip(self._stats_chunks, self._stats_views):
            holder.append(self.stats_arena[off:off + n].view(*shape))
        self._finalize_plans()

    # ---- buffers -----------------------------------------------------------
    def take_stats(self, N, C, reps=1):
        """Reserve [reps, N, C, 2] float64 in the stats arena (bound after the plan is known)."""
        holder: List[torch.Tensor] = []
        n = reps * N * C * 2
        self._stats_chunks.append((self._stats_total, n, (reps, N, C, 2)))
        self._stats_views.append(holder)
        self._stats_total += n
        return holder

    def hl(self, rows, cols) -> Tuple[HL, List[torch.Tensor]]:
        a, b = self.pool.get(rows * cols * 2), self.pool.get(rows * cols * 2)
        return HL(a.view(torch.bfloat16).view(rows, cols), b.view(torch.bfloat16).view(rows, cols)), [a, b]

    def release(self, stores) -> None:
        for s in stores:
            self.pool.put(s)

    def free_act(self, a: _Act) -> None:
        a.refs -= 1
        if a.refs == 0:
            self.pool.put(a._raw_store)

    # ---- weights -----------------------------------------------------------
    def weight(self, fn, rows, cols) -> HL:
        hl = HL.empty(rows, cols, self.device)
        self.packers.append((fn, hl))
        return hl

    def vec(self, fn, n) -> torch.Tensor:
        v = torch.empty(n, dtype=torch.float32, device=self.device)
        self.bias_packers.append((fn, v))
        return v

    def refresh_weights(self, model: UNetModel, force=False) -> None:
        params = list(model.parameters())
        key = (tuple((p.data_ptr(), p._version) for p in params), packing.content_key(self, params))
        if not force and key == self._wkey:
            return
        with torch.no_grad():
            for fn, hl in self.packers:
                w = fn().detach().to(self.device, torch.float32)
                hi = w.to(torch.bfloat16)
                hl.hi.copy_(hi)
                hl.lo.copy_((w - hi.float()).to(torch.bfloat16))
            for fn, v in self.bias_packers:
                v.copy_(fn().detach().to(self.device, torch.float32).reshape(-1))
        self._wkey = key
        self._task_key = None

    # ---- deferred plan creation (stats views exist only after _build) -------
    def _step(self, fn, tag: str) -> None:
        self.steps.append(fn)
        self.tags.append(tag)

    def add_igemm(self, **kw):
        self._pending = getattr(self, "_pending", [])
        slot = [None]
        self._pending.append((slot, kw))
        self._step(lambda s=slot: s[0].run(), "igemm")
        return slot

    def add_igemm_dual(self, spatial_kw: dict, temporal_kw: dict):
        """One launch for a Conv3d: the spatial and the temporal implicit GEMM interleaved (ops.IgemmDual)."""
        self._pending = getattr(self, "_pending", [])
        slot = [None]
        self._pending.append((slot, ("dual", spatial_kw, temporal_kw)))
        self._step(lambda s=slot: s[0].run(), "igemm")
        return slot

    def add_prep(self, **kw):
        self._pending = getattr(self, "_pending", [])
        slot = [None]
        self._pending.append((slot, ("prep", kw)))
        self._step(lambda s=slot: s[0].run(), "prep" + ("_gn" if kw.get("stats0") is not None else f"_mode{kw.get('mode', 0)}"))
        return slot

    def _finalize_plans(self):
        def res(v):
            return v[0] if isinstance(v, list) and len(v) == 1 and isinstance(v[0], torch.Tensor) else v
        for slot, kw in self._pending:
            if isinstance(kw, tuple) and kw[0] == "dual":
                skw = {k: res(v) for k, v in kw[1].items()}
                tkw = {k: res(v) for k, v in kw[2].items()}
                g = ops.IgemmDual(dict(passes=self.passes, **skw), dict(passes=self.passes, **tkw), self.Fr)
                slot[0] = g
                self.igemms.append(g)
            elif isinstance(kw, tuple):
                args = {k: res(v) for k, v in kw[1].items()}
                slot[0] = ops.Prep(**args)
            else:
                args = {k: res(v) for k, v in kw.items()}
                g = ops.Igemm(passes=self.passes, fill_sms="block_n" not in args, **args)
                slot[0] = g
                self.igemms.append(g)
        self._pending = []
        self.flops = sum(g.flops for g in self.igemms)
        self.algo_flops = sum(g.algo_flops for g in self.igemms)

    # ---- building blocks ---------------------------------------------------
    def conv3d(self, m: Conv3d, a_hl: HL, cin, N, H, W, *, stride2=False, upsample=False, rowvec=None,
               residual=None, skip=None, extra_bias=None) -> _Act:
        """spatial 3x3 (TMA zero-padded taps) -> hl y -> temporal k3 (+skip K) -> fp32 raw + sums.
        H, W are the sizes of ``a_hl``'s grid; ``upsample``: the conv runs on the nearest-x2 upsampled grid
        (gd/unet.py:107-114) as four sub-pixel phases over the coarse operand, see convs.upsample3x3_phase."""
        B, Fr = self.B, self.Fr
        cout = m.spatial_conv.out_channels
        Ho, Wo = (H // 2, W // 2) if stride2 else ((2 * H, 2 * W) if upsample else (H, W))
        b_s = self.vec(lambda: m.spatial_conv.bias, cout)
        y_hl, y_st = self.hl(N * Ho * Wo, cout)
        if upsample:
            for py in range(2):
                for px in range(2):
                    prog = convs.upsample3x3_phase(cin, N, H, W, py, px)
                    w_p = self.weight(lambda py=py, px=px: convs.upsample3x3_phase_weight(m.spatial_conv.weight, py, px),
                                      cout, prog.ktot)
                    self.add_igemm(srcs=[(a_hl, cin, prog.src_dims[0])], taps=prog.taps, w=w_p,
                                   out_dims=prog.out_dims, cout=cout, out_hl=y_hl, bias=b_s,
                                   out_pix=convs.upsample3x3_out_pix(H, W, py, px), algo_flops_scale=9 / 4)
        spatial_kw = None
        if not upsample:
            prog = convs.spatial3x3_s2(cin, N, H, W) if stride2 else convs.spatial3x3(cin, N, H, W)
            w_s = self.weight(lambda: convs.spatial3x3_weight(m.spatial_conv.weight), cout, prog.ktot)
            spatial_kw = dict(srcs=[(a_hl, cin, prog.src_dims[0])], taps=prog.taps, w=w_s, out_dims=prog.out_dims,
                              cout=cout, out_hl=y_hl, bias=b_s)
        out = _Act(self, N, Ho, Wo, cout)
        HW = Ho * Wo
        srcs = [(y_hl, cout, (HW, Fr, B, 1))]
        if skip is not None:
            x_hl, cx, mskip = skip
            progt = convs.temporal3(cout, B, Fr, HW, skip_channels=cx)
            srcs.append((x_hl, cx, (HW, Fr, B, 1)))
            w_t = self.weight(lambda: convs.temporal3_weight(m.temporal_conv.weight, mskip.spatial_conv.weight),
                              cout, progt.ktot)
            b_t = self.vec(lambda: m.temporal_conv.bias + mskip.spatial_conv.bias, cout)
        else:
            progt = convs.temporal3(cout, B, Fr, HW)
            w_t = self.weight(lambda: convs.temporal3_weight(m.temporal_conv.weight), cout, progt.ktot)
            b_t = self.vec(lambda: m.temporal_conv.bias, cout)
        temporal_kw = dict(srcs=srcs, taps=progt.taps, w=w_t, out_dims=progt.out_dims, cout=cout, out_f32=out.raw,
                           bias=b_t, rowvec=rowvec, rowvec_mul=(0, 0, 1, 0), residual=residual, stats=out.stats,
                           stats_mul=(0, 1, Fr, 0))
        # Narrow layers (Cout <= 128): the temporal conv is bound by its epilogue, the spatial conv by its MMAs --
        # ONE dual launch interleaves them tile by tile so each hides the other (ops.IgemmDual); V2A_DUAL=0 = two
        # launches (A/B probe)
        if spatial_kw is not None and os.environ.get("V2A_DUAL", "1") != "0" and \
                ops.dual_conv3d_ok(spatial_kw["out_dims"], progt.out_dims, cout, self.passes, Fr):
            bn = ops.choose_block_n(cout)
            self.add_igemm_dual(dict(block_n=bn, **spatial_kw), dict(block_n=bn, **temporal_kw))
        else:
            if spatial_kw is not None:
                # Few output tiles and a long K (the deep levels at batch 1-2: 448 / 1792 rows, K up to 11.5 k): the
                # planes output rules out split-K, so the spatial GEMM goes to an fp32 scratch WITH split-K (widest N
                # tile: the activation tile is read once per split, not once per narrow N tile) and one small kernel
                # splits the finished sums into the planes.  V2A_SPLITK_SPATIAL=0: planes straight from the epilogue.
                rows_o = N * Ho * Wo
                bn0 = ops.choose_block_n(cout)
                tiles = -(-rows_o // 128) * -(-cout // bn0)
                if os.environ.get("V2A_SPLITK_SPATIAL", "1") != "0" and tiles * 2 <= 148 and \
                        spatial_kw["w"].hi.shape[1] // 64 >= 32 and cout % 16 == 0:
                    slices = 16                                   # the plan never splits further (csrc/igemm.cu)
                    sc_store = self.pool.get(slices * rows_o * cout * 4)
                    sc = sc_store.view(torch.float32).view(slices, rows_o, cout)
                    kw = dict(spatial_kw)
                    kw.pop("out_hl")
                    # every split stores its partial sums in its own slice and ONE kernel adds the slices in order
                    # and splits the result into the planes: no atomics, so repeated runs agree bit for bit
                    slot = self.add_igemm(out_f32=sc[0], block_n=bn0, split_stride=rows_o * cout, **kw)
                    lib = _lib.load()
                    self._step(lambda: _lib.check(lib.v2a_sum_slices_hl(sc.data_ptr(), slot[0].k_splits, rows_o * cout,
                                                                       rows_o, cout, y_hl.hi.data_ptr(),
                                                                       y_hl.lo.data_ptr(), ops._stream()), "sum_slices_hl"),
                               "sum_slices_hl")
                    self.pool.put(sc_store)
                else:
                    self.add_igemm(**spatial_kw)
            self.add_igemm(**temporal_kw)
        self.release(y_st)
        return out

    def gn_prep(self, parts: List[_Act], norm: nn.GroupNorm, *, per_frame, act, want_raw=False, mode=0):
        x0 = parts[0]
        x1 = parts[1] if len(parts) > 1 else None
        C = x0.C + (x1.C if x1 else 0)
        rows_out = x0.rows * (4 if mode == 1 else 1)
        a_hl, a_st = self.hl(rows_out, C)
        raw_hl, raw_st = (self.hl(rows_out, C) if want_raw else (None, []))
        gamma = self.vec(lambda: norm.weight, C)
        beta = self.vec(lambda: norm.bias, C)
        self.add_prep(x0=x0.raw, x1=None if x1 is None else x1.raw, stats0=x0.stats,
                      stats1=None if x1 is None else x1.stats, pixels_per_inst=x0.H * x0.W,
                      inst_per_group=1 if per_frame else self.Fr, groups=norm.num_groups, eps=norm.eps,
                      gamma=gamma, beta=beta, act=act, mode=mode, H=x0.H, W=x0.W, out_hl=a_hl, raw_hl=raw_hl)
        return a_hl, a_st, raw_hl, raw_st, C

    def res_block(self, m: ResBlock, parts: List[_Act], emb_slice) -> _Act:
        x0 = parts[0]
        N, H, W = x0.N, x0.H, x0.W
        has_skip = not isinstance(m.skip_connection, nn.Identity)
        a1, a1_st, x_hl, x_st, cin = self.gn_prep(parts, m.in_layers[0], per_frame=False, act=ops.ACT_SILU,
                                                  want_raw=has_skip)
        assert cin == m.channels
        h1 = self.conv3d(m.in_layers[2], a1, cin, N, H, W, rowvec=emb_slice)
        self.release(a1_st)
        a2, a2_st, _, _, _ = self.gn_prep([h1], m.out_layers[0], per_frame=False, act=ops.ACT_SILU)
        if has_skip:
            out = self.conv3d(m.out_layers[3], a2, m.out_channels, N, H, W, skip=(x_hl, cin, m.skip_connection))
        else:
            out = self.conv3d(m.out_layers[3], a2, m.out_channels, N, H, W, residual=x0.raw)
        self.release(a2_st)
        self.release(x_st)
        self.free_act(h1)
        for p in parts:
            self.free_act(p)
        return out

    def attn_block(self, m: AttentionBlock, x: _Act) -> _Act:
        N, H, W, C = x.N, x.H, x.W, x.C
        L = H * W
        assert C // m.num_heads == 32, "attention kernel is specialised for 32-channel heads"
        a, a_st, _, _, _ = self.gn_prep([x], m.norm, per_frame=True, act=ops.ACT_NONE)
        qkv_store = self.pool.get(x.rows * 3 * C * 4)
        qkv = qkv_store.view(torch.float32).view(x.rows, 3 * C)
        prog = convs.pointwise(C, (x.rows,))
        w_qkv = self.weight(lambda: convs.pointwise_weight(m.qkv.weight), 3 * C, prog.ktot)
        b_qkv = self.vec(lambda: m.qkv.bias, 3 * C)
        self.add_igemm(srcs=[(a, C, prog.src_dims[0])], taps=prog.taps, w=w_qkv, out_dims=prog.out_dims,
                       cout=3 * C, out_f32=qkv, bias=b_qkv)
        h_hl, h_st = self.hl(x.rows, C)
        heads = m.num_heads
        self._step(lambda: ops.attention(qkv, N, L, heads, h_hl), f"attention L{L}")
        out = _Act(self, N, H, W, C)
        prog2 = convs.pointwise(C, (L, N))
        w_p = self.weight(lambda: convs.pointwise_weight(m.proj_out.weight), C, prog2.ktot)
        b_p = self.vec(lambda: m.proj_out.bias, C)
        self.add_igemm(srcs=[(h_hl, C, prog2.src_dims[0])], taps=prog2.taps, w=w_p, out_dims=prog2.out_dims,
                       cout=C, out_f32=out.raw, bias=b_p, residual=x.raw, stats=out.stats, stats_mul=(0, 1, 0, 0))
        self.release(a_st)
        self.release(h_st)
        self.pool.put(qkv_store)
        self.free_act(x)
        return out

    def resample(self, conv: Conv3d, x: _Act, *, down: bool) -> _Act:
        s_hl, s_st = self.hl(x.rows, x.C)
        # down: stride-2 phase split of the operand; up: a plain hi/lo split of the COARSE tensor -- the upsampled
        # tensor (4x the rows) is never written, the conv's sub-pixel phases read the coarse one
        self.add_prep(x0=x.raw, mode=2 if down else 0, H=x.H, W=x.W, out_hl=s_hl)
        if down:
            out = self.conv3d(conv, s_hl, x.C, x.N, x.H, x.W, stride2=True)
        else:
            out = self.conv3d(conv, s_hl, x.C, x.N, x.H, x.W, upsample=True)
        self.release(s_st)
        self.free_act(x)
        return out

    def run_block(self, block: nn.Sequential, parts: List[_Act]) -> _Act:
        h = None
        for layer in block:
            if isinstance(layer, ResBlock):
                h = self.res_block(layer, parts if h is None else [h], self.emb_slices[id(layer)])
            elif isinstance(layer, AttentionBlock):
                h = self.attn_block(layer, h)
            elif isinstance(layer, Downsample):
                h = self.resample(layer.op, parts[0] if h is None else h, down=True)
            elif isinstance(layer, Upsample):
                h = self.resample(layer.conv, h, down=False)
            else:
                raise RuntimeError(f"unexpected layer {type(layer).__name__}")
        return h

    # ---- whole-network plan ------------------------------------------------
    def _build(self, model: UNetModel) -> None:
        B, Fr, H, W, N = self.B, self.Fr, self.H, self.W, self.N
        dev = self.device
        # timestep embedding MLP + all ResBlock emb_layers as ONE concatenated linear
        res_blocks = [m for m in model.modules() if isinstance(m, ResBlock)]
        tot = sum(m.out_channels for m in res_blocks)
        self.emb_all = torch.empty(B, tot, dtype=torch.float32, device=dev)
        self.emb_slices, off = {}, 0
        for m in res_blocks:
            self.emb_slices[id(m)] = self.emb_all[:, off:off + m.out_channels]
            off += m.out_channels
        te0, te2 = model.time_embed[0], model.time_embed[2]
        self.w_te0 = self.vec(lambda: te0.weight, te0.weight.numel()).view(te0.weight.shape)
        self.b_te0 = self.vec(lambda: te0.bias, te0.bias.numel())
        self.w_te2 = self.vec(lambda: te2.weight, te2.weight.numel()).view(te2.weight.shape)
        self.b_te2 = self.vec(lambda: te2.bias, te2.bias.numel())
        self.w_emb = self.vec(lambda: torch.cat([m.emb_layers[1].weight for m in res_blocks], 0),
                              tot * self.ted).view(tot, self.ted)
        self.b_emb = self.vec(lambda: torch.cat([m.emb_layers[1].bias for m in res_blocks], 0), tot)

        # SiLU(emb) once, not inside the wide linear: with act_in every one of its ~9 k output warps re-evaluated the
        # 16 x 512 activations (0.2 ms of a 0.29 ms emb_path)
        self.emb_act = torch.empty(B, self.ted, dtype=torch.float32, device=dev)
        emb_silu = ops.Prep(x0=self.emb, act=ops.ACT_SILU, out_f32=self.emb_act)

        def emb_path():
            ops.timestep_embedding(self.t_buf, self.mc, 0, self.temb)
            ops.linear(self.temb, self.w_te0, self.b_te0, self.temb_h, act_out=ops.ACT_SILU)
            ops.linear(self.temb_h, self.w_te2, self.b_te2, self.emb, add=self.task_emb)
            emb_silu.run()
            ops.linear(self.emb_act, self.w_emb, self.b_emb, self.emb_all)
        self._step(emb_path, "emb_path")

        # input conv: im2col'd 6-channel 3x3 (K = 54 -> one 64-wide chunk)
        conv0: Conv3d = model.input_blocks[0][0]
        c0 = conv0.spatial_conv.out_channels
        self.in_hl, _ = self.hl(N * H * W, 64)
        self._step(lambda: ops.unet_input_pack(self.io["x"], self.io["xs"], self.io["c"], self.io["cs"],
                                               B, Fr, H, W, self.in_hl), "input_pack")
        prog = convs.pointwise(64, (N * H * W,))
        w0 = self.weight(lambda: convs.input_conv_weight(conv0.spatial_conv.weight), c0, 64)
        b0 = self.vec(lambda: conv0.spatial_conv.bias, c0)
        y0, y0_st = self.hl(N * H * W, c0)
        self.add_igemm(srcs=[(self.in_hl, 64, prog.src_dims[0])], taps=prog.taps, w=w0, out_dims=prog.out_dims,
                       cout=c0, out_hl=y0, bias=b0)
        h = _Act(self, N, H, W, c0)
        progt = convs.temporal3(c0, B, Fr, H * W)
        wt0 = self.weight(lambda: convs.temporal3_weight(conv0.temporal_conv.weight), c0, progt.ktot)
        bt0 = self.vec(lambda: conv0.temporal_conv.bias, c0)
        self.add_igemm(srcs=[(y0, c0, progt.src_dims[0])], taps=progt.taps, w=wt0, out_dims=progt.out_dims,
                       cout=c0, out_f32=h.raw, bias=bt0, stats=h.stats, stats_mul=(0, 1, Fr, 0))
        self.release(y0_st)

        hs = [h]
        h.refs += 1
        for block in list(model.input_blocks)[1:]:
            h = self.run_block(block, [h])
            hs.append(h)
            h.refs += 1
        h = self.run_block(model.middle_block, [h])
        for block in model.output_blocks:
            h = self.run_block(block, [h, hs.pop()])
        # out head: GN -> SiLU -> 3x3 conv to 3 channels -> temporal conv + layout kernel.  With 3 output channels the
        # nine-tap implicit GEMM re-reads the 128-channel operand nine times to feed a 16-wide N tile (1.29 ms at
        # 9.8 TFLOP/s in round 1); instead ONE 1x1 GEMM to 27 columns P[pix][tap*3 + co], then a 9-point gather.
        a, a_st, _, _, C = self.gn_prep([h], model.out[0], per_frame=False, act=ops.ACT_SILU)
        convo: Conv3d = model.out[2]
        prog = convs.pointwise(C, (N * H * W,))
        wo = self.weight(lambda: convs.taps_as_columns_weight(convo.spatial_conv.weight), 27, prog.ktot)
        bo = self.vec(lambda: convo.spatial_conv.bias, 3)
        p_out = self.pool.get(N * H * W * 32 * 4)
        self.p_out = p_out.view(torch.float32).view(N * H * W, 32)
        self.y_out = torch.empty(N * H * W, 4, dtype=torch.float32, device=dev)
        self.add_igemm(srcs=[(a, C, prog.src_dims[0])], taps=prog.taps, w=wo, out_dims=prog.out_dims, cout=27,
                       ldc=32, out_f32=self.p_out, block_n=32, algo_flops_scale=1.0)
        self._step(lambda: ops.stencil9(self.p_out, bo, N, H, W, 3, self.y_out), "stencil9")
        self.wt_out = self.vec(lambda: convo.temporal_conv.weight, 27)
        self.bt_out = self.vec(lambda: convo.temporal_conv.bias, 3)
        self._step(lambda: ops.unet_output_head(self.y_out, 4, self.wt_out, self.bt_out, B, Fr, H, W,
                                                self.io["o"], self.io["os"]), "output_head")
        self.release(a_st)
        self.pool.put(p_out)
        self.free_act(h)

    # ---- conditioning (step-invariant): PerceiverResampler on the text tokens ---
    def set_task_embed(self, model: UNetModel, y: torch.Tensor) -> None:
        # The cache is keyed on the token VALUES: the caller's tensor is usually a temporary (video_model.py:68-70
        # builds a fresh embedding per call) and the caching allocator hands the same address back for the next
        # prompt, so (data_ptr, _version) would silently condition a new task on the previous one.  B*L*512 floats
        # compared once per sample() call.
        y = y.detach().to(self.device, torch.float32).contiguous()
        prev = self._task_key
        if prev is not None and prev.shape == y.shape and torch.equal(prev, y):
            return
        with torch.no_grad(), torch.autocast("cuda", enabled=False):
            _task_pool_cuda(model.task_attnpool, y, self.task_emb)
        self._task_key = y.clone()

    # ---- execution ---------------------------------------------------------
    def _launch_all(self) -> None:
        self.stats_arena.zero_()
        for s in self.steps:
            s()

    def forward(self, x_ptr, xs, c_ptr, cs, timesteps, y, out, os_, keep=()):
        model = self.model_ref()
        self.refresh_weights(model)
        self.set_task_embed(model, y)
        t = timesteps if torch.is_tensor(timesteps) else torch.tensor([timesteps], device=self.device)
        self.t_buf.copy_(t.to(self.device, torch.int64).expand(self.B))
        self.io = dict(x=x_ptr, xs=xs, c=c_ptr, cs=cs, o=out, os=os_)
        self._launch_all()

    # static-buffer interface for the sampler (CUDA-graph capturable)
    def bind_static(self, x, cond, out):
        HW = self.H * self.W
        self.io = dict(x=x.data_ptr(), xs=(x.stride(0), 3 * HW, HW), c=cond.data_ptr(), cs=(cond.stride(0), 0, HW),
                       o=out, os=(3 * self.Fr * HW, 3 * HW, HW))
        self._static = (x, cond, out)

    def run_static(self):
        self._launch_all()


_POOL_MAX_ROWS = 512


def _task_pool_cuda(seq: nn.Sequential, y: torch.Tensor, out: torch.Tensor) -> None:
    """`task_attnpool(y).mean(1)` on the v2a kernels (`csrc/perceiver.cu` + `v2a_linear`): out [B, D].

    Same operation order as the reference (gd/imagen.py:254-372) except the final `Linear(D, D)` and the mean over
    the latents, which commute (the layer is affine): the mean is taken first, on 68x fewer rows.
    torch only allocates the scratch buffers here.
    """
    pr: PerceiverResampler = seq[0]
    B, n, D = y.shape
    dev = y.device
    f32 = dict(device=dev, dtype=torch.float32)
    n_pool = 0
    n_lat = pr.latents.shape[0]
    mp = pr.to_latents_from_mean_pooled_seq
    if mp is not None:
        n_pool = mp[1].weight.shape[0] // D
    NL = n_pool + n_lat
    # samples are independent: keep every dense layer within `_POOL_MAX_ROWS` token rows per launch (the small-batch
    # `v2a_linear` covers 8 rows x 64 row-blocks per pass), i.e. 6 samples of 12 + 68 tokens at a time
    per_call = max(1, _POOL_MAX_ROWS // (n + NL))
    if B > per_call:
        for b0 in range(0, B, per_call):
            _task_pool_cuda(seq, y[b0:b0 + per_call], out[b0:b0 + per_call])
        return
    lat = torch.empty(B, NL, D, **f32)
    ops.pr_broadcast_rows(pr.latents.detach(), lat[:, n_pool:])
    if mp is not None:
        ymean = torch.empty(B, D, **f32)
        ops.pr_token_mean(y, ymean)
        yn = torch.empty(B, 1, D, **f32)
        ops.pr_layernorm(ymean.unsqueeze(1), mp[0].g.detach(), None, yn)
        # [B, n_pool * D] written as the first n_pool latent rows of every sample (row stride NL * D)
        ops.linear(yn.reshape(B, D), mp[1].weight.detach(), mp[1].bias.detach(), lat.reshape(B, NL * D))
    pos = pr.pos_emb.weight.detach()
    for attn, ff in pr.layers:
        h = attn.heads
        inner = attn.to_q.weight.shape[0]
        kvin = torch.empty(B, n + NL, D, **f32)                 # cat([norm(x + pos), norm_latents(lat)], dim=1)
        ops.pr_layernorm(y, attn.norm.weight.detach(), attn.norm.bias.detach(), kvin[:, :n], pos=pos)
        ops.pr_layernorm(lat, attn.norm_latents.weight.detach(), attn.norm_latents.bias.detach(), kvin[:, n:])
        kv = torch.empty(B * (n + NL), 2 * inner, **f32)
        ops.linear(kvin.reshape(B * (n + NL), D), attn.to_kv.weight.detach(), None, kv)
        ln_rows = torch.empty(B * NL, D, **f32)                 # the latent rows of kvin, contiguous for to_q
        ops.pr_layernorm(lat, attn.norm_latents.weight.detach(), attn.norm_latents.bias.detach(),
                         ln_rows.reshape(B, NL, D))
        q = torch.empty(B * NL, inner, **f32)
        ops.linear(ln_rows, attn.to_q.weight.detach(), None, q)
        qn = torch.empty_like(q)
        ops.pr_l2norm_scale(q, h, attn.q_scale.detach(), qn)
        kn = torch.empty(B * (n + NL), inner, **f32)
        ops.pr_l2norm_scale(kv[:, :inner], h, attn.k_scale.detach(), kn)
        o = torch.empty(B * NL, inner, **f32)
        ops.pr_attention(qn, kn, kv[:, inner:], B, h, float(attn.scale), o)
        po = torch.empty(B * NL, D, **f32)
        ops.linear(o, attn.to_out[0].weight.detach(), None, po)
        lat2 = torch.empty(B, NL, D, **f32)
        ops.pr_layernorm(po.reshape(B, NL, D), attn.to_out[1].weight.detach(), attn.to_out[1].bias.detach(), lat2)
        ops.add_rows_(lat2.reshape(B * NL, D), lat.reshape(B * NL, D))        # lat = to_out(...) + lat
        lat = lat2
        hidden = ff[1].weight.shape[0]
        g1 = torch.empty(B, NL, D, **f32)
        ops.pr_layernorm(lat, ff[0].g.detach(), None, g1)
        hdn = torch.empty(B * NL, hidden, **f32)
        ops.linear(g1.reshape(B * NL, D), ff[1].weight.detach(), None, hdn)
        g2 = torch.empty(B, NL, hidden, **f32)
        ops.pr_layernorm(hdn.reshape(B, NL, hidden), ff[3].g.detach(), None, g2, act=ops.ACT_GELU)
        lat3 = torch.empty(B * NL, D, **f32)
        ops.linear(g2.reshape(B * NL, hidden), ff[4].weight.detach(), None, lat3, add=lat.reshape(B * NL, D))
        lat = lat3.reshape(B, NL, D)
    lmean = torch.empty(B, D, **f32)
    ops.pr_token_mean(lat, lmean)
    assert out.shape == (B, seq[1].weight.shape[0]) and out.stride(1) == 1      # Linear(D, time_embed_dim)
    ops.linear(lmean, seq[1].weight.detach(), seq[1].bias.detach(), out)
