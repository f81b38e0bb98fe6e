"""B200-native ``ConditionalUnet1D`` (forward AND backward) behind the reference's module surface.

Mirrors diffuser/diffusion_policy/model/conditional_unet1d.py:69-246 (same constructor,
``state_dict`` keys/shapes, ``forward(sample, timestep, local_cond=None, global_cond=None)``).
The module tree only holds parameters.  ``forward`` runs a planned list of CUDA launches and
is wrapped in a ``torch.autograd.Function`` whose backward runs a second planned list:

  forward, per ConditionalResidualBlock1D
      conv k5 (tcgen05 implicit GEMM, TMA zero-padded taps, concat inputs as two tap groups)
      -> GroupNorm+Mish+FiLM (one CTA per sample)           -> conv k5 -> GroupNorm+Mish
      -> + residual (identity: fused add; 1x1 conv: GEMM with the block result as residual)
  backward
      GroupNorm/Mish/FiLM backward (one CTA per sample) emits dy (hi/lo), dy^T and the
      bias / gamma / beta / FiLM gradients in one pass
      data gradient   = implicit GEMM over dy with flipped weights (fan-in accumulates in its epilogue)
      weight gradient = GEMM dy^T x im2col^T(x), written straight into the parameter-gradient slab
  all FiLM linears run as ONE concatenated GEMM (forward, dgrad and wgrad).

Activations are channels-last [B, T, C] — exactly the layout ``sample`` arrives in, so the
reference's two rearranges (conditional_unet1d.py:189,245) disappear.
"""
from __future__ import annotations

import ctypes as C
import os
import weakref
from typing import Callable, Dict, List, Optional, Union

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import _lib, convs, ops, packing
from .ops import HL
from .packing import PackedParams

# ---------------------------------------------------------------------------
# parameter holders (reference-identical names)
# ---------------------------------------------------------------------------


class SinusoidalPosEmb(nn.Module):
    def __init__(self, dim):
        super().__init__()
        self.dim = dim


class Conv1dBlock(nn.Module):
    """Conv1d --> GroupNorm --> Mish (conv1d_components.py:23-40)."""

    def __init__(self, inp_channels, out_channels, kernel_size, n_groups=8):
        super().__init__()
        self.block = nn.Sequential(nn.Conv1d(inp_channels, out_channels, kernel_size, padding=kernel_size // 2),
                                   nn.GroupNorm(n_groups, out_channels), nn.Mish())


class Downsample1d(nn.Module):
    def __init__(self, dim):
        super().__init__()
        self.conv = nn.Conv1d(dim, dim, 3, 2, 1)


class Upsample1d(nn.Module):
    def __init__(self, dim):
        super().__init__()
        self.conv = nn.ConvTranspose1d(dim, dim, 4, 2, 1)


class ConditionalResidualBlock1D(nn.Module):
    def __init__(self, in_channels, out_channels, cond_dim, kernel_size=3, n_groups=8, cond_predict_scale=False):
        super().__init__()
        self.blocks = nn.ModuleList([Conv1dBlock(in_channels, out_channels, kernel_size, n_groups=n_groups),
                                     Conv1dBlock(out_channels, out_channels, kernel_size, n_groups=n_groups)])
        cond_channels = out_channels * 2 if cond_predict_scale else out_channels
        self.cond_predict_scale = cond_predict_scale
        self.out_channels = out_channels
        self.in_channels = in_channels
        self.cond_encoder = nn.Sequential(nn.Mish(), nn.Linear(cond_dim, cond_channels), nn.Identity())
        self.residual_conv = nn.Conv1d(in_channels, out_channels, 1) if in_channels != out_channels else nn.Identity()


_ENGINES: "weakref.WeakKeyDictionary[nn.Module, Dict[tuple, _PolicyEngine]]" = weakref.WeakKeyDictionary()
# modules whose parameter gradients stay in the engine's flat slab (train_step.PolicyTrainStep) instead of
# being handed to autograd as per-parameter tensors
_SLAB_GRADS: "weakref.WeakKeyDictionary[nn.Module, bool]" = weakref.WeakKeyDictionary()
_LAST_ENGINE: "weakref.WeakKeyDictionary[nn.Module, _PolicyEngine]" = weakref.WeakKeyDictionary()


def set_slab_grads(model: nn.Module, on: bool = True) -> None:
    """Leave parameter gradients in ``engine.gslab`` (flat, parameters() order) and return None to autograd."""
    _SLAB_GRADS[model] = bool(on)


def last_engine(model: nn.Module) -> "Optional[_PolicyEngine]":
    """The engine the most recent forward() of ``model`` ran on (holds that step's gradient slab)."""
    return _LAST_ENGINE.get(model)


def invalidate_weights(model: nn.Module) -> None:
    """Parameters were updated in place through raw pointers (fused optimiser): repack on next forward."""
    for eng in _ENGINES.get(model, {}).values():
        eng._wkey = None


class ConditionalUnet1D(nn.Module):
    def __init__(self, input_dim, local_cond_dim=None, global_cond_dim=None, diffusion_step_embed_dim=256,
                 down_dims=[256, 512, 1024], kernel_size=3, n_groups=8, cond_predict_scale=False,
                 cond_unet1d_config={}):
        super().__init__()
        if local_cond_dim is not None or not cond_predict_scale or cond_unet1d_config.get("no_down_up", False):
            raise NotImplementedError("v2a_b200.ConditionalUnet1D covers the Libero policy configuration "
                                      "(global conditioning, FiLM scale+bias, down/up-sampling)")
        all_dims = [input_dim] + list(down_dims)
        start_dim = down_dims[0]
        dsed = diffusion_step_embed_dim
        self.no_down_up = False
        self.input_dim, self.dsed, self.kernel_size, self.n_groups = input_dim, dsed, kernel_size, n_groups
        self.global_cond_dim = global_cond_dim or 0
        cond_dim = dsed + self.global_cond_dim
        in_out = list(zip(all_dims[:-1], all_dims[1:]))
        mk = lambda i, o: ConditionalResidualBlock1D(i, o, cond_dim=cond_dim, kernel_size=kernel_size,
                                                     n_groups=n_groups, cond_predict_scale=cond_predict_scale)
        mid_dim = all_dims[-1]
        # registration order follows the reference (mid, step encoder, up, down, final) so state_dict order matches
        self.mid_modules = nn.ModuleList([mk(mid_dim, mid_dim), mk(mid_dim, mid_dim)])
        down_modules = nn.ModuleList([])
        for ind, (dim_in, dim_out) in enumerate(in_out):
            is_last = ind >= (len(in_out) - 1)
            down_modules.append(nn.ModuleList([mk(dim_in, dim_out), mk(dim_out, dim_out),
                                               Downsample1d(dim_out) if not is_last else nn.Identity()]))
        up_modules = nn.ModuleList([])
        for ind, (dim_in, dim_out) in enumerate(reversed(in_out[1:])):
            is_last = ind >= (len(in_out) - 1)
            up_modules.append(nn.ModuleList([mk(dim_out * 2, dim_in), mk(dim_in, dim_in),
                                             Upsample1d(dim_in) if not is_last else nn.Identity()]))
        final_conv = nn.Sequential(Conv1dBlock(start_dim, start_dim, kernel_size=kernel_size),
                                   nn.Conv1d(start_dim, input_dim, 1))
        self.diffusion_step_encoder = nn.Sequential(SinusoidalPosEmb(dsed), nn.Linear(dsed, dsed * 4), nn.Mish(),
                                                    nn.Linear(dsed * 4, dsed))
        self.local_cond_encoder = None
        self.up_modules = up_modules
        self.down_modules = down_modules
        self.final_conv = final_conv

    def forward(self, sample: torch.Tensor, timestep: Union[torch.Tensor, float, int], local_cond=None,
                global_cond=None, **kwargs):
        """sample [B, T, input_dim]; timestep int / 0-d / [B]; global_cond [B, global_cond_dim] -> [B, T, input_dim]."""
        assert local_cond is None, "local conditioning is not part of the Libero policy path"
        if not sample.is_cuda:
            raise RuntimeError("v2a_b200.ConditionalUnet1D runs on CUDA only (no CPU fallback)")
        B, T, D = sample.shape
        assert D == self.input_dim
        t = timestep
        if not torch.is_tensor(t):
            t = torch.tensor([t], dtype=torch.long, device=sample.device)
        elif t.dim() == 0:
            t = t[None].to(sample.device)
        t = t.expand(B).to(torch.int64)
        gc = global_cond if global_cond is not None else sample.new_zeros(B, 0)
        eng = _policy_engine(self, B, T, sample.device)
        eng.training_call = packing.training_call(eng.params)
        _LAST_ENGINE[self] = eng
        return _UNet1DFunction.apply(self, eng, sample, t, gc, *list(self.parameters()))


def _policy_engine(model, B, T, device) -> "_PolicyEngine":
    per = _ENGINES.setdefault(model, {})
    device = torch.device(device)
    if device.index is None:
        device = torch.device("cuda", torch.cuda.current_device())
    key = (B, T, str(device))
    eng = per.get(key)
    if eng is None:
        if len(per) >= 4:
            per.pop(next(iter(per)))
        eng = _PolicyEngine(model, B, T, device)
        per[key] = eng
    return eng


class _UNet1DFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, model, eng, sample, t, gc, *params):
        with torch.autocast("cuda", enabled=False):
            out = eng.forward(sample.detach().float(), t, gc.detach().float())
        ctx.eng = eng
        ctx.slab = _SLAB_GRADS.get(model, False)
        ctx.token = eng.fwd_token
        ctx.in_dtypes = (sample.dtype, gc.dtype)
        return out

    @staticmethod
    def backward(ctx, grad_out):
        eng: _PolicyEngine = ctx.eng
        if ctx.token != eng.fwd_token:
            raise RuntimeError("v2a_b200.ConditionalUnet1D: backward() after another forward() on the same "
                               "module/shape — activations live in static buffers (one forward per backward)")
        with torch.autocast("cuda", enabled=False):
            d_sample, d_gc, pgrads = eng.backward(grad_out.float(), clone_param_grads=not ctx.slab)
        hook = getattr(eng, "on_backward_done", None)
        if hook is not None:       # one-shot: PolicyTrainStep starts this slab's all-reduce under the encoders' backward
            eng.on_backward_done = None
            hook()
        return (None, None, d_sample.to(ctx.in_dtypes[0]), None, d_gc.to(ctx.in_dtypes[1]), *pgrads)


# ---------------------------------------------------------------------------
# engine
# ---------------------------------------------------------------------------
def _c16(n):
    return -(-n // 16) * 16


def _c64(n):
    return -(-n // 64) * 64


class _Ref:
    """Column window [off, off+C) of a row-pitched fp32 matrix t [rows, ld]."""

    def __init__(self, t: torch.Tensor, off: int, Cc: int):
        self.t, self.off, self.C = t, off, Cc
        self.ld = t.stride(0)

    @property
    def win(self):
        return self.t[:, self.off:self.off + self.C]

    @property
    def ptr(self):
        return self.t.data_ptr() + 4 * self.off


class _Node:
    """Activation [B*T, C] stored with row pitch ld (zero padded) as fp32 and/or hi/lo planes."""

    def __init__(self, B, T, Cc, ld=None, f32=None, hl=None):
        self.B, self.T, self.C = B, T, Cc
        self.ld = ld or Cc
        self.f32, self.hl = f32, hl
        self.grad: Optional[_Ref] = None


class _Steps(list):
    """Launch list that remembers what each entry is (developer timing probes read .tags)."""

    def __init__(self):
        super().__init__()
        self.tags: List[str] = []
        self.lanes: List[int] = []
        self.lane = 0            # lane given to steps added while it is set (1 = side stream)

    def add(self, tag: str, fn: Callable, lane: Optional[int] = None) -> None:
        self.tags.append(tag)
        self.lanes.append(self.lane if lane is None else lane)
        list.append(self, fn)

    def append(self, fn: Callable) -> None:
        self.add(getattr(fn, "__name__", "step"), fn)

    def tail(self, start: int) -> "_Steps":
        """The same launches from index ``start`` on (tags and lanes kept)."""
        out = _Steps()
        for fn, tag, lane in list(zip(self, self.tags, self.lanes))[start:]:
            out.add(tag, fn, lane)
        return out


class _PolicyEngine(PackedParams):
    def __init__(self, model: ConditionalUnet1D, B, T, device):
        self.model_ref = weakref.ref(model)
        self.B, self.T, self.device = B, T, device
        self.passes = int(os.environ.get("V2A_PASSES", "3"))
        self.lib = _lib.load()
        self.fwd: _Steps = _Steps()
        self.bwd: _Steps = _Steps()
        self._init_packing()
        self.keep = []
        self._graphs: Dict[str, object] = {}
        self._gn_partials = None
        self._side = None if os.environ.get("V2A_NO_SIDE_STREAM", "0") == "1" else torch.cuda.Stream(device=device)
        self.fwd_token = 0
        self._wkey = None
        self.igemms: List[ops.Igemm] = []
        self.params = list(model.parameters())
        tot = sum(p.numel() for p in self.params)
        self.gslab = torch.zeros(tot, dtype=torch.float32, device=device)
        self.pgrad: Dict[int, torch.Tensor] = {}
        off = 0
        for p in self.params:
            self.pgrad[id(p)] = self.gslab[off:off + p.numel()].view(p.shape)
            off += p.numel()
        self._dcat: Dict[tuple, torch.Tensor] = {}
        self._zero_each_bwd: List[torch.Tensor] = [self.gslab]
        self._mn_wgrad = os.environ.get("V2A_POLICY_WGRAD", "mn") == "mn"   # "im2col": the transposed-im2col path
        self._wg_arenas: List[torch.Tensor] = []
        self._wg_used = 0
        self.wgrads: List[ops.Wgrad] = []
        self._build(model)
        self._trace_packers()

    # ---- buffers / weights ---------------------------------------------------
    def zeros(self, rows, cols):
        return torch.zeros(rows, cols, dtype=torch.float32, device=self.device)

    def hlz(self, rows, cols) -> HL:
        return HL(torch.zeros(rows, cols, dtype=torch.bfloat16, device=self.device),
                  torch.zeros(rows, cols, dtype=torch.bfloat16, device=self.device))

    def _needs_dyT(self, k: int, ins) -> bool:
        """True when a conv over `ins` falls back to the transposed-im2col weight gradient (which wants dy^T);
        the MN-major weight gradient reads dy itself, so GroupNorm backward need not write the transpose."""
        return not (self._mn_wgrad and all(k * ops.nchunks(n.ld) <= _lib.V2A_WGRAD_MAX_UNITS for n in ins))

    def _wg_scratch(self, rows: int, cols: int) -> torch.Tensor:
        """[rows, cols] fp32 carved from arenas that are cleared once per backward (the MN-major weight-gradient
        GEMM accumulates its split pixel reduction with REDs)."""
        n = rows * cols
        if not self._wg_arenas or self._wg_used + n > self._wg_arenas[-1].numel():
            self._wg_arenas.append(torch.zeros(max(n, 32 << 20), dtype=torch.float32, device=self.device))
            self._zero_each_bwd.append(self._wg_arenas[-1])
            self._wg_used = 0
        t = self._wg_arenas[-1][self._wg_used:self._wg_used + n].view(rows, cols)
        self._wg_used += n
        return t

    # ---- launch wrappers -------------------------------------------------------
    def igemm(self, steps, **kw):
        g = ops.Igemm(passes=self.passes, **kw)
        self.igemms.append(g)
        steps.add(f"igemm M{g.rows} N{g.cout} K{g.ktot}", g.run)

    def gn(self, steps, backward: bool, **kw):
        d = _lib.PolicyGnDesc()
        if backward:   # shared scratch for the per-sample bias / gamma / beta sums (launches are stream ordered)
            if self._gn_partials is None:
                self._gn_partials = torch.zeros(self.B * 3 * 2048, dtype=torch.float32, device=self.device)
            assert kw["B"] * 3 * kw["C"] <= self._gn_partials.numel()
            kw["partials"] = self._gn_partials
        for k, v in kw.items():
            setattr(d, k, v.data_ptr() if isinstance(v, torch.Tensor) else v)
        self.keep.append((d, kw))
        fn = self.lib.v2a_policy_gn_act_bwd if backward else self.lib.v2a_policy_gn_act_fwd
        steps.add("gn_bwd" if backward else "gn_fwd", lambda: _lib.check(fn(C.byref(d), ops._stream()), "policy_gn_act"))

    def im2col_t(self, steps, src: HL, ld, Bn, Tin, Tout, Cc, offsets, stride) -> HL:
        """-> planes [Cc*len(offsets), pad64(Bn*Tout)] (zero padded K), the B operand of a weight-gradient GEMM."""
        kpad = _c64(Bn * Tout)
        out = self.hlz(Cc * len(offsets), kpad)
        arr = (C.c_int * len(offsets))(*offsets)
        self.keep.append(arr)
        steps.add(f"im2col_t C{Cc} taps{len(offsets)} K{kpad}", lambda: _lib.check(self.lib.v2a_policy_im2col_t(
            src.hi.data_ptr(), src.lo.data_ptr(), ld, 0, Bn, Tin, Tout, Cc, len(offsets), stride, arr,
            out.hi.data_ptr(), out.lo.data_ptr(), kpad, ops._stream()), "im2col_t"))
        return out

    def grad_prep(self, steps, ref: _Ref, rows, *, want_hl=True, colsum=None):
        """fp32 gradient window -> (planes [rows, c16(C)], transposed planes [C, pad64(rows)]) (+ column sums)."""
        ldh = _c16(ref.C)
        hl = self.hlz(rows, ldh) if want_hl else None
        kpad = _c64(rows)
        tr = self.hlz(ref.C, kpad)
        steps.add(f"grad_prep rows{rows} C{ref.C}", lambda: _lib.check(self.lib.v2a_grad_prep(
            ref.ptr, rows, ref.C, ref.ld, None if hl is None else hl.hi.data_ptr(),
            None if hl is None else hl.lo.data_ptr(), ldh, tr.hi.data_ptr(), tr.lo.data_ptr(), kpad,
            None if colsum is None else colsum.data_ptr(), ops._stream()), "grad_prep"))
        return hl, ldh, tr

    def add_into(self, steps, dst: _Ref, src: _Ref, rows, accumulate=True):
        steps.add("add_strided", lambda: _lib.check(self.lib.v2a_add_strided(dst.ptr, dst.ld, src.ptr, src.ld, rows, dst.C,
                                                                1 if accumulate else 0, ops._stream()), "add"))

    def act_bwd(self, steps, x, dy, dx, n, act):
        steps.add("act_bwd", lambda: _lib.check(self.lib.v2a_act_bwd(x.data_ptr(), dy.data_ptr(), dx.data_ptr(), None, None,
                                                            n, act, ops._stream()), "act_bwd"))

    # ---- generic stride-1 conv over [B, T, C] (k taps; inputs may be a 2-way channel concat) -------------
    def conv_fwd(self, wfn, bfn, cout, k, pad, ins: List[_Node], *, out_f32=None, out_hl=None, residual=None):
        Bn, T = ins[0].B, ins[0].T
        prog = convs.conv1d_cat([n.ld for n in ins], Bn, T, k, pad)

        def packed(wfn=wfn, ins=ins):
            w = wfn()
            parts, off = [], 0
            for n in ins:
                for j in range(k):
                    wj = w[:, off:off + n.C, j]
                    parts.append(F.pad(wj, (0, n.ld - n.C)) if n.ld != n.C else wj)
                off += n.C
            return ops.pack_weight_taps(parts)
        w = self.weight(packed, cout, prog.ktot)
        b = self.vec(bfn, cout) if bfn is not None else None
        self.igemm(self.fwd, srcs=[(n.hl, n.ld, d) for n, d in zip(ins, prog.src_dims)], taps=prog.taps, w=w,
                   out_dims=prog.out_dims, cout=cout, out_f32=out_f32, out_hl=out_hl, bias=b, residual=residual)

    def wgrad(self, steps, dyT: HL, m_rows, col: HL, ncols, target: torch.Tensor, col_off):
        """target[:, col_off:col_off+ncols] = dyT [m_rows, K] x col [ncols, K]^T (fp32, straight into the slab)."""
        kpad = dyT.hi.shape[1]
        assert col.hi.shape[1] == kpad
        prog = convs.pointwise(kpad, (m_rows,))
        ntot = target.shape[1]
        if ncols % 16 == 0 and col_off % 4 == 0 and ntot % 4 == 0:
            self.igemm(steps, srcs=[(dyT, kpad, prog.src_dims[0])], taps=prog.taps, w=col, out_dims=prog.out_dims,
                       cout=ncols, out_f32=target[:, col_off:col_off + ncols])
        else:  # narrow / unaligned parameter (7-channel input or output): padded scratch, then a strided copy
            tmp = self.zeros(m_rows, _c16(ncols))
            self.igemm(steps, srcs=[(dyT, kpad, prog.src_dims[0])], taps=prog.taps, w=col, out_dims=prog.out_dims,
                       cout=ncols, out_f32=tmp)
            steps.add("copy_narrow", lambda: target[:, col_off:col_off + ncols].copy_(tmp[:, :ncols]))

    def conv_bwd(self, steps, wfn, wgrad_target, k, pad, ins: List[_Node], dy: HL, ld_dy, dyT: HL, cout,
                 dx_residual: Optional[_Ref] = None):
        """Weight gradient into `wgrad_target` [cout, cin_tot*k] and input gradient (fan-in aware)."""
        Bn, T = ins[0].B, ins[0].T
        rows = Bn * T
        cin_tot = sum(n.C for n in ins)
        off = 0
        steps.lane = 1
        for n in ins:
            units = [(0, (j - pad, 0, 0, 0), ch) for j in range(k) for ch in range(ops.nchunks(n.ld))]
            if self._mn_wgrad and len(units) <= _lib.V2A_WGRAD_MAX_UNITS:
                # MN-major tcgen05 weight gradient straight from the channels-last planes of x and dy: no
                # transposed im2col copy (44 launches, 1.8 ms of the 3.5 ms backward before) and no dy^T operand
                sc = self._wg_scratch(64 * len(units), ld_dy)
                wg = ops.Wgrad(srcs=[(n.hl, n.ld, (T, Bn, 1, 1))], units=units, dy=dy, dy_channels=ld_dy,
                               dy_dims=(T, Bn, 1, 1), cout=ld_dy, out=sc, passes=ops.wgrad_passes(self.passes, T * Bn))
                self.wgrads.append(wg)
                steps.add(f"wgrad M{64 * len(units)} N{ld_dy}", wg.run)
                win = wgrad_target[:, off * k:(off + n.C) * k]
                steps.add("wgrad_scatter", lambda sc=sc, n=n, win=win: ops.wgrad_scatter(
                    sc, cout, n.C, k, win, ld_dw=wgrad_target.stride(0), accumulate=False))
            else:
                col = self.im2col_t(steps, n.hl, n.ld, Bn, T, T, n.C, [j - pad for j in range(k)], 1)
                self.wgrad(steps, dyT, cout, col, n.C * k, wgrad_target, off * k)
            off += n.C
        steps.lane = 0
        progd = convs.conv1d(ld_dy, Bn, T, k, pad)

        def packed_d(wfn=wfn):
            w = wfn()  # [cout, cin_tot, k]; dgrad taps ascend in offset: j = k-1 .. 0
            parts = [F.pad(w[:, :, j].t(), (0, ld_dy - cout)) if ld_dy != cout else w[:, :, j].t()
                     for j in reversed(range(k))]
            return ops.pack_weight_taps(parts)
        wd = self.weight(packed_d, cin_tot, progd.ktot, bwd=True)
        src = [(dy, ld_dy, progd.src_dims[0])]
        if len(ins) == 1:
            n = ins[0]
            if n.grad is None:
                n.grad = _Ref(self.zeros(rows, _c16(n.C)), 0, n.C)
                res = dx_residual
            else:
                if dx_residual is not None:
                    self.add_into(steps, n.grad, dx_residual, rows)
                res = n.grad
            self.igemm(steps, srcs=src, taps=progd.taps, w=wd, out_dims=progd.out_dims, cout=cin_tot,
                       out_f32=n.grad.t[:, n.grad.off:n.grad.off + _c16(n.C)] if n.grad.C % 16 else n.grad.win,
                       residual=None if res is None else (res.t[:, res.off:res.off + _c16(res.C)] if res.C % 16 else res.win))
            return
        assert dx_residual is None
        key = tuple(id(n) for n in ins)
        dcat = self._dcat.get(key)
        if dcat is None:  # first gradient into this concat: parts become windows of one buffer
            dcat = self.zeros(rows, cin_tot)
            self._dcat[key] = dcat
            self.igemm(steps, srcs=src, taps=progd.taps, w=wd, out_dims=progd.out_dims, cout=cin_tot, out_f32=dcat)
            off = 0
            for n in ins:
                part = _Ref(dcat, off, n.C)
                if n.grad is None:
                    n.grad = part
                else:
                    self.add_into(steps, n.grad, part, rows)
                off += n.C
        else:         # further gradients accumulate in place in the GEMM epilogue
            assert all(n.grad is not None and n.grad.t is dcat for n in ins)
            self.igemm(steps, srcs=src, taps=progd.taps, w=wd, out_dims=progd.out_dims, cout=cin_tot, out_f32=dcat,
                       residual=dcat)

    # ---- network plan ------------------------------------------------------------
    def _build(self, model: ConditionalUnet1D):
        Bn, T, dev = self.B, self.T, self.device
        dsed, gcd = model.dsed, model.global_cond_dim
        cd = dsed + gcd
        G = model.n_groups
        g = lambda p: self.pgrad[id(p)]
        self.tape: List[Callable] = []       # backward planners, run in reverse
        # ---------------- conditioning path ----------------
        self.t_buf = torch.zeros(Bn, dtype=torch.int64, device=dev)
        temb = self.zeros(Bn, dsed)
        temb_hl = self.hlz(Bn, dsed)
        self.fwd.add("timestep_embedding", lambda: ops.timestep_embedding(self.t_buf, dsed, 1, temb))
        self.fwd.add("split_hl", lambda: _lib.check(self.lib.v2a_split_hl(temb.data_ptr(), Bn, dsed, dsed, temb_hl.hi.data_ptr(),
                                                                temb_hl.lo.data_ptr(), ops._stream()), "split"))
        lin1, lin3 = model.diffusion_step_encoder[1], model.diffusion_step_encoder[3]
        n_temb = _Node(Bn, 1, dsed, hl=temb_hl)
        a1 = self.zeros(Bn, 4 * dsed)
        self.conv_fwd(lambda: lin1.weight.unsqueeze(-1), lambda: lin1.bias, 4 * dsed, 1, 0, [n_temb], out_f32=a1)
        m1_hl = self.hlz(Bn, 4 * dsed)
        p1 = ops.Prep(x0=a1, act=ops.ACT_MISH, out_hl=m1_hl)
        self.fwd.add("prep_mish", p1.run)
        n_m1 = _Node(Bn, 1, 4 * dsed, hl=m1_hl)
        self.gf = self.zeros(Bn, cd)
        self.conv_fwd(lambda: lin3.weight.unsqueeze(-1), lambda: lin3.bias, dsed, 1, 0, [n_m1],
                      out_f32=self.gf[:, :dsed])
        # everything up to here depends on the timestep and the weights only (not on the observation): the single-graph
        # predict_action keeps gf[:, :dsed] of its 8 DDIM timesteps in a table and starts each forward after this point
        self.n_temb_steps = len(self.fwd)
        self.temb_out = self.gf[:, :dsed]
        self.gc_in = self.gf[:, dsed:]
        mgf_hl = self.hlz(Bn, cd)
        p2 = ops.Prep(x0=self.gf, act=ops.ACT_MISH, out_hl=mgf_hl)
        self.fwd.add("prep_mish", p2.run)
        n_mgf = _Node(Bn, 1, cd, hl=mgf_hl)
        blocks = [m for m in model.modules() if isinstance(m, ConditionalResidualBlock1D)]
        ftot = sum(2 * m.out_channels for m in blocks)
        self.film = self.zeros(Bn, ftot)
        self.dfilm = self.zeros(Bn, ftot)
        self._zero_each_bwd.append(self.dfilm)
        film_off, o = {}, 0
        for m in blocks:
            film_off[id(m)] = o
            o += 2 * m.out_channels
        wfilm = lambda: torch.cat([m.cond_encoder[1].weight for m in blocks], 0).unsqueeze(-1)
        self.conv_fwd(wfilm, lambda: torch.cat([m.cond_encoder[1].bias for m in blocks], 0), ftot, 1, 0, [n_mgf],
                      out_f32=self.film)

        # ---------------- residual blocks ----------------
        def res_block(m: ConditionalResidualBlock1D, ins: List[_Node]) -> _Node:
            Tn, rows, Co = ins[0].T, Bn * ins[0].T, m.out_channels
            k, pad = m.blocks[0].block[0].kernel_size[0], m.blocks[0].block[0].padding[0]
            c1, gn1 = m.blocks[0].block[0], m.blocks[0].block[1]
            c2, gn2 = m.blocks[1].block[0], m.blocks[1].block[1]
            fo = film_off[id(m)]
            y1, y2 = self.zeros(rows, Co), self.zeros(rows, Co)
            mr1, mr2 = self.zeros(Bn, 2 * G), self.zeros(Bn, 2 * G)
            hf = _Node(Bn, Tn, Co, hl=self.hlz(rows, Co))
            out = _Node(Bn, Tn, Co, f32=self.zeros(rows, Co), hl=self.hlz(rows, Co))
            ga1, be1 = self.vec(lambda: gn1.weight, Co), self.vec(lambda: gn1.bias, Co)
            ga2, be2 = self.vec(lambda: gn2.weight, Co), self.vec(lambda: gn2.bias, Co)
            film_ptr = self.film.data_ptr() + 4 * fo
            has_res = not isinstance(m.residual_conv, nn.Identity)
            # The 1x1 residual conv reads only the block input, so it runs on the side lane beside conv1 -> GN -> conv2
            # and the second GroupNorm adds its result (same two fp32 values added once, whichever kernel does it): 4
            # dependent launches per block instead of 5.  `predict_action` (every launch is latency): 4.6 -> 4.2 ms per
            # call; training batch: forward 1.244 -> 1.155 ms (gpurun_out/r2c28_*).  V2A_SIDE_RES=0: the serial form.
            side_res = has_res and os.environ.get("V2A_SIDE_RES", "1") != "0"
            if side_res:
                rc = m.residual_conv
                r_side = self.zeros(rows, Co)
                self.fwd.lane = 1
                self.conv_fwd(lambda: rc.weight, lambda: rc.bias, Co, 1, 0, ins, out_f32=r_side)
                self.fwd.lane = 0
            self.conv_fwd(lambda: c1.weight, lambda: c1.bias, Co, k, pad, ins, out_f32=y1)
            self.gn(self.fwd, False, B=Bn, T=Tn, C=Co, groups=G, eps=gn1.eps, y=y1, gamma=ga1, beta=be1,
                    film=film_ptr, ld_film=ftot, out_hi=hf.hl.hi, out_lo=hf.hl.lo, ld_hl=Co, mean_rstd=mr1)
            self.conv_fwd(lambda: c2.weight, lambda: c2.bias, Co, k, pad, [hf], out_f32=y2)
            if side_res:
                self.fwd.lane = 2
                self.gn(self.fwd, False, B=Bn, T=Tn, C=Co, groups=G, eps=gn2.eps, y=y2, gamma=ga2, beta=be2,
                        addend=r_side, ld_add=Co, out_f32=out.f32, ld_out=Co, out_hi=out.hl.hi,
                        out_lo=out.hl.lo, ld_hl=Co, mean_rstd=mr2)
                self.fwd.lane = 0
            elif has_res:
                h2 = self.zeros(rows, Co)
                self.gn(self.fwd, False, B=Bn, T=Tn, C=Co, groups=G, eps=gn2.eps, y=y2, gamma=ga2, beta=be2,
                        out_f32=h2, ld_out=Co, mean_rstd=mr2)
                rc = m.residual_conv
                self.conv_fwd(lambda: rc.weight, lambda: rc.bias, Co, 1, 0, ins, out_f32=out.f32, out_hl=out.hl,
                              residual=h2)
            else:
                x = ins[0]
                self.gn(self.fwd, False, B=Bn, T=Tn, C=Co, groups=G, eps=gn2.eps, y=y2, gamma=ga2, beta=be2,
                        addend=x.f32, ld_add=x.f32.stride(0), out_f32=out.f32, ld_out=Co, out_hi=out.hl.hi,
                        out_lo=out.hl.lo, ld_hl=Co, mean_rstd=mr2)

            def plan_bwd():
                st = self.bwd
                dO = out.grad
                assert dO is not None
                kp = _c64(rows)
                none = HL(None, None)
                dy2 = self.hlz(rows, Co)
                dy2T = self.hlz(Co, kp) if self._needs_dyT(k, [hf]) else none
                self.gn(st, True, B=Bn, T=Tn, C=Co, groups=G, eps=gn2.eps, y=y2, gamma=ga2, beta=be2, mean_rstd=mr2,
                        dout=dO.ptr, ld_dout=dO.ld, dy_hi=dy2.hi, dy_lo=dy2.lo, dyT_hi=dy2T.hi, dyT_lo=dy2T.lo,
                        ld_T=kp, dbias=g(c2.bias), dgamma=g(gn2.weight), dbeta=g(gn2.bias))
                self.conv_bwd(st, lambda: c2.weight, g(c2.weight).view(Co, -1), k, pad, [hf], dy2, Co, dy2T, Co)
                dy1 = self.hlz(rows, Co)
                dy1T = self.hlz(Co, kp) if self._needs_dyT(k, ins) else none
                self.gn(st, True, B=Bn, T=Tn, C=Co, groups=G, eps=gn1.eps, y=y1, gamma=ga1, beta=be1, mean_rstd=mr1,
                        film=film_ptr, ld_film=ftot, dout=hf.grad.ptr, ld_dout=hf.grad.ld, dy_hi=dy1.hi, dy_lo=dy1.lo,
                        dyT_hi=dy1T.hi, dyT_lo=dy1T.lo, ld_T=kp, dbias=g(c1.bias), dgamma=g(gn1.weight),
                        dbeta=g(gn1.bias), dfilm=self.dfilm.data_ptr() + 4 * fo, ld_dfilm=ftot)
                if has_res:
                    rc = m.residual_conv
                    dOh, ldh, dOT = self.grad_prep(st, dO, rows, colsum=g(rc.bias))
                    self.conv_bwd(st, lambda: rc.weight, g(rc.weight).view(Co, -1), 1, 0, ins, dOh, ldh, dOT, Co)
                    self.conv_bwd(st, lambda: c1.weight, g(c1.weight).view(Co, -1), k, pad, ins, dy1, Co, dy1T, Co)
                else:
                    self.conv_bwd(st, lambda: c1.weight, g(c1.weight).view(Co, -1), k, pad, ins, dy1, Co, dy1T, Co,
                                  dx_residual=dO)
            self.tape.append(plan_bwd)
            return out

        def down(mod: Downsample1d, x: _Node) -> _Node:
            Cc, Tn = x.C, x.T
            conv = mod.conv
            out = _Node(Bn, Tn // 2, Cc, f32=self.zeros(Bn * Tn // 2, Cc), hl=self.hlz(Bn * Tn // 2, Cc))
            prog = convs.down1d(Cc, Bn, Tn)
            w = self.weight(lambda: convs.conv1d_weight(conv.weight), Cc, prog.ktot)
            b = self.vec(lambda: conv.bias, Cc)
            self.igemm(self.fwd, srcs=[(x.hl, Cc, prog.src_dims[0])], taps=prog.taps, w=w, out_dims=prog.out_dims,
                       cout=Cc, out_f32=out.f32, out_hl=out.hl, bias=b)

            def plan_bwd():
                st, rows_o = self.bwd, Bn * Tn // 2
                dOh, ldh, dOT = self.grad_prep(st, out.grad, rows_o, colsum=g(conv.bias))
                st.lane = 1
                col = self.im2col_t(st, x.hl, x.ld, Bn, Tn, Tn // 2, Cc, [-1, 0, 1], 2)
                self.wgrad(st, dOT, Cc, col, Cc * 3, g(conv.weight).view(Cc, -1), 0)
                st.lane = 0
                progd = convs.down1d_dgrad(ldh, Bn, Tn)
                wd = self.weight(lambda: convs.down1d_dgrad_weight(conv.weight), 2 * Cc, progd.ktot, bwd=True)
                tmp = self.zeros(rows_o, 2 * Cc)  # == [B*T, C] memory
                self.igemm(st, srcs=[(dOh, ldh, progd.src_dims[0])], taps=progd.taps, w=wd, out_dims=progd.out_dims,
                           cout=2 * Cc, out_f32=tmp)
                part = _Ref(tmp.view(Bn * Tn, Cc), 0, Cc)
                if x.grad is None:
                    x.grad = part
                else:
                    self.add_into(st, x.grad, part, Bn * Tn)
            self.tape.append(plan_bwd)
            return out

        def up(mod: Upsample1d, x: _Node) -> _Node:
            Cc, Tn = x.C, x.T
            conv = mod.conv
            rows = Bn * Tn
            o32, ohl = self.zeros(rows, 2 * Cc), self.hlz(rows, 2 * Cc)
            out = _Node(Bn, 2 * Tn, Cc, f32=o32.view(2 * rows, Cc),
                        hl=HL(ohl.hi.view(2 * rows, Cc), ohl.lo.view(2 * rows, Cc)))
            prog = convs.up1d(Cc, Bn, Tn)
            w = self.weight(lambda: convs.up1d_weight(conv.weight), 2 * Cc, prog.ktot)
            b = self.vec(lambda: torch.cat([conv.bias, conv.bias]), 2 * Cc)
            self.igemm(self.fwd, srcs=[(x.hl, Cc, prog.src_dims[0])], taps=prog.taps, w=w, out_dims=prog.out_dims,
                       cout=2 * Cc, out_f32=o32, out_hl=ohl, bias=b)

            def plan_bwd():
                st = self.bwd
                dOh, ldh, _ = self.grad_prep(st, out.grad, 2 * rows, colsum=g(conv.bias))
                # dWt[ci, co, k] = x^T [C, B*T] x im2col^T(dy, stride 2, offsets k-1) [C*4, B*T]
                st.lane = 1
                xT = self.im2col_t(st, x.hl, x.ld, Bn, Tn, Tn, Cc, [0], 1)
                col = self.im2col_t(st, dOh, ldh, Bn, 2 * Tn, Tn, Cc, [-1, 0, 1, 2], 2)
                self.wgrad(st, xT, Cc, col, Cc * 4, g(conv.weight).view(Cc, -1), 0)
                st.lane = 0
                progd = convs.up1d_dgrad(ldh, Bn, Tn)
                wd = self.weight(lambda: convs.up1d_dgrad_weight(conv.weight), Cc, progd.ktot, bwd=True)
                if x.grad is None:
                    x.grad = _Ref(self.zeros(rows, Cc), 0, Cc)
                    res = None
                else:
                    res = x.grad.win
                self.igemm(st, srcs=[(dOh, ldh, progd.src_dims[0])], taps=progd.taps, w=wd, out_dims=progd.out_dims,
                           cout=Cc, out_f32=x.grad.win, residual=res)
            self.tape.append(plan_bwd)
            return out

        # ---------------- UNet ----------------
        din = model.input_dim
        self.x_in_base = self.zeros(Bn * T, 16)               # input rows zero padded to 16 channels
        self.x_in = self.x_in_base[:, :din]
        x0 = _Node(Bn, T, din, ld=16, f32=None, hl=self.hlz(Bn * T, 16))
        self.x0 = x0
        self.fwd.add("split_hl", lambda: _lib.check(self.lib.v2a_split_hl(self.x_in_base.data_ptr(), Bn * T, 16, 16,
                                                                x0.hl.hi.data_ptr(), x0.hl.lo.data_ptr(),
                                                                ops._stream()), "split"))
        x, hs = x0, []
        for (r1, r2, dn) in model.down_modules:
            x = res_block(r1, [x])
            x = res_block(r2, [x])
            hs.append(x)
            if not isinstance(dn, nn.Identity):
                x = down(dn, x)
        for m in model.mid_modules:
            x = res_block(m, [x])
        for (r1, r2, upm) in model.up_modules:
            x = res_block(r1, [x, hs.pop()])
            x = res_block(r2, [x])
            if not isinstance(upm, nn.Identity):
                x = up(upm, x)
        # final: Conv1dBlock (default 8 groups) + 1x1 conv to input_dim (row padded to 16)
        fb, fc = model.final_conv[0].block, model.final_conv[1]
        Cs, rows = x.C, Bn * T
        assert x.T == T
        yf, mrf = self.zeros(rows, Cs), self.zeros(Bn, 2 * fb[1].num_groups)
        hfin = _Node(Bn, T, Cs, hl=self.hlz(rows, Cs))
        gaf, bef = self.vec(lambda: fb[1].weight, Cs), self.vec(lambda: fb[1].bias, Cs)
        kf, pf = fb[0].kernel_size[0], fb[0].padding[0]
        self.conv_fwd(lambda: fb[0].weight, lambda: fb[0].bias, Cs, kf, pf, [x], out_f32=yf)
        self.gn(self.fwd, False, B=Bn, T=T, C=Cs, groups=fb[1].num_groups, eps=fb[1].eps, y=yf, gamma=gaf, beta=bef,
                out_hi=hfin.hl.hi, out_lo=hfin.hl.lo, ld_hl=Cs, mean_rstd=mrf)
        self.out16 = self.zeros(rows, 16)
        self.conv_fwd(lambda: fc.weight, lambda: fc.bias, din, 1, 0, [hfin], out_f32=self.out16)
        self.dout16 = self.zeros(rows, 16)
        x_last = x

        def plan_final_bwd():
            st = self.bwd
            dO = _Ref(self.dout16, 0, din)
            dOh, ldh, dOT = self.grad_prep(st, dO, rows, colsum=g(fc.bias))
            self.conv_bwd(st, lambda: fc.weight, g(fc.weight).view(din, -1), 1, 0, [hfin], dOh, ldh, dOT, din)
            kp = _c64(rows)
            dyf = self.hlz(rows, Cs)
            dyfT = self.hlz(Cs, kp) if self._needs_dyT(kf, [x_last]) else HL(None, None)
            self.gn(st, True, B=Bn, T=T, C=Cs, groups=fb[1].num_groups, eps=fb[1].eps, y=yf, gamma=gaf, beta=bef,
                    mean_rstd=mrf, dout=hfin.grad.ptr, ld_dout=hfin.grad.ld, dy_hi=dyf.hi, dy_lo=dyf.lo,
                    dyT_hi=dyfT.hi, dyT_lo=dyfT.lo, ld_T=kp, dbias=g(fb[0].bias), dgamma=g(fb[1].weight),
                    dbeta=g(fb[1].bias))
            self.conv_bwd(st, lambda: fb[0].weight, g(fb[0].weight).view(Cs, -1), kf, pf, [x_last], dyf, Cs, dyfT, Cs)
        self.tape.append(plan_final_bwd)

        # ---------------- backward plan: tape in reverse, then the conditioning path ----------------
        for planner in reversed(self.tape):
            planner()
        st = self.bwd
        dF = _Ref(self.dfilm, 0, ftot)
        self.dbfilm = torch.zeros(ftot, dtype=torch.float32, device=dev)
        self.dwfilm = self.zeros(ftot, cd)
        self._zero_each_bwd.append(self.dbfilm)
        dFh, ldF, dFT = self.grad_prep(st, dF, Bn, colsum=self.dbfilm)
        self.conv_bwd(st, wfilm, self.dwfilm, 1, 0, [n_mgf], dFh, ldF, dFT, ftot)
        dgf = self.zeros(Bn, cd)
        self.act_bwd(st, self.gf, n_mgf.grad.t, dgf, Bn * cd, 2)
        self.dgf = dgf

        # rows of the concatenated FiLM gradient -> the per-block windows of the gradient slab (2 launches)
        base = self.gslab.data_ptr()
        w_off, b_off, o = [], [], 0
        for m in blocks:
            n2 = 2 * m.out_channels
            gw, gb = g(m.cond_encoder[1].weight), g(m.cond_encoder[1].bias)
            w_off += [(gw.data_ptr() - base) // 4 + r * cd for r in range(n2)]
            b_off += [(gb.data_ptr() - base) // 4 + r for r in range(n2)]
            o += n2
        w_off_t = torch.tensor(w_off, dtype=torch.int64, device=dev)
        b_off_t = torch.tensor(b_off, dtype=torch.int64, device=dev)
        self.keep.append((w_off_t, b_off_t))
        st.add("scatter_film_w", lambda: _lib.check(self.lib.v2a_scatter_rows(
            self.dwfilm.data_ptr(), cd, ftot, cd, w_off_t.data_ptr(), base, ops._stream()), "scatter_rows"), lane=1)
        st.add("scatter_film_b", lambda: _lib.check(self.lib.v2a_scatter_rows(
            self.dbfilm.data_ptr(), 1, ftot, 1, b_off_t.data_ptr(), base, ops._stream()), "scatter_rows"))
        ddse = _Ref(dgf, 0, dsed)
        d3h, ld3, d3T = self.grad_prep(st, ddse, Bn, colsum=g(lin3.bias))
        self.conv_bwd(st, lambda: lin3.weight.unsqueeze(-1), g(lin3.weight), 1, 0, [n_m1], d3h, ld3, d3T, dsed)
        da1 = self.zeros(Bn, 4 * dsed)
        self.act_bwd(st, a1, n_m1.grad.t, da1, Bn * 4 * dsed, 2)
        d1h, ld1, d1T = self.grad_prep(st, _Ref(da1, 0, 4 * dsed), Bn, colsum=g(lin1.bias))
        self.conv_bwd(st, lambda: lin1.weight.unsqueeze(-1), g(lin1.weight), 1, 0, [n_temb], d1h, ld1, d1T, 4 * dsed)

    # ---- execution -----------------------------------------------------------------
    def _run(self, name: str, steps, pre=(), force_eager: bool = False):
        """Launch a planned list.  Every buffer is static, so after one eager (warm-up) run the list is
        captured into a CUDA graph and replayed (one launch instead of 68 / 169; V2A_NO_GRAPH=1 disables)."""
        def eager():
            # lane 1 = weight-gradient chains (im2col^T -> wgrad GEMM): nothing on the main lane reads their
            # results, so they run on a side stream beside the data-gradient chain and join at the end
            main = torch.cuda.current_stream()
            for z in pre:
                z.zero_()
            used_side = False
            for fn, lane in zip(steps, steps.lanes):
                if lane == 1 and self._side is not None:
                    ev = torch.cuda.Event()
                    ev.record(main)
                    self._side.wait_event(ev)
                    with torch.cuda.stream(self._side):
                        fn()
                    used_side = True
                else:
                    if lane == 2 and used_side:      # a main-lane step that consumes what the side lane produced
                        main.wait_stream(self._side)
                    fn()
            if used_side:
                main.wait_stream(self._side)
        if force_eager or os.environ.get("V2A_NO_GRAPH", "0") == "1":   # force_eager: the caller captures a larger graph
            return eager()
        seen = self._graphs.get(name)
        if seen is None:            # first call: eager (lazy CUDA module loads must not happen under capture)
            self._graphs[name] = False
            return eager()
        if seen is False:
            g = torch.cuda.CUDAGraph()
            cur = torch.cuda.current_stream()
            side = torch.cuda.Stream()
            side.wait_stream(cur)
            with torch.cuda.stream(side):
                with torch.cuda.graph(g, stream=side):
                    eager()
            cur.wait_stream(side)
            self._graphs[name] = seen = g
        seen.replay()

    def forward(self, sample, t, gc):
        self.refresh_weights()
        Bn, T = self.B, self.T
        self.t_buf.copy_(t)
        self.x_in.copy_(sample.reshape(Bn * T, -1))
        if self.gc_in.shape[1]:
            self.gc_in.copy_(gc)
        self._run("fwd", self.fwd)
        self.fwd_token += 1
        return self.out16[:, :self.x0.C].reshape(Bn, T, -1).clone()

    def backward(self, grad_out, clone_param_grads=True):
        Bn, T = self.B, self.T
        self.dout16[:, :self.x0.C].copy_(grad_out.reshape(Bn * T, -1))
        self.wait_bwd_weights()
        self._run("bwd", self.bwd, pre=self._zero_each_bwd)
        d_sample = self.x0.grad.win.reshape(Bn, T, -1).clone()
        d_gc = self.dgf[:, self.dgf.shape[1] - self.gc_in.shape[1]:].clone()
        if clone_param_grads:
            pgrads = [self.pgrad[id(p)].clone() for p in self.params]
        else:  # slab mode: the gradients stay in self.gslab for the fused optimiser
            pgrads = [None] * len(self.params)
        return d_sample, d_gc, pgrads
