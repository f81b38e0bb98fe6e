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
      GroupNorm/Mish/FiLM backward (one CTA per sample) emits dy (hi/lo), dy^T, bias/gamma/beta/FiLM grads
      data gradient  = implicit GEMM over dy with flipped weights (accumulating in its epilogue)
      weight gradient = GEMM dy^T x im2col^T(x), written straight into the parameter-gradient slab
  all 12 FiLM linears run as ONE concatenated GEMM (forward, dgrad and wgrad).

Activations are channels-last [B, T, C] — exactly the layout ``sample`` arrives in, so the
reference's two rearranges (conditional_unet1d.py:189,245) disappear.
"""
from __future__ import annotations

import ctypes as C
import os
import weakref
from typing import Dict, List, Optional, Tuple, Union

import torch
import torch.nn as nn

from . import _lib, convs, ops
from .ops import HL

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
        params = list(self.parameters())
        return _UNet1DFunction.apply(self, eng, sample, t, gc, *params)


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
            out = eng.forward(model, sample.detach().float(), t, gc.detach().float())
        ctx.eng = eng
        ctx.token = eng.fwd_token
        ctx.nparams = len(params)
        ctx.in_dtypes = (sample.dtype, gc.dtype)
        return out

    @staticmethod
    def backward(ctx, grad_out):
        eng: _PolicyEngine = ctx.eng
        if ctx.token != eng.fwd_token:
            raise RuntimeError("v2a_b200.ConditionalUnet1D: backward() after another forward() on the same "
                               "module/shape — activations are kept in static buffers (one forward per backward)")
        with torch.autocast("cuda", enabled=False):
            d_sample, d_gc, pgrads = eng.backward(grad_out.float())
        return (None, None, d_sample.to(ctx.in_dtypes[0]), None, d_gc.to(ctx.in_dtypes[1]), *pgrads)


# ---------------------------------------------------------------------------
# engine
# ---------------------------------------------------------------------------
class _Ref:
    """A column window of a row-pitched fp32 matrix (gradient buffers are shared through these)."""

    def __init__(self, t: torch.Tensor, off: int, C: int):
        self.t, self.off, self.C = t, off, C
        self.ld = t.shape[1]

    @property
    def view(self):
        return self.t[:, self.off:self.off + self.C]

    @property
    def ptr(self):
        return self.t.data_ptr() + 4 * self.off


class _Node:
    """An activation [B*T, C]: fp32 and/or hi/lo planes, plus its gradient window (set during backward planning)."""

    def __init__(self, B, T, Cc, f32=None, hl=None, ld=None):
        self.B, self.T, self.C = B, T, Cc
        self.f32, self.hl = f32, hl
        self.ld = ld or Cc           # row pitch of hl / f32 (>= C when zero padded)
        self.grad: Optional[_Ref] = None


def _gn_desc(**kw) -> _lib.PolicyGnDesc:
    d = _lib.PolicyGnDesc()
    for k, v in kw.items():
        if isinstance(v, torch.Tensor):
            v = v.data_ptr()
        setattr(d, k, v)
    return d


class _PolicyEngine:
    def __init__(self, model: ConditionalUnet1D, B, T, device):
        self.model_ref = weakref.ref(model)
        self.B, self.T, self.device = B, T, device
        self.passes = int(os.environ.get("V2A_PASSES", "3"))
        self.lib = _lib.load()
        self.fwd: List = []
        self.bwd: List = []
        self.packers: List = []
        self.vec_packers: List = []
        self.keep: List = []
        self.fwd_token = 0
        self._wkey = None
        self.igemms: List[ops.Igemm] = []
        f32 = dict(dtype=torch.float32, device=device)
        params = list(model.parameters())
        self.params = params
        # flat parameter-gradient slab with per-parameter views
        tot = sum(p.numel() for p in params)
        self.gslab = torch.zeros(tot, **f32)
        self.pgrad: Dict[int, torch.Tensor] = {}
        off = 0
        for p in params:
            self.pgrad[id(p)] = self.gslab[off:off + p.numel()].view(p.shape)
            off += p.numel()
        self._build(model)

    # ---- small helpers ------------------------------------------------------
    def f32(self, rows, cols):
        return torch.zeros(rows, cols, dtype=torch.float32, device=self.device)

    def hlbuf(self, rows, cols) -> HL:
        h = HL.empty(rows, cols, self.device)
        h.hi.zero_()
        h.lo.zero_()
        return h

    def weight(self, fn, rows, cols) -> HL:
        hl = HL.empty(rows, cols, self.device)
        self.packers.append((fn, hl))
        return hl

    def vec(self, fn, n) -> torch.Tensor:
        v = torch.empty(n, dtype=torch.float32, device=self.device)
        self.vec_packers.append((fn, v))
        return v

    def refresh_weights(self):
        key = tuple((p.data_ptr(), p._version) for p in self.params)
        if key == self._wkey:
            return
        with torch.no_grad():
            for fn, hl in self.packers:
                w = fn().detach().to(self.device, torch.float32)
                hi = w.to(torch.bfloat16)
                hl.hi.copy_(hi)
                hl.lo.copy_((w - hi.float()).to(torch.bfloat16))
            for fn, v in self.vec_packers:
                v.copy_(fn().detach().to(self.device, torch.float32).reshape(-1))
        self._wkey = key

    def igemm(self, steps, **kw) -> ops.Igemm:
        g = ops.Igemm(passes=self.passes, **kw)
        self.igemms.append(g)
        steps.append(g.run)
        return g

    def call(self, steps, fn, *args):
        steps.append(lambda: _lib.check(fn(*args, ops._stream()), fn.__name__ if hasattr(fn, "__name__") else "call"))

    def gn_fwd(self, desc):
        self.keep.append(desc)
        self.fwd.append(lambda: _lib.check(self.lib.v2a_policy_gn_act_fwd(C.byref(desc), ops._stream()), "gn_act_fwd"))

    def gn_bwd(self, steps, desc):
        self.keep.append(desc)
        steps.append(lambda: _lib.check(self.lib.v2a_policy_gn_act_bwd(C.byref(desc), ops._stream()), "gn_act_bwd"))

    def im2col_t(self, steps, src: HL, ld, c_off, B, Tin, Tout, Cc, offsets, stride, out: HL):
        arr = (C.c_int * len(offsets))(*offsets)
        self.keep.append(arr)
        steps.append(lambda: _lib.check(self.lib.v2a_policy_im2col_t(
            src.hi.data_ptr(), src.lo.data_ptr(), ld, c_off, B, Tin, Tout, Cc, len(offsets), stride, arr,
            out.hi.data_ptr(), out.lo.data_ptr(), ops._stream()), "im2col_t"))

    def grad_prep(self, steps, ref: _Ref, rows, hl: Optional[HL], ld_hl, tr: Optional[HL], colsum):
        steps.append(lambda: _lib.check(self.lib.v2a_grad_prep(
            ref.ptr, rows, ref.C, ref.ld, None if hl is None else hl.hi.data_ptr(),
            None if hl is None else hl.lo.data_ptr(), ld_hl, None if tr is None else tr.hi.data_ptr(),
            None if tr is None else tr.lo.data_ptr(), None if colsum is None else colsum.data_ptr(),
            ops._stream()), "grad_prep"))

    def add_into(self, steps, dst: _Ref, src: _Ref, rows, accumulate=True):
        steps.append(lambda: _lib.check(self.lib.v2a_add_strided(dst.ptr, dst.ld, src.ptr, src.ld, rows, dst.C,
                                                                1 if accumulate else 0, ops._stream()), "add"))

    # ---- gradient fan-in ----------------------------------------------------
    def grad_target(self, node: _Node) -> Tuple[_Ref, bool]:
        """Where a producer of d(node) must write, and whether it must accumulate."""
        if node.grad is None:
            node.grad = _Ref(self.f32(node.B * node.T, node.C), 0, node.C)
            return node.grad, False
        return node.grad, True

    # ---- generic conv (stride-1 Conv1d over [B, T, C], optionally over a 2-source concat) -----------------
    def conv_fwd(self, conv: nn.Conv1d, ins: List[_Node], out_f32=None, out_hl=None, residual=None, ldc=None):
        B, T = ins[0].B, ins[0].T
        k, pad = conv.kernel_size[0], conv.padding[0]
        cins = [n.C for n in ins]
        cout = conv.out_channels
        prog = convs.conv1d_cat([n.ld for n in ins], B, T, k, pad)
        # weights sliced per source; sources narrower than their pitch (zero-padded input) read zeros
        def wfn(conv=conv, ins=ins):
            parts, off = [], 0
            for n in ins:
                for j in range(k):
                    w = conv.weight[:, off:off + n.C, j]
                    parts.append(torch.nn.functional.pad(w, (0, n.ld - n.C)) if n.ld != n.C else w)
                off += n.C
            return ops.pack_weight_taps(parts)
        w = self.weight(wfn, cout, prog.ktot)
        b = self.vec(lambda conv=conv: conv.bias, cout)
        self.igemm(self.fwd, srcs=[(n.hl, n.ld, d) for n, d in zip(ins, prog.src_dims)], taps=prog.taps, w=w,
                   out_dims=prog.out_dims, cout=cout, ldc=ldc, out_f32=out_f32, out_hl=out_hl, bias=b,
                   residual=residual)

    def conv_bwd(self, steps, conv: nn.Conv1d, ins: List[_Node], dy_hl: HL, dyT_hl: HL, ld_dy: int,
                 extra_residual: Optional[_Ref] = None):
        """dW (into the gradient slab) and dX (fan-in aware) of a stride-1 Conv1d given dy planes."""
        B, T = ins[0].B, ins[0].T
        rows = B * T
        k, pad = conv.kernel_size[0], conv.padding[0]
        cout = conv.out_channels
        cin_tot = sum(n.C for n in ins)
        dW = self.pgrad[id(conv.weight)].view(cout, cin_tot * k)
        off = 0
        for n in ins:  # weight gradient per concat part: dy^T [cout, rows] x im2col^T(x) [C*k, rows]
            col = self.scratch_hl("col", n.C * k, rows)
            self.im2col_t(steps, n.hl, n.ld, 0, B, T, T, n.C, [j - pad for j in range(k)], 1, col)
            progw = convs.pointwise(rows, (cout,))
            self.igemm(steps, srcs=[(dyT_hl, rows, progw.src_dims[0])], taps=progw.taps, w=col,
                       out_dims=progw.out_dims, cout=n.C * k, ldc=cin_tot * k,
                       out_f32=_window(dW, off * k, n.C * k))
            off += n.C
        # data gradient over the whole (concatenated) input, then fan out to the parts
        need = [n for n in ins if n.f32 is not None or n.grad is not None or getattr(n, "needs_grad", True)]
        if not need:
            return
        progd = convs.conv1d(ld_dy, B, T, k, pad)
        wd = self.weight(lambda conv=conv: torch.nn.functional.pad(
            convs.conv1d_dgrad_weight(conv.weight), (0, 0)), cin_tot, None)
        # (cols fixed below: K = k * pad64(ld_dy); conv1d_dgrad_weight pads Cout to 64 per tap already)
        if len(ins) == 1:
            tgt, acc = self.grad_target(ins[0])
            res = tgt if acc else extra_residual
            if acc and extra_residual is not None:
                self.add_into(steps, tgt, extra_residual, rows)
            self.igemm(steps, srcs=[(dy_hl, ld_dy, progd.src_dims[0])], taps=progd.taps, w=wd,
                       out_dims=progd.out_dims, cout=cin_tot, ldc=tgt.ld, out_f32=_window(tgt.t, tgt.off, tgt.C),
                       residual=None if res is None else _window(res.t, res.off, res.C))
        else:
            dcat = self.f32(rows, cin_tot)
            self.igemm(steps, srcs=[(dy_hl, ld_dy, progd.src_dims[0])], taps=progd.taps, w=wd,
                       out_dims=progd.out_dims, cout=cin_tot, out_f32=dcat,
                       residual=None if extra_residual is None else _window(extra_residual.t, extra_residual.off,
                                                                           extra_residual.C))
            off = 0
            for n in ins:
                part = _Ref(dcat, off, n.C)
                if n.grad is None:
                    n.grad = part
                else:
                    self.add_into(steps, n.grad, part, rows)
                off += n.C

    def scratch_hl(self, tag, rows, cols) -> HL:
        """Reusable scratch planes (backward runs sequentially on one stream)."""
        key = (tag, rows * cols)
        pool = self.__dict__.setdefault("_scratch", {})
        if key not in pool:
            pool[key] = HL.empty(rows * cols, 1, self.device)
        h = pool[key]
        return HL(h.hi.view(rows, cols), h.lo.view(rows, cols))

    # ---- plan -----------------------------------------------------------------
    def _build(self, model: ConditionalUnet1D):
        raise NotImplementedError  # replaced below (kept separate for readability)


def _window(t: torch.Tensor, off: int, C_: int) -> torch.Tensor:
    """Column window view whose data_ptr / row pitch the kernels use (rows stay pitched by t.shape[1])."""
    return t[:, off:off + C_]
