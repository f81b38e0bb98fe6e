"""B200-native ``VisualCore`` (ResNet18-GroupNorm trunk -> SpatialSoftmax -> Linear), forward AND backward.

Row P6 / N1 of SURVEY.md section 8: the two observation encoders are 80 % of the FLOPs of
``DiffusionUnetImagePolicy.compute_loss``.  The parameter-holding modules stay the reference's
(diffuser/diffusion_policy/common/vision_nets.py:9-177, common/base_nets.py:153-285,
model/multi_image_obs_encoder.py:67-74; same ``state_dict``); this module plans, per (VisualCore, batch):

  forward   stem im2col (7x7 s2, K = 147 -> 192) -> tcgen05 GEMM (+ GroupNorm sums in its epilogue)
            -> GroupNorm + ReLU + MaxPool(3, 2, 1)
            8 BasicBlocks: 3x3 conv (TMA zero-padded taps; stride 2 on a phase-split operand) -> GN + ReLU
            -> 3x3 conv -> relu(GN + identity | GN(1x1 stride-2 downsample conv))
            keypoint 1x1 conv -> spatial softmax expectation -> Linear
  backward  GroupNorm / ReLU backward in two passes (per-(image, channel) sums, then the gradient as bf16
            hi/lo planes), data gradient = implicit GEMM with flipped weights (stride 2: one GEMM producing
            the 4 input phases), weight gradient = MN-major tcgen05 GEMM straight from the channels-last
            planes (no transposes, no im2col), MaxPool / SpatialSoftmax / Linear backward kernels.

Parameter gradients land in one flat slab in ``parameters()`` order (``engine.gslab``).
"""
from __future__ import annotations

import ctypes as C
import math
import os
import weakref
from typing import Dict, List, Optional

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import _lib, convs, ops, packing
from .ops import HL
from .packing import PackedParams
from .policy_unet1d import _Steps

# Forward operand planes: bf16 (hi, lo) pairs by default, like every other contraction of the library.
# V2A_ENCODER_FWD_PLANES=fp16 switches activations and weights of the FORWARD convs to fp16 (hi, lo) pairs (22+
# significant bits; bf16 twins are then written for the weight-gradient GEMM, whose gradient operand stays bf16
# for range and a tcgen05 MMA takes one operand format).  Measured (tests/test_encoder_gpu.py, B200): forward
# rel-L2 vs float64 2.4e-5 (bf16 planes) / 1.4e-5 (fp16 planes) -- the floor is the tensor core's truncating fp32
# accumulation, not the operand width, so the ReLU-mask flips that limit gradient parity (see the test) do not go
# away and the cheaper format is the default.
FWD16 = os.environ.get("V2A_ENCODER_FWD_PLANES", "bf16") == "fp16"

_ENGINES: "weakref.WeakKeyDictionary[nn.Module, Dict[tuple, _EncoderEngine]]" = weakref.WeakKeyDictionary()
_SLAB_GRADS: "weakref.WeakKeyDictionary[nn.Module, bool]" = weakref.WeakKeyDictionary()
_LAST_ENGINE: "weakref.WeakKeyDictionary[nn.Module, _EncoderEngine]" = weakref.WeakKeyDictionary()


def set_slab_grads(core: nn.Module, on: bool = True) -> None:
    """Leave parameter gradients in ``engine.gslab`` (flat, parameters() order) and return None to autograd."""
    _SLAB_GRADS[core] = bool(on)


def last_engine(core: nn.Module) -> "Optional[_EncoderEngine]":
    return _LAST_ENGINE.get(core)


def invalidate_weights(core: nn.Module) -> None:
    for eng in _ENGINES.get(core, {}).values():
        eng._wkey = None


def encoder_engine(core, B: int, device) -> "_EncoderEngine":
    per = _ENGINES.setdefault(core, {})
    device = torch.device(device)
    if device.index is None:
        device = torch.device("cuda", torch.cuda.current_device())
    key = (B, str(device))
    eng = per.get(key)
    if eng is None:
        if len(per) >= 3:
            per.pop(next(iter(per)))
        eng = _EncoderEngine(core, B, device)
        per[key] = eng
    return eng


def visual_core_forward(core, x: torch.Tensor) -> torch.Tensor:
    """x [B, 3, H, W] (already normalised) -> features [B, feature_dimension] through the planned CUDA engine."""
    if not x.is_cuda:
        raise RuntimeError("v2a_b200 VisualCore runs on CUDA only (no CPU fallback)")
    eng = encoder_engine(core, x.shape[0], x.device)
    eng.training_call = packing.training_call(eng.params)
    _LAST_ENGINE[core] = eng
    return _VisualCoreFunction.apply(core, eng, x, *list(core.parameters()))


class _VisualCoreFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, core, eng, x, *params):
        with torch.autocast("cuda", enabled=False):
            out = eng.forward(x.detach().float())
        ctx.eng = eng
        ctx.slab = _SLAB_GRADS.get(core, False)
        ctx.token = eng.fwd_token
        return out

    @staticmethod
    def backward(ctx, grad_out):
        eng: _EncoderEngine = ctx.eng
        if ctx.token != eng.fwd_token:
            raise RuntimeError("v2a_b200.VisualCore: backward() after another forward() on the same module/shape "
                               "(activations live in static buffers: one forward per backward)")
        with torch.autocast("cuda", enabled=False):
            pgrads = eng.backward(grad_out.float(), clone_param_grads=not ctx.slab)
        hook = getattr(eng, "on_backward_done", None)
        if hook is not None:       # one-shot: PolicyTrainStep starts this slab's all-reduce the moment it is complete
            eng.on_backward_done = None
            hook()
        return (None, None, None, *pgrads)


class _Act:
    """Activation [N, H, W, C] channels-last: fp32 and/or bf16 planes (normal or stride-2 phase-split layout)."""

    def __init__(self, N, H, W, Cc, f32=None, hl=None, hl_ps=None):
        self.N, self.H, self.W, self.C = N, H, W, Cc
        self.rows = N * H * W
        self.f32, self.hl, self.hl_ps = f32, hl, hl_ps
        self.tw: Optional[HL] = None      # bf16 twin of hl / hl_ps (same layout): x operand of the weight gradient
        self.grad: Optional[torch.Tensor] = None


def dgrad3x3_weight(w: torch.Tensor) -> torch.Tensor:
    """[Cout, Cin, 3, 3] -> [Cin, 9 * pad64(Cout)]: dx[p] = sum_taps W[:, :, 2-kh, 2-kw]^T dy[p + (kh-1, kw-1)]."""
    return ops.pack_weight_taps([w[:, :, 2 - kh, 2 - kw].t() for kh in range(3) for kw in range(3)])


_S2_K = {(0, 0): 1, (1, 0): 2, (1, 1): 0}   # (input phase, dy offset) -> kernel index, stride-2 3x3 pad 1


def dgrad3x3_s2_weight(w: torch.Tensor) -> torch.Tensor:
    """[Cout, Cin, 3, 3] -> [4 * Cin, 4 * pad64(Cout)]: rows (py, px, ci); taps dy[(j + dj, i + di)], (dj, di) in
    {0,1}^2; input row 2j + py gets kernel row ky = py - 2 dj + 1 when that is in 0..2."""
    z = torch.zeros_like(w[:, :, 0, 0].t())
    rows = []
    for py in range(2):
        for px in range(2):
            parts = []
            for dj in range(2):
                for di in range(2):
                    ky, kx = _S2_K.get((py, dj)), _S2_K.get((px, di))
                    parts.append(z if ky is None or kx is None else w[:, :, ky, kx].t())
            rows.append(ops.pack_weight_taps(parts))
    return torch.cat(rows, 0)


class _EncoderEngine(PackedParams):
    GN_EPS_DEFAULT = 1e-5

    def __init__(self, core, B: int, device):
        self.core_ref = weakref.ref(core)
        self.B, self.device = B, device
        self.passes = int(os.environ.get("V2A_PASSES", "3"))
        self.lib = _lib.load()
        self.fwd: _Steps = _Steps()
        self.bwd: _Steps = _Steps()
        self._init_packing()
        self.keep: List = []
        self.igemms: List = []
        self.wgrads: List = []
        self.fwd_token = 0
        self.params = list(core.parameters())
        tot = sum(p.numel() for p in self.params)
        self.gslab = torch.zeros(tot, dtype=torch.float32, device=device)
        self.pgrad: Dict[int, torch.Tensor] = {}
        off = 0
        for p in self.params:
            self.pgrad[id(p)] = self.gslab[off:off + p.numel()].view(p.shape)
            off += p.numel()
        self._stats_req: List = []       # (holder, reps, N, C)
        self._scratch_req: List = []     # (holder, rows, cols) weight-gradient scratch, zeroed every backward
        self._graphs: Dict[str, object] = {}
        self._side = None if os.environ.get("V2A_NO_SIDE_STREAM", "0") == "1" else torch.cuda.Stream(device=device)
        self.done = torch.cuda.Event()
        self._build(core)
        self._bind_arenas()
        self._trace_packers()

    # ---- buffers -------------------------------------------------------------------------------
    def zeros(self, *shape):
        return torch.zeros(*shape, dtype=torch.float32, device=self.device)

    def hlz(self, rows, cols, fp16: bool = False) -> HL:
        """zeroed operand planes; fp16=True: fp16 (hi, lo) pairs (forward operands), else bf16 (gradients)"""
        return HL(torch.zeros(rows, cols, dtype=torch.bfloat16, device=self.device),
                  torch.zeros(rows, cols, dtype=torch.bfloat16, device=self.device), fp16)

    def stats(self, N, Cc, HW):
        """float64 [reps, N, C, 2] GroupNorm sums filled by the producing igemm (bound after planning)."""
        holder: List[torch.Tensor] = []
        self._stats_req.append((holder, max(1, min(8, HW // 512)), N, Cc))
        return holder

    def scratch(self, rows, cols):
        holder: List[torch.Tensor] = []
        self._scratch_req.append((holder, rows, cols))
        return holder

    def _bind_arenas(self):
        n = sum(r * N * Cc * 2 for _, r, N, Cc in self._stats_req)
        self.stats_arena = torch.zeros(max(n, 1), dtype=torch.float64, device=self.device)
        off = 0
        for holder, r, N, Cc in self._stats_req:
            m = r * N * Cc * 2
            holder.append(self.stats_arena[off:off + m].view(r, N, Cc, 2))
            off += m
        n = sum(r * c for _, r, c in self._scratch_req)
        self.wg_arena = torch.zeros(max(n, 1), dtype=torch.float32, device=self.device)
        off = 0
        for holder, r, c in self._scratch_req:
            holder.append(self.wg_arena[off:off + r * c].view(r, c))
            off += r * c
        for fn in self._deferred:
            fn()
        self._deferred = []

    # ---- launch wrappers (plans are created after the arenas exist) ------------------------------
    @staticmethod
    def _block_n(rows: int, cout: int) -> int:
        """N tile of a forward conv: its epilogue accumulates GroupNorm sums, so it cannot split K to fill the
        machine; when the widest N tile leaves fewer tiles than SMs (layers 3-4: 64 ... 32 tiles), narrow it."""
        n = ops.choose_block_n(cout)
        m_tiles = -(-rows // 128)
        while n > 64 and n % 32 == 0 and m_tiles * -(-cout // n) < 120:
            n //= 2
        return n

    def igemm(self, steps, tag="igemm", **kw):
        slot = [None]
        if kw.get("stats") is not None and "block_n" not in kw:
            od = kw["out_dims"]
            kw["block_n"] = self._block_n(od[0] * od[1] * od[2] * od[3], kw["cout"])

        def make():
            args = {k: (v[0] if isinstance(v, list) and len(v) == 1 and isinstance(v[0], torch.Tensor) else v)
                    for k, v in kw.items()}
            g = ops.Igemm(passes=self.passes, **args)
            slot[0] = g
            self.igemms.append(g)
        self._deferred.append(make)
        steps.add(tag, lambda: slot[0].run())

    def wgrad(self, steps, *, srcs, units, dy: HL, dy_channels, dy_dims, cout, cin, ntaps, param):
        sc = self.scratch(64 * len(units), cout)
        slot = [None]

        def make():
            g = ops.Wgrad(srcs=srcs, units=units, dy=dy, dy_channels=dy_channels, dy_dims=dy_dims, cout=cout,
                          out=sc[0], passes=ops.wgrad_passes(self.passes, int(math.prod(dy_dims))))
            slot[0] = g
            self.wgrads.append(g)
        self._deferred.append(make)
        steps.add(f"wgrad M{64 * len(units)} N{cout}", lambda: slot[0].run(), lane=1)
        dst = self.pgrad[id(param)]
        steps.add("wgrad_scatter", lambda: ops.wgrad_scatter(sc[0], cout, cin, ntaps, dst, accumulate=False), lane=1)

    def gn_finalize(self, steps, st, N, Cc, groups, HW, eps, mr):
        steps.add("gn_finalize", lambda: _lib.check(self.lib.v2a_gn_finalize(
            st[0].data_ptr(), st[0].shape[0], st[0].stride(0), N, Cc, groups, HW, eps, mr.data_ptr(), ops._stream()),
            "gn_finalize"))

    def enc_prep(self, steps, *, xa, mra, gn_a, N, H, W, Cc, xb=None, mrb=None, gn_b=None, idn=None, relu=True,
                 out_f32=None, out_hl=None, out_tw=None, phase_split=False):
        d = _lib.EncPrepDesc()
        ga, ba = self.vec(lambda: gn_a.weight, Cc), self.vec(lambda: gn_a.bias, Cc)
        d.xa, d.mean_rstd_a, d.gamma_a, d.beta_a = xa.data_ptr(), mra.data_ptr(), ga.data_ptr(), ba.data_ptr()
        keep = [xa, mra, ga, ba, xb, mrb, idn, out_f32, out_hl]
        if xb is not None:
            gb, bb = self.vec(lambda: gn_b.weight, Cc), self.vec(lambda: gn_b.bias, Cc)
            d.xb, d.mean_rstd_b, d.gamma_b, d.beta_b = xb.data_ptr(), mrb.data_ptr(), gb.data_ptr(), bb.data_ptr()
            keep += [gb, bb]
        if idn is not None:
            d.idn = idn.data_ptr()
        d.groups, d.C, d.H, d.W, d.images = gn_a.num_groups, Cc, H, W, N
        d.relu, d.phase_split = int(relu), int(phase_split)
        d.plane_fmt = int(bool(out_hl is not None and out_hl.fp16))
        if out_f32 is not None:
            d.out_f32 = out_f32.data_ptr()
        if out_hl is not None:
            d.out_hi, d.out_lo = out_hl.hi.data_ptr(), out_hl.lo.data_ptr()
        if out_tw is not None and out_tw is not out_hl:
            d.out2_hi, d.out2_lo = out_tw.hi.data_ptr(), out_tw.lo.data_ptr()
            keep.append(out_tw)
        self.keep.append((d, keep))
        steps.add("enc_prep", lambda: _lib.check(self.lib.v2a_enc_prep(C.byref(d), ops._stream()), "enc_prep"))
        return ga, ba

    def gn_bwd(self, steps, *, dout, raw, mr, gn: nn.GroupNorm, N, HW, Cc, mask_mode, outv=None, d_hl: HL, g_out=None):
        d = _lib.EncGnBwdDesc()
        ga, ba = self.vec(lambda: gn.weight, Cc), self.vec(lambda: gn.bias, Cc)
        sums = self.scratch(N * Cc * 2, 1)      # pass-1 sums: part of the arena cleared once per backward
        coef = None
        d.dout, d.raw, d.mean_rstd = dout.data_ptr(), raw.data_ptr(), mr.data_ptr()
        d.outv = None if outv is None else outv.data_ptr()
        d.gamma, d.beta = ga.data_ptr(), ba.data_ptr()
        d.mask_mode, d.groups, d.C, d.HW, d.images = mask_mode, gn.num_groups, Cc, HW, N
        self._deferred.append(lambda: setattr(d, "sums", sums[0].data_ptr()))
        d.d_hi, d.d_lo = d_hl.hi.data_ptr(), d_hl.lo.data_ptr()
        d.g_out = None if g_out is None else g_out.data_ptr()
        d.dgamma, d.dbeta = self.pgrad[id(gn.weight)].data_ptr(), self.pgrad[id(gn.bias)].data_ptr()
        self.keep.append((d, [dout, raw, mr, outv, ga, ba, sums, coef, d_hl, g_out]))
        steps.add("enc_gn_bwd", lambda: _lib.check(self.lib.v2a_enc_gn_bwd(C.byref(d), ops._stream()), "enc_gn_bwd"))

    # ---- network plan --------------------------------------------------------------------------
    def _build(self, core):
        self._deferred: List = []
        N, dev = self.B, self.device
        backbone, pool, lin = core.backbone, core.pool, core.nets[3] if len(core.nets) > 3 else None
        if lin is None:
            raise NotImplementedError("VisualCore engine expects the trailing Linear (feature_dimension)")
        conv1, gn0 = backbone.nets[0], backbone.nets[1]
        layers = [backbone.nets[i] for i in range(4, 8)]
        Cin, H, W = core.input_shape
        assert Cin == 3 and H % 32 == 0 and W % 32 == 0, "stem engine is specialised for 3-channel images"
        self.x_in = self.zeros(N, 3, H, W)
        tape: List = []
        G = lambda gn: gn.num_groups

        # ---------------- stem: 7x7 s2 conv as im2col GEMM -> GN -> ReLU -> maxpool ----------------
        H0, W0 = H // 2, W // 2
        HW0 = H0 * W0
        C0 = conv1.out_channels
        col = self.hlz(N * HW0, 192, FWD16)
        col_tw = self.hlz(N * HW0, 192) if FWD16 else col
        self.fwd.add("stem_pack", lambda: _lib.check(self.lib.v2a_enc_stem_pack(
            self.x_in.data_ptr(), 1.0, 0.0, N, H, W, col.hi.data_ptr(), col.lo.data_ptr(), int(FWD16),
            col_tw.hi.data_ptr() if FWD16 else None, col_tw.lo.data_ptr() if FWD16 else None, ops._stream()), "stem_pack"))
        raw0, st0, mr0 = self.zeros(N * HW0, C0), self.stats(N, C0, HW0), self.zeros(N, G(gn0) * 2)
        w0 = self.weight(lambda: F.pad(conv1.weight.reshape(C0, 147), (0, 45)), C0, 192, FWD16)
        prog = convs.pointwise(192, (HW0, N))
        self.igemm(self.fwd, "igemm stem", srcs=[(col, 192, prog.src_dims[0])], taps=prog.taps, w=w0,
                   out_dims=prog.out_dims, cout=C0, out_f32=raw0, stats=st0, stats_mul=(0, 1, 0, 0))
        self.gn_finalize(self.fwd, st0, N, C0, G(gn0), HW0, gn0.eps, mr0)
        self.stem_probe = dict(raw0=raw0, mr0=mr0, H=H0, W=W0)     # parity probes: the stem ReLU's sign pattern
        H1, W1 = H0 // 2, W0 // 2
        P0 = _Act(N, H1, W1, C0, f32=self.zeros(N * H1 * W1, C0), hl=self.hlz(N * H1 * W1, C0, FWD16))
        P0.tw = self.hlz(N * H1 * W1, C0) if FWD16 else P0.hl
        ga0, be0 = self.vec(lambda: gn0.weight, C0), self.vec(lambda: gn0.bias, C0)
        self.fwd.add("gn_relu_maxpool", lambda: _lib.check(self.lib.v2a_enc_gn_relu_maxpool(
            raw0.data_ptr(), mr0.data_ptr(), G(gn0), ga0.data_ptr(), be0.data_ptr(), N, H0, W0, C0,
            P0.f32.data_ptr(), P0.hl.hi.data_ptr(), P0.hl.lo.data_ptr(), int(FWD16),
            P0.tw.hi.data_ptr() if FWD16 else None, P0.tw.lo.data_ptr() if FWD16 else None, ops._stream()),
            "gn_relu_maxpool"))

        def stem_bwd():
            st = self.bwd
            g0 = self.zeros(N * HW0, C0)
            st.add("maxpool_relu_bwd", lambda: _lib.check(self.lib.v2a_enc_maxpool_relu_bwd(
                raw0.data_ptr(), mr0.data_ptr(), G(gn0), ga0.data_ptr(), be0.data_ptr(), P0.f32.data_ptr(),
                P0.grad.data_ptr(), N, H0, W0, C0, g0.data_ptr(), ops._stream()), "maxpool_relu_bwd"))
            d0 = self.hlz(N * HW0, C0)
            self.stem_probe.update(g0=g0, d0=d0)
            self.gn_bwd(st, dout=g0, raw=raw0, mr=mr0, gn=gn0, N=N, HW=HW0, Cc=C0, mask_mode=0, d_hl=d0)
            self.wgrad(st, srcs=[(col_tw, 192, (HW0, N, 1, 1))], units=[(0, (0, 0, 0, 0), ch) for ch in range(3)],
                       dy=d0, dy_channels=C0, dy_dims=(HW0, N, 1, 1), cout=C0, cin=147, ntaps=1, param=conv1.weight)
        tape.append(stem_bwd)

        # ---------------- residual stages ----------------
        blocks = [b for layer in layers for b in layer]
        X = P0
        self.probes: List[dict] = []   # per-block backward intermediates (developer parity probes)
        self.inner_acts: List[HL] = []
        self.acts = [P0]          # block inputs / outputs in order (parity probes read .f32 / .grad)
        for bi, blk in enumerate(blocks):
            nxt_s2 = bi + 1 < len(blocks) and blocks[bi + 1].conv1.stride[0] == 2
            X = self._basic_block(blk, X, nxt_s2, tape)

        # ---------------- head: keypoint conv -> spatial softmax -> Linear ----------------
        K = pool._num_kp
        Pn = X.H * X.W
        kconv = pool.nets
        logits = self.zeros(X.rows, 32 if K <= 32 else -(-K // 16) * 16)
        ldk = logits.shape[1]
        progk = convs.pointwise(X.C, (X.rows,))
        wk = self.weight(lambda: convs.pointwise_weight(kconv.weight), K, progk.ktot, FWD16)
        bk = self.vec(lambda: kconv.bias, K)
        self.igemm(self.fwd, "igemm kp", srcs=[(X.hl, X.C, progk.src_dims[0])], taps=progk.taps, w=wk,
                   out_dims=progk.out_dims, cout=K, out_f32=logits, bias=bk)
        att, kp = self.zeros(N * Pn * K), self.zeros(N, 2 * K)
        pos_x, pos_y = pool.pos_x.reshape(-1).float().to(dev).contiguous(), pool.pos_y.reshape(-1).float().to(dev).contiguous()
        temp = float(pool.temperature.reshape(-1)[0])
        self.fwd.add("spatial_softmax", lambda: _lib.check(self.lib.v2a_enc_spatial_softmax_fwd(
            logits.data_ptr(), ldk, N, Pn, K, temp, pos_x.data_ptr(), pos_y.data_ptr(), att.data_ptr(), kp.data_ptr(),
            ops._stream()), "spatial_softmax_fwd"))
        Fo = lin.out_features
        self.feat = self.zeros(N, Fo)
        w_lin = self.vec(lambda: lin.weight, lin.weight.numel()).view(Fo, 2 * K)
        b_lin = self.vec(lambda: lin.bias, Fo)
        self.fwd.add("linear", lambda: ops.linear(kp, w_lin, b_lin, self.feat))
        self.dfeat = self.zeros(N, Fo)

        def head_bwd():
            st = self.bwd
            dkp = self.zeros(N, 2 * K)
            st.add("linear_bwd", lambda: _lib.check(self.lib.v2a_enc_linear_bwd(
                kp.data_ptr(), self.dfeat.data_ptr(), w_lin.data_ptr(), N, 2 * K, Fo, dkp.data_ptr(),
                self.pgrad[id(lin.weight)].data_ptr(), self.pgrad[id(lin.bias)].data_ptr(), ops._stream()), "linear_bwd"))
            dlog = self.hlz(N * Pn, K)
            st.add("spatial_softmax_bwd", lambda: _lib.check(self.lib.v2a_enc_spatial_softmax_bwd(
                att.data_ptr(), kp.data_ptr(), dkp.data_ptr(), N, Pn, K, temp, pos_x.data_ptr(), pos_y.data_ptr(),
                dlog.hi.data_ptr(), dlog.lo.data_ptr(), self.pgrad[id(kconv.bias)].data_ptr(), ops._stream()),
                "spatial_softmax_bwd"))
            self.wgrad(st, srcs=[(X.tw, X.C, (X.rows, 1, 1, 1))],
                       units=[(0, (0, 0, 0, 0), ch) for ch in range(ops.nchunks(X.C))], dy=dlog, dy_channels=K,
                       dy_dims=(X.rows, 1, 1, 1), cout=K, cin=X.C, ntaps=1, param=kconv.weight)
            X.grad = self.zeros(X.rows, X.C)
            progd = convs.pointwise(K, (X.rows,))
            wd = self.weight(lambda: ops.pack_weight_taps([kconv.weight.reshape(K, X.C).t()]), X.C, progd.ktot, bwd=True)
            self.igemm(st, "igemm dgrad kp", srcs=[(dlog, K, progd.src_dims[0])], taps=progd.taps, w=wd,
                       out_dims=progd.out_dims, cout=X.C, out_f32=X.grad)
        tape.append(head_bwd)
        for plan in reversed(tape):
            plan()

    def _basic_block(self, blk, X: _Act, next_stride2: bool, tape: List) -> _Act:
        N = self.B
        s = blk.conv1.stride[0]
        Ci, Co = X.C, blk.conv1.out_channels
        Ho, Wo = X.H // s, X.W // s
        HWo, rows_o = Ho * Wo, N * Ho * Wo
        gn1, gn2 = blk.bn1, blk.bn2
        ds = blk.downsample
        G = gn1.num_groups
        # conv1 (+ stats) -> GN + ReLU
        raw1, st1, mr1 = self.zeros(rows_o, Co), self.stats(N, Co, HWo), self.zeros(N, G * 2)
        if s == 1:
            prog1, src1, smul = convs.spatial3x3(Ci, N, X.H, X.W), X.hl, (0, 0, 1, 0)
        else:
            prog1, src1, smul = convs.spatial3x3_s2(Ci, N, X.H, X.W), X.hl_ps, (0, 0, 0, 1)
        w1 = self.weight(lambda: convs.spatial3x3_weight(blk.conv1.weight), Co, prog1.ktot, FWD16)
        self.igemm(self.fwd, f"igemm conv1 C{Ci}->{Co} s{s}", srcs=[(src1, Ci, prog1.src_dims[0])], taps=prog1.taps,
                   w=w1, out_dims=prog1.out_dims, cout=Co, out_f32=raw1, stats=st1, stats_mul=smul)
        self.gn_finalize(self.fwd, st1, N, Co, G, HWo, gn1.eps, mr1)
        a1 = self.hlz(rows_o, Co, FWD16)
        a1_tw = self.hlz(rows_o, Co) if FWD16 else a1
        self.enc_prep(self.fwd, xa=raw1, mra=mr1, gn_a=gn1, N=N, H=Ho, W=Wo, Cc=Co, out_hl=a1, out_tw=a1_tw)
        # conv2 (+ stats)
        raw2, st2, mr2 = self.zeros(rows_o, Co), self.stats(N, Co, HWo), self.zeros(N, G * 2)
        prog2 = convs.spatial3x3(Co, N, Ho, Wo)
        w2 = self.weight(lambda: convs.spatial3x3_weight(blk.conv2.weight), Co, prog2.ktot, FWD16)
        self.igemm(self.fwd, f"igemm conv2 C{Co}", srcs=[(a1, Co, prog2.src_dims[0])], taps=prog2.taps, w=w2,
                   out_dims=prog2.out_dims, cout=Co, out_f32=raw2, stats=st2, stats_mul=(0, 0, 1, 0))
        self.gn_finalize(self.fwd, st2, N, Co, G, HWo, gn2.eps, mr2)
        rawd = mrd = gnd = None
        if ds is not None:
            dconv, gnd = ds[0], ds[1]
            rawd, std, mrd = self.zeros(rows_o, Co), self.stats(N, Co, HWo), self.zeros(N, G * 2)
            taps_d = [(0, (0, 0, 0, 0), ops.nchunks(Ci))]
            wdn = self.weight(lambda: convs.pointwise_weight(dconv.weight), Co, 64 * ops.nchunks(Ci), FWD16)
            self.igemm(self.fwd, f"igemm downsample C{Ci}->{Co}", srcs=[(X.hl_ps, Ci, (X.W // 2, X.H // 2, 4, N))],
                       taps=taps_d, w=wdn, out_dims=(Wo, Ho, 1, N), cout=Co, out_f32=rawd, stats=std,
                       stats_mul=(0, 0, 0, 1))
            self.gn_finalize(self.fwd, std, N, Co, G, HWo, gnd.eps, mrd)
        out = _Act(N, Ho, Wo, Co, f32=self.zeros(rows_o, Co))
        if next_stride2:
            out.hl_ps = self.hlz(rows_o, Co, FWD16)
        else:
            out.hl = self.hlz(rows_o, Co, FWD16)
        out.tw = self.hlz(rows_o, Co) if FWD16 else (out.hl_ps if next_stride2 else out.hl)
        self.enc_prep(self.fwd, xa=raw2, mra=mr2, gn_a=gn2, N=N, H=Ho, W=Wo, Cc=Co, xb=rawd, mrb=mrd, gn_b=gnd,
                      idn=None if ds is not None else X.f32, out_f32=out.f32,
                      out_hl=out.hl_ps if next_stride2 else out.hl, out_tw=out.tw, phase_split=next_stride2)

        def plan_bwd():
            st = self.bwd
            dO = out.grad
            assert dO is not None
            # relu + GN2 backward (+ the identity / downsample branch gradient g)
            d2 = self.hlz(rows_o, Co)
            g = self.zeros(rows_o, Co) if ds is None else None
            self.gn_bwd(st, dout=dO, raw=raw2, mr=mr2, gn=gn2, N=N, HW=HWo, Cc=Co, mask_mode=1, outv=out.f32,
                        d_hl=d2, g_out=g)
            taps9 = [(kw - 1, kh - 1, 0, 0) for kh in range(3) for kw in range(3)]
            self.wgrad(st, srcs=[(a1_tw, Co, (Wo, Ho, N, 1))],
                       units=[(0, d, ch) for d in taps9 for ch in range(ops.nchunks(Co))], dy=d2, dy_channels=Co,
                       dy_dims=(Wo, Ho, N, 1), cout=Co, cin=Co, ntaps=9, param=blk.conv2.weight)
            dA1 = self.zeros(rows_o, Co)
            wd2 = self.weight(lambda: dgrad3x3_weight(blk.conv2.weight), Co, prog2.ktot, bwd=True)
            self.igemm(st, f"igemm dgrad conv2 C{Co}", srcs=[(d2, Co, prog2.src_dims[0])], taps=prog2.taps, w=wd2,
                       out_dims=prog2.out_dims, cout=Co, out_f32=dA1)
            # relu + GN1 backward
            d1 = self.hlz(rows_o, Co)
            self.probes.append(dict(d2=d2, g=g, dA1=dA1, d1=d1, raw1=raw1, raw2=raw2))
            self.gn_bwd(st, dout=dA1, raw=raw1, mr=mr1, gn=gn1, N=N, HW=HWo, Cc=Co, mask_mode=2, d_hl=d1)
            X.grad = self.zeros(X.rows, Ci)
            if s == 1:
                self.wgrad(st, srcs=[(X.tw, Ci, (X.W, X.H, N, 1))],
                           units=[(0, d, ch) for d in taps9 for ch in range(ops.nchunks(Ci))], dy=d1, dy_channels=Co,
                           dy_dims=(Wo, Ho, N, 1), cout=Co, cin=Ci, ntaps=9, param=blk.conv1.weight)
                progd = convs.spatial3x3(Co, N, Ho, Wo)
                wd1 = self.weight(lambda: dgrad3x3_weight(blk.conv1.weight), Ci, progd.ktot, bwd=True)
                self.igemm(st, f"igemm dgrad conv1 C{Co}->{Ci}", srcs=[(d1, Co, progd.src_dims[0])], taps=progd.taps,
                           w=wd1, out_dims=progd.out_dims, cout=Ci, out_f32=X.grad, residual=g)
            else:
                self.wgrad(st, srcs=[(X.tw, Ci, (X.W // 2, X.H // 2, 4, N))],
                           units=[(0, tuple(t[1]), ch) for t in prog1.taps for ch in range(ops.nchunks(Ci))], dy=d1,
                           dy_channels=Co, dy_dims=(Wo, Ho, 1, N), cout=Co, cin=Ci, ntaps=9, param=blk.conv1.weight)
                taps4 = [(0, (di, dj, 0, 0), ops.nchunks(Co)) for dj in range(2) for di in range(2)]
                wd1 = self.weight(lambda: dgrad3x3_s2_weight(blk.conv1.weight), 4 * Ci, 4 * 64 * ops.nchunks(Co), bwd=True)
                blocked = self.zeros(rows_o, 4 * Ci)
                self.igemm(st, f"igemm dgrad conv1 s2 C{Co}->4x{Ci}", srcs=[(d1, Co, (Wo, Ho, N, 1))], taps=taps4,
                           w=wd1, out_dims=(Wo, Ho, N, 1), cout=4 * Ci, out_f32=blocked)
                # downsample branch: GN backward -> weight gradient + data gradient at phase (0, 0)
                dconv = ds[0]
                dd = self.hlz(rows_o, Co)
                self.gn_bwd(st, dout=dO, raw=rawd, mr=mrd, gn=gnd, N=N, HW=HWo, Cc=Co, mask_mode=1, outv=out.f32,
                            d_hl=dd)
                self.wgrad(st, srcs=[(X.tw, Ci, (X.W // 2, X.H // 2, 4, N))],
                           units=[(0, (0, 0, 0, 0), ch) for ch in range(ops.nchunks(Ci))], dy=dd, dy_channels=Co,
                           dy_dims=(Wo, Ho, 1, N), cout=Co, cin=Ci, ntaps=1, param=dconv.weight)
                dxd = self.zeros(rows_o, Ci)
                progp = convs.pointwise(Co, (rows_o,))
                wdd = self.weight(lambda: ops.pack_weight_taps([dconv.weight.reshape(Co, Ci).t()]), Ci, progp.ktot, bwd=True)
                self.igemm(st, f"igemm dgrad downsample C{Co}->{Ci}", srcs=[(dd, Co, progp.src_dims[0])],
                           taps=progp.taps, w=wdd, out_dims=progp.out_dims, cout=Ci, out_f32=dxd)
                st.add("unblock_add", lambda: _lib.check(self.lib.v2a_enc_unblock_add(
                    blocked.data_ptr(), dxd.data_ptr(), N, X.H, X.W, Ci, X.grad.data_ptr(), ops._stream()),
                    "unblock_add"))
        tape.append(plan_bwd)
        self.acts.append(out)
        self.inner_acts.append(a1)     # post-ReLU operand of conv2 (parity probes read its sign pattern)
        return out

    def planned_launches(self) -> int:
        """kernel launches of one forward + backward (gn_bwd = 2 kernels; + the weight repack chunks)"""
        n = len(self.fwd) + len(self.bwd) + sum(1 for t in self.bwd.tags if t == "enc_gn_bwd")
        return n + len(self._wchunks) + len(self._vchunks)

    # ---- execution -----------------------------------------------------------------------------
    def _run(self, name: str, steps, pre=(), force_eager: bool = False):
        """Launch a planned list.  Every buffer is static, so after one eager (warm-up) run the list is captured
        into a CUDA graph and replayed (V2A_NO_GRAPH=1 disables).  Lane 1 = weight-gradient GEMMs + their scatter:
        nothing on the main chain reads their results, so they run on a side stream beside the data-gradient chain."""
        def eager():
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

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        self.refresh_weights()
        self.x_in.copy_(x)
        self._run("fwd", self.fwd, pre=(self.stats_arena,))
        self.fwd_token += 1
        return self.feat.clone()

    def backward(self, dfeat: torch.Tensor, clone_param_grads: bool = True):
        self.dfeat.copy_(dfeat)
        self.wait_bwd_weights()
        self._run("bwd", self.bwd, pre=(self.gslab, self.wg_arena))
        self.done.record(torch.cuda.current_stream())   # consumers of gslab on other streams wait on this
        if clone_param_grads:
            return [self.pgrad[id(p)].clone() if p.requires_grad else None for p in self.params]
        return [None] * len(self.params)


_SIDE_STREAMS: Dict[tuple, "torch.cuda.Stream"] = {}


def side_stream(device, index: int) -> "torch.cuda.Stream":
    """Per-device streams on which MultiImageObsEncoder runs all but its last encoder, so the independent
    ResNets (many launches that do not fill 148 SMs at 8x8 / 4x4) overlap, forward and -- since autograd runs a
    Function's backward on its forward stream -- backward."""
    device = torch.device(device)
    if device.index is None:
        device = torch.device("cuda", torch.cuda.current_device())
    key = (str(device), index)
    if key not in _SIDE_STREAMS:
        _SIDE_STREAMS[key] = torch.cuda.Stream(device=device)
    return _SIDE_STREAMS[key]
