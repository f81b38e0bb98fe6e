"""Table-driven weight re-layout shared by the planned engines (policy UNet1D, observation encoder).

Every "packer" is a pure gather of parameter elements (slices, transposes, zero padding, cat): it is run
ONCE on index-valued stand-ins of the parameters to record ``out[i] <- flat parameter index``, so the
re-pack after every optimiser step is one ``v2a_gather_split`` launch per arena chunk instead of hundreds
of small torch ops.  A class using the mixin provides ``device``, ``params`` (list, parameters() order),
``lib`` and calls ``_init_packing()`` before registering packers.
"""
from __future__ import annotations

from contextlib import contextmanager

import torch

from . import _lib, ops
from .ops import HL


# ---------------------------------------------------------------------------------------------------
# cache key of the packed weights
# ---------------------------------------------------------------------------------------------------
# (data_ptr, _version) per parameter catches optimiser steps made through autograd-visible in-place ops, and the
# fused optimiser invalidates explicitly.  Writes through `.data` (`p.data.copy_()`, `p.data.lerp_()`: how
# ema_pytorch.EMA.update() maintains the EMA model the reference evaluates, lb_online_trainer_v7.py:624,1077) bump
# nothing, so every INFERENCE forward also keys on a 64-bit fingerprint of the parameter values (one launch over
# the parameters + an 8-byte read-back, which synchronises).  Training forwards (grad enabled on trainable
# parameters) skip it: their optimiser either bumps the versions or invalidates explicitly.
_SCOPE = [None]
_SCOPE_SEQ = [0]


@contextmanager
def one_content_check():
    """Inside this scope every engine fingerprints its parameters at most once (a `predict_action` call runs the
    same UNet1D engine 8 times: the weights cannot change in between)."""
    _SCOPE_SEQ[0] += 1
    prev, _SCOPE[0] = _SCOPE[0], _SCOPE_SEQ[0]
    try:
        yield
    finally:
        _SCOPE[0] = prev


def training_call(params) -> bool:
    """True when the caller runs a TRAINING forward (grad mode on and something to train).  Must be evaluated in the
    module's forward, not inside a torch.autograd.Function.forward (grad mode is always off in there)."""
    return torch.is_grad_enabled() and any(p.requires_grad for p in params)


def content_key(engine, params):
    """Fingerprint part of an engine's weight-cache key; engines set ``engine.training_call`` per forward (engines
    that never train -- the video UNet -- leave it False and always check)."""
    if getattr(engine, "training_call", False) or not all(p.is_cuda for p in params):
        return getattr(engine, "_last_fp", None)
    scope = _SCOPE[0]
    if scope is not None and getattr(engine, "_fp_scope", None) == scope:
        return engine._last_fp
    engine._last_fp = ops.params_fingerprint(params)
    engine._fp_scope = scope
    return engine._last_fp


class PackedParams:
    def _init_packing(self):
        self.packers, self.vec_packers = [], []
        self._wchunks, self._vchunks = [], []
        self._wkey = None
        self._stage = None
        self._trace_checked = False
        self._pack_stream = None          # side stream for the backward-only packs (data-gradient weights)
        self._bwd_pack_done = None        # event the backward waits on

    def _carve(self, kind: str, n: int, bwd: bool = False):
        """n elements (multiple of 128) from the growing weight ('w': bf16 hi/lo planes) or vector ('v': fp32)
        arena; every arena chunk is re-packed by ONE v2a_gather_split launch."""
        n_al = -(-n // 128) * 128
        # 'h': weight planes in the fp16 (hi, lo) format (stored in bf16-typed tensors; only the bits matter)
        chunks = ([c for c in self._wchunks if c["fmt"] == (1 if kind == "h" else 0) and c["bwd"] == bwd]
                  if kind in "wh" else self._vchunks)
        if not chunks or chunks[-1]["used"] + n_al > chunks[-1]["cap"]:
            cap = max(n_al, (16 << 20) if kind in "wh" else (1 << 20))
            c = dict(cap=cap, used=0, map=torch.zeros(cap, dtype=torch.int32, device=self.device),
                     fmt=1 if kind == "h" else 0, bwd=bwd)
            if kind in "wh":
                c["hi"] = torch.zeros(cap, dtype=torch.bfloat16, device=self.device)
                c["lo"] = torch.zeros(cap, dtype=torch.bfloat16, device=self.device)
            else:
                c["f32"] = torch.zeros(cap, dtype=torch.float32, device=self.device)
            (self._wchunks if kind in "wh" else self._vchunks).append(c)
            chunks = [c]
        c = chunks[-1]
        lo, hi = c["used"], c["used"] + n
        c["used"] += n_al
        return c, lo, hi

    def weight(self, fn, rows, cols, fp16: bool = False, bwd: bool = False) -> HL:
        """bwd=True: a pack only the backward pass reads (data-gradient layouts) -- re-packed on a side stream
        while the forward pass runs (`wait_bwd_weights()` before the first backward launch)."""
        c, lo, hi = self._carve("h" if fp16 else "w", rows * cols, bwd)
        hl = HL(c["hi"][lo:hi].view(rows, cols), c["lo"][lo:hi].view(rows, cols))
        hl.fp16 = fp16
        self.packers.append((fn, hl, c["map"][lo:hi].view(rows, cols)))
        return hl

    def vec(self, fn, n) -> torch.Tensor:
        c, lo, hi = self._carve("v", n)
        v = c["f32"][lo:hi]
        self.vec_packers.append((fn, v, c["map"][lo:hi]))
        return v

    def _trace_packers(self):
        """Every packer is a pure gather of parameter elements (slices, transposes, zero padding, cat):
        run each ONCE on index-valued stand-ins of the parameters to record out[i] <- flat parameter index
        (+1; 0 = structural zero).  Re-packing after an optimiser step is then table driven."""
        saved = [p.data for p in self.params]
        try:
            off = 0
            for p in self.params:
                n = p.numel()
                p.data = torch.arange(off + 1, off + n + 1, dtype=torch.float64, device=self.device).view(p.shape)
                off += n
            assert off < 2 ** 31 - 1
            with torch.no_grad():
                for fn, hl, mp in self.packers:
                    mp.copy_(fn().to(self.device).reshape(mp.shape).to(torch.int32))
                for fn, v, mp in self.vec_packers:
                    mp.copy_(fn().to(self.device).reshape(-1).to(torch.int32))
        finally:
            for p, d in zip(self.params, saved):
                p.data = d
        self._stage = None
        self._trace_checked = False

    def _param_slab(self) -> torch.Tensor:
        """The parameters as ONE flat fp32 tensor in parameters() order: the live slab when they already are
        consecutive views of one (train_step.PolicyTrainStep), else a staging copy (one cat launch)."""
        ps = [p for p in self.params if p.numel()]
        ptr, ok = ps[0].data_ptr(), True
        for p in ps:
            if p.data_ptr() != ptr or not p.is_contiguous() or p.dtype != torch.float32 or p.device != self.device:
                ok = False
                break
            ptr += 4 * p.numel()
        tot = sum(p.numel() for p in ps)
        if ok:
            st = ps[0].untyped_storage()
            first = (ps[0].data_ptr() - st.data_ptr()) // 4
            return torch.empty(0, dtype=torch.float32, device=self.device).set_(st, first, (tot,))
        if self._stage is None:
            self._stage = torch.empty(tot, dtype=torch.float32, device=self.device)
        with torch.no_grad():
            torch.cat([p.detach().to(self.device, torch.float32).reshape(-1) for p in ps], out=self._stage)
        return self._stage

    def _repack_reference(self):
        """The packers evaluated with torch ops on the real parameters (one-off check of the traced tables)."""
        with torch.no_grad():
            for fn, hl, _ in self.packers:
                w = fn().detach().to(self.device, torch.float32).reshape(hl.hi.shape)
                if getattr(hl, "fp16", False):
                    hi = w.to(torch.float16)
                    yield hl.hi.view(torch.float16), hi
                    yield hl.lo.view(torch.float16), (w - hi.float()).to(torch.float16)
                    continue
                hi = w.to(torch.bfloat16)
                yield hl.hi, hi
                yield hl.lo, (w - hi.float()).to(torch.bfloat16)
            for fn, v, _ in self.vec_packers:
                yield v, fn().detach().to(self.device, torch.float32).reshape(-1)

    def refresh_weights(self):
        key = (tuple((p.data_ptr(), p._version) for p in self.params), content_key(self, self.params))
        if key == self._wkey:
            return
        src = self._param_slab()
        st = ops._stream()

        def pack(c, stream):
            _lib.check(self.lib.v2a_gather_split_fmt(src.data_ptr(), c["map"].data_ptr(), c["used"], c["hi"].data_ptr(),
                                                     c["lo"].data_ptr(), None, c["fmt"], stream), "gather_split")
        for c in self._wchunks:
            if not c["bwd"]:
                pack(c, st)
        for c in self._vchunks:
            _lib.check(self.lib.v2a_gather_split(src.data_ptr(), c["map"].data_ptr(), c["used"], None, None,
                                                 c["f32"].data_ptr(), st), "gather_split")
        bwd_chunks = [c for c in self._wchunks if c["bwd"]]
        if bwd_chunks:
            # nothing in the forward pass reads these: re-pack them beside it
            cur = torch.cuda.current_stream()
            if self._pack_stream is None:
                self._pack_stream = torch.cuda.Stream(device=self.device)
            self._pack_stream.wait_stream(cur)
            with torch.cuda.stream(self._pack_stream):
                for c in bwd_chunks:
                    pack(c, self._pack_stream.cuda_stream)
                self._bwd_pack_done = torch.cuda.Event()
                self._bwd_pack_done.record(self._pack_stream)
            if not self._trace_checked:
                cur.wait_stream(self._pack_stream)
        if not self._trace_checked:   # first pack: the traced gather must reproduce the torch packers bit for bit
            for got, want in self._repack_reference():
                if not torch.equal(got, want):
                    raise RuntimeError("v2a_b200 engine: traced weight table disagrees with its packer")
            self._trace_checked = True
        self._wkey = key


    def wait_bwd_weights(self) -> None:
        """Order the current stream after the side-stream re-pack of the backward-only weight layouts."""
        if self._bwd_pack_done is not None:
            torch.cuda.current_stream().wait_event(self._bwd_pack_done)
