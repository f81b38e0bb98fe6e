"""Fused optimisation step of the policy training loop.

Replaces, for a module whose hot part is ``ConditionalUnet1D``, what the reference trainer does
each iteration (diffuser/libero/lb_online_trainer_v7.py:593-624; hyper-parameters from
config/libero/lb_tk8_65to72.py:138-153):

    loss = compute_loss(batch); loss.backward()          -> planned CUDA forward / backward
    [DDP gradient all-reduce under multi-process]        -> distributed.allreduce_mean_ (NCCL)
    clip_grad_norm_(params, 1.0)                         -> v2a_grad_sumsq
    AdamW.step(); zero_grad()                            -> v2a_adamw_ema_step (ONE pass over flat
    EMA.update()                                            slabs: p, g, m, v, ema)

Parameters are re-pointed into one flat fp32 slab per segment (their ``state_dict`` names and
shapes are unchanged); the gradients of the UNet1D and of the observation encoders never leave their
engines' gradient slabs, any other parameter accumulates through autograd into slab views.

EMA follows ``ema_pytorch`` 0.2.3 as the trainer configures it (``EMA(model, beta=0.9999,
update_after_step=0, inv_gamma=1, power=0.75, min_value=0, update_every=1)``; third-party, not
vendored in the reference — restated from its published algorithm, parity unpinned by reference
tests): the first two ``update()`` calls copy the online parameters, call c >= 2 uses
decay = clamp(1 - (1 + c)^-0.75, 0, beta).
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Callable, Iterable, List, Optional

import torch
import torch.nn as nn

from . import _lib, distributed, obs_encoder, ops, policy_unet1d


def ema_decay(call_index: int, *, beta: float = 0.9999, inv_gamma: float = 1.0, power: float = 0.75,
              min_value: float = 0.0, update_after_step: int = 0) -> float:
    """Decay used by the ``call_index``-th (0-based) ``EMA.update()`` call; 0.0 means 'copy online params'."""
    if call_index <= update_after_step + 1:
        return 0.0
    epoch = call_index - update_after_step
    value = 1.0 - (1.0 + epoch / inv_gamma) ** (-power)
    return float(min(max(value, min_value), beta))


class _Segment:
    """A group of parameters living in one flat slab, with optimiser state beside it."""

    def __init__(self, params: List[nn.Parameter], device, own_grad: bool, ema: bool):
        self.params = params
        self.numel = sum(p.numel() for p in params)
        f32 = dict(dtype=torch.float32, device=device)
        self.p = torch.empty(self.numel, **f32)
        self.m = torch.zeros(self.numel, **f32)
        self.v = torch.zeros(self.numel, **f32)
        self.g: Optional[torch.Tensor] = torch.zeros(self.numel, **f32) if own_grad else None
        off = 0
        with torch.no_grad():
            for p in params:
                n = p.numel()
                view = self.p[off:off + n].view(p.shape)
                view.copy_(p.data)
                p.data = view
                if own_grad:
                    p.grad = self.g[off:off + n].view(p.shape)
                off += n
        self.ema = self.p.clone() if ema else None

    def ema_views(self):
        off = 0
        for p in self.params:
            n = p.numel()
            yield p, self.ema[off:off + n].view(p.shape)
            off += n


# ---------------------------------------------------------------------------------------------------
# checkpoint / resume: the optimiser and EMA state in the formats the reference trainer saves
# (lb_online_trainer_v7.py:367-407: data['opt'] = AdamW.state_dict(), data['ema'] = EMA.state_dict())
# ---------------------------------------------------------------------------------------------------
def export_optimizer_state(module: nn.Module, segments: Iterable["_Segment"], steps_done: int, *, lr: float, betas,
                           eps: float, weight_decay: float) -> dict:
    """``torch.optim.AdamW(module.parameters()).state_dict()`` equivalent of the slab state: per-parameter
    ``exp_avg`` / ``exp_avg_sq`` / ``step`` indexed by position in ``module.parameters()``."""
    index = {id(p): i for i, p in enumerate(module.parameters())}
    state = {}
    if steps_done > 0:                                   # torch creates the state lazily at the first step
        for seg in segments:
            off = 0
            for p in seg.params:
                n = p.numel()
                if n:
                    state[index[id(p)]] = {"step": torch.tensor(float(steps_done)),
                                           "exp_avg": seg.m[off:off + n].view(p.shape).clone(),
                                           "exp_avg_sq": seg.v[off:off + n].view(p.shape).clone()}
                off += n
    group = {"lr": lr, "betas": tuple(betas), "eps": eps, "weight_decay": weight_decay, "amsgrad": False,
             "maximize": False, "foreach": None, "capturable": False, "differentiable": False, "fused": None,
             "params": list(range(len(index)))}
    return {"state": dict(sorted(state.items())), "param_groups": [group]}


def import_optimizer_state(module: nn.Module, segments: Iterable["_Segment"], opt_state: dict) -> int:
    """Inverse of ``export_optimizer_state`` (accepts a real ``AdamW.state_dict()``); returns the step count."""
    index = {id(p): i for i, p in enumerate(module.parameters())}
    state = opt_state["state"]
    steps = 0
    with torch.no_grad():
        for seg in segments:
            off = 0
            for p in seg.params:
                n = p.numel()
                st = state.get(index[id(p)])
                if n and st is not None:
                    seg.m[off:off + n].copy_(st["exp_avg"].reshape(-1))
                    seg.v[off:off + n].copy_(st["exp_avg_sq"].reshape(-1))
                    steps = max(steps, int(float(st["step"])))
                elif n:
                    seg.m[off:off + n].zero_()
                    seg.v[off:off + n].zero_()
                off += n
    return steps


def export_ema_state(module: nn.Module, segments: Iterable["_Segment"], steps_done: int) -> dict:
    """``ema_pytorch.EMA(include_online_model=False).state_dict()`` layout (the trainer's setting,
    config/libero/lb_tk8_65to72.py:145-153: the online model is then held in a plain list and contributes no
    ``online_model.*`` keys): ``ema_model.<name>`` for every parameter (from the EMA slabs) and buffer (from the
    module), plus the two bookkeeping buffers as ema_pytorch 0.2.3 registers them -- ``initted`` =
    ``torch.Tensor([bool])`` (float32, shape [1]; set by the SECOND ``update()`` call when ``update_after_step``
    is 0) and ``step`` = ``torch.tensor([n])`` (int64, shape [1]).  The package is not installed offline: restated
    from its published source, unpinned by a round trip through the real class."""
    names = {id(p): n for n, p in module.named_parameters()}
    out = {}
    for seg in segments:
        if seg.ema is None:
            raise RuntimeError("no EMA state: PolicyTrainStep was built with ema=False")
        for p, e in seg.ema_views():
            out["ema_model." + names[id(p)]] = e.clone()
    for n, p in module.named_parameters():                      # zero-size placeholders etc.
        out.setdefault("ema_model." + n, p.detach().clone())
    for n, b in module.named_buffers():
        out["ema_model." + n] = b.detach().clone()
    out["initted"] = torch.Tensor([steps_done >= 2])
    out["step"] = torch.tensor([steps_done])
    return out


def import_ema_state(module: nn.Module, segments: Iterable["_Segment"], ema_state: dict) -> None:
    names = {id(p): n for n, p in module.named_parameters()}
    with torch.no_grad():
        for seg in segments:
            if seg.ema is None:
                continue
            for p, e in seg.ema_views():
                key = "ema_model." + names[id(p)]
                if key not in ema_state:
                    raise KeyError(f"EMA checkpoint lacks {key}")
                e.copy_(ema_state[key])


class PolicyTrainStep:
    """``step(loss_fn)`` = forward + backward + (all-reduce) + clip + AdamW + EMA on CUDA."""

    def __init__(self, module: nn.Module, *, unet1d: Optional[nn.Module] = None, lr: float = 1e-4,
                 betas=(0.95, 0.999), eps: float = 1e-8, weight_decay: float = 1e-6, max_norm: float = 1.0,
                 ema: bool = True, ema_beta: float = 0.9999, ema_power: float = 0.75, ema_inv_gamma: float = 1.0,
                 group=None, bucket_bytes: int = 64 << 20):
        if unet1d is None:
            unet1d = module if isinstance(module, policy_unet1d.ConditionalUnet1D) else getattr(module, "model", None)
        if not isinstance(unet1d, policy_unet1d.ConditionalUnet1D):
            raise TypeError("PolicyTrainStep needs a v2a_b200 ConditionalUnet1D (module itself or module.model)")
        dev = next(unet1d.parameters()).device
        if dev.type != "cuda":
            raise RuntimeError("PolicyTrainStep runs on CUDA only (no CPU fallback): move the module first")
        self.module, self.unet = module, unet1d
        self.lr, self.betas, self.eps, self.wd, self.max_norm = lr, tuple(betas), eps, weight_decay, max_norm
        self.ema_cfg = dict(beta=ema_beta, power=ema_power, inv_gamma=ema_inv_gamma)
        self.group, self.bucket_bytes = group, bucket_bytes
        unet_params = list(unet1d.parameters())
        ids = {id(p) for p in unet_params}
        # observation encoders that run on the planned CUDA engine keep their gradients in the engine's slab too
        from .diffusion_policy import VisualCore
        self.cores = [m for m in module.modules() if isinstance(m, VisualCore)]
        self.seg_cores = []
        for core in self.cores:
            cp = [p for p in core.parameters()]
            ids |= {id(p) for p in cp}
            self.seg_cores.append(_Segment(cp, dev, own_grad=False, ema=ema))
            obs_encoder.set_slab_grads(core, True)
        other = [p for p in module.parameters() if id(p) not in ids and p.requires_grad and p.numel() > 0]
        self.seg_unet = _Segment(unet_params, dev, own_grad=False, ema=ema)
        self.seg_other = _Segment(other, dev, own_grad=True, ema=ema) if other else None
        policy_unet1d.set_slab_grads(unet1d, True)
        self.sumsq = torch.zeros(1, dtype=torch.float64, device=dev)
        self.steps_done = 0
        self.collectives = 0
        self._early = {}           # id(gradient slab) -> pending collectives started from inside backward
        # Multi-GPU: every engine's gradient slab starts its all-reduce the moment that engine's backward has been
        # enqueued (autograd runs the UNet1D's backward first, then the two encoders'), so the exchange travels
        # under the rest of backward; only the last encoder's 45 MB is exposed.  Graph replay joins the
        # weight-gradient side lane before the graph ends, so a slab is complete (in stream order) when its
        # backward returns; the eager-launch debugging mode gives no such guarantee.  V2A_OVERLAP_ALLREDUCE=0 = one
        # blocking exchange after backward (A/B probe).
        self._overlap = os.environ.get("V2A_OVERLAP_ALLREDUCE", "1") != "0" and not os.environ.get("V2A_NO_GRAPH")
        self._lib = _lib.load()

    def close(self) -> None:
        """Hand gradients back to autograd (per-parameter .grad tensors) for anyone who keeps using the modules
        without this object."""
        try:
            policy_unet1d.set_slab_grads(self.unet, False)
            for core in self.cores:
                obs_encoder.set_slab_grads(core, False)
        except Exception:
            pass

    def __del__(self):
        self.close()

    # ---- pieces (also used one by one in tests) ------------------------------------------------
    def _unet_grad_slab(self) -> torch.Tensor:
        eng = policy_unet1d.last_engine(self.unet)
        if eng is None:
            raise RuntimeError("PolicyTrainStep.step: the loss closure did not run the ConditionalUnet1D")
        return eng.gslab

    def _segments(self):
        yield self.seg_unet, self._unet_grad_slab()
        for core, seg in zip(self.cores, self.seg_cores):
            eng = obs_encoder.last_engine(core)
            if eng is None:
                raise RuntimeError("PolicyTrainStep.step: the loss closure did not run an observation encoder")
            torch.cuda.current_stream().wait_event(eng.done)   # its backward may have run on a side stream
            yield seg, eng.gslab
        if self.seg_other is not None:
            yield self.seg_other, self.seg_other.g

    def optimizer_tail(self) -> None:
        """all-reduce (mean) -> global grad-norm -> clip + AdamW + EMA, all asynchronous on the stream."""
        st = ops._stream()
        segs = list(self._segments())
        early, self._early = self._early, {}
        works = []
        for _, g in segs:          # same slab order on every rank; slabs already in flight are not started twice
            w = early.pop(g.data_ptr(), None)
            works += w if w is not None else distributed.allreduce_mean_start([g], self.group, self.bucket_bytes)
        self.collectives += distributed.allreduce_wait(works)
        self.sumsq.zero_()
        for _, g in segs:
            _lib.check(self._lib.v2a_grad_sumsq(g.data_ptr(), g.numel(), self.sumsq.data_ptr(), st), "grad_sumsq")
        decay = ema_decay(self.steps_done, **self.ema_cfg)
        self.steps_done += 1
        for seg, g in segs:
            _lib.check(self._lib.v2a_adamw_ema_step(
                seg.p.data_ptr(), g.data_ptr(), seg.m.data_ptr(), seg.v.data_ptr(),
                None if seg.ema is None else seg.ema.data_ptr(), seg.numel, self.sumsq.data_ptr(),
                self.max_norm, self.lr, self.betas[0], self.betas[1], self.eps, self.wd, self.steps_done,
                decay, st), "adamw_ema_step")
        policy_unet1d.invalidate_weights(self.unet)
        for core in self.cores:
            obs_encoder.invalidate_weights(core)
        if self.seg_other is not None:
            self.seg_other.g.zero_()          # zero_grad() for the autograd-accumulated segment

    def step(self, loss_fn: Callable[[], torch.Tensor]) -> torch.Tensor:
        """Run ``loss_fn`` (it must call the module), backward, and the optimiser tail.  Returns the loss
        tensor (device resident; reading it synchronises)."""
        loss = loss_fn()
        if self._overlap and distributed.world_size(self.group) > 1:
            engines = [policy_unet1d.last_engine(self.unet)] + [obs_encoder.last_engine(c) for c in self.cores]
            for eng in engines:
                if eng is not None:
                    def start(eng=eng):
                        self._early[eng.gslab.data_ptr()] = distributed.allreduce_mean_start(
                            [eng.gslab], self.group, self.bucket_bytes)
                    eng.on_backward_done = start
        loss.backward()
        self.optimizer_tail()
        return loss.detach()

    # ---- checkpoint / resume (lb_online_trainer_v7.py:367-407) ----------------------------------
    def _all_segments(self):
        return [seg for seg in (self.seg_unet, *self.seg_cores, self.seg_other) if seg is not None]

    def state_dict(self) -> dict:
        """``{'opt': AdamW-format, 'ema': ema_pytorch-format, 'steps_done': int}`` — what the reference trainer
        stores under ``data['opt']`` / ``data['ema']``; loadable into the stock ``AdamW`` / ``EMA`` objects."""
        segs = self._all_segments()
        out = {"opt": export_optimizer_state(self.module, segs, self.steps_done, lr=self.lr, betas=self.betas,
                                             eps=self.eps, weight_decay=self.wd),
               "steps_done": self.steps_done}
        if self.seg_unet.ema is not None:
            out["ema"] = export_ema_state(self.module, segs, self.steps_done)
        return out

    def load_state_dict(self, state: dict) -> None:
        """Resume from ``state_dict()`` output or from the reference trainer's ``data['opt']`` / ``data['ema']``
        (pass ``{'opt': ..., 'ema': ...}``).  The module's own weights are loaded separately, as in the trainer."""
        segs = self._all_segments()
        steps = import_optimizer_state(self.module, segs, state["opt"])
        self.steps_done = int(state.get("steps_done", steps))
        if "ema" in state and self.seg_unet.ema is not None:
            import_ema_state(self.module, segs, state["ema"])
        policy_unet1d.invalidate_weights(self.unet)
        for core in self.cores:
            obs_encoder.invalidate_weights(core)

    def grad_norm(self) -> torch.Tensor:
        """Global gradient L2 norm of the last step (before clipping), as a device tensor."""
        return self.sumsq.sqrt()

    # ---- EMA export --------------------------------------------------------------------------
    @torch.no_grad()
    def copy_ema_to(self, target: nn.Module) -> None:
        """Write the EMA weights into ``target`` (same architecture), e.g. the trainer's ``ema.ema_model``."""
        if self.seg_unet.ema is None:
            raise RuntimeError("PolicyTrainStep was built with ema=False")
        src = {}
        names = {id(p): n for n, p in self.module.named_parameters()}
        for seg in (self.seg_unet, *self.seg_cores, self.seg_other):
            if seg is None:
                continue
            for p, e in seg.ema_views():
                src[names[id(p)]] = e
        for n, p in target.named_parameters():
            if n in src:
                p.copy_(src[n])          # in-place through autograd's version counter (under no_grad)
        # the target's engines key their packed weights on (data_ptr, _version) + a content fingerprint at
        # inference; drop the caches explicitly as well so the very next forward repacks
        from .diffusion_policy import VisualCore
        for m in target.modules():
            if isinstance(m, policy_unet1d.ConditionalUnet1D):
                policy_unet1d.invalidate_weights(m)
            elif isinstance(m, VisualCore):
                obs_encoder.invalidate_weights(m)
