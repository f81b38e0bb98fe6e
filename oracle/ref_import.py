"""TEST INFRASTRUCTURE — imports the UNMODIFIED reference modules from /root/reference.

Only usable in the build container (the GPU box has no /root/reference).  Used
by tests/golden/make_golden.py to generate the committed golden vectors and by
the CPU tests that pin oracle/*.py against the real reference.  Nothing on the
product path imports this file.

The reference's hot-path modules import many unrelated packages at module top
(accelerate, ema_pytorch, matplotlib, gym, mujoco_py, einops_exts, diffusers,
omegaconf ...) that are not installed here; they are replaced by inert stubs
in sys.modules BEFORE the import.  No reference source is
copied: the real files are executed from where they lie.
"""
from __future__ import annotations

import copy
import importlib
import os
import sys
import types

REF_ROOT = os.environ.get("V2A_REFERENCE_ROOT", "/root/reference")


def available() -> bool:
    return os.path.isdir(os.path.join(REF_ROOT, "flowdiffusion", "flowdiffusion"))


def _stub(name: str, **attrs) -> types.ModuleType:
    m = sys.modules.get(name)
    if m is None:
        m = types.ModuleType(name)
        m.__dict__["__stub__"] = True
        sys.modules[name] = m
    for k, v in attrs.items():
        setattr(m, k, v)
    return m


def _namespace(name: str, path: str) -> None:
    if name in sys.modules:
        return
    m = types.ModuleType(name)
    m.__path__ = [path]
    sys.modules[name] = m


_installed = False


def install_shims() -> None:
    global _installed
    if _installed:
        return
    if not available():
        raise RuntimeError(f"reference checkout not found at {REF_ROOT}")
    from einops import rearrange, repeat

    # einops_exts==0.0.4: *_many are maps of the einops functions
    _stub("einops_exts",
          rearrange_many=lambda ts, pattern, **kw: [rearrange(t, pattern, **kw) for t in ts],
          repeat_many=lambda ts, pattern, **kw: [repeat(t, pattern, **kw) for t in ts],
          check_shape=lambda t, pattern, **kw: t)

    class _EinopsToAndFrom:  # only referenced at class-definition time off the hot path
        def __init__(self, *a, **k):
            raise NotImplementedError

    _stub("einops_exts.torch", EinopsToAndFrom=_EinopsToAndFrom)

    class _EMA:  # holder with the attribute the wrappers touch
        def __init__(self, model, *a, **k):
            self.ema_model = copy.deepcopy(model)

    _stub("ema_pytorch", EMA=_EMA)
    _stub("accelerate", Accelerator=type("Accelerator", (), {}))
    _stub("matplotlib")
    _stub("matplotlib.pyplot")
    _stub("gym")
    _stub("mujoco_py", MjSimState=type("MjSimState", (), {}))
    _stub("omegaconf", OmegaConf=type("OmegaConf", (), {}))
    _stub("imageio")
    _stub("termcolor", colored=lambda s, *a, **k: s, cprint=print)
    _stub("h5py")
    # diffusers schedulers: the policy oracle supplies its own restatement (policy_oracle.py)
    _stub("diffusers")
    _stub("diffusers.schedulers")
    _stub("diffusers.schedulers.scheduling_ddpm", DDPMScheduler=type("DDPMScheduler", (), {}))
    _stub("diffusers.schedulers.scheduling_ddim", DDIMScheduler=type("DDIMScheduler", (), {}))

    # namespace packages: bypass the reference's __init__.py files (they pull tap/h5py/omegaconf)
    d = os.path.join(REF_ROOT, "diffuser")
    _namespace("diffuser", d)
    for sub in ("utils", "models", "diffusion_policy", "datasets", "libero"):
        _namespace(f"diffuser.{sub}", os.path.join(d, sub))
    for sub in ("model", "common"):
        _namespace(f"diffuser.diffusion_policy.{sub}", os.path.join(d, "diffusion_policy", sub))
    f = os.path.join(REF_ROOT, "flowdiffusion")
    _namespace("flowdiffusion", f)
    _namespace("flowdiffusion.flowdiffusion", os.path.join(f, "flowdiffusion"))
    _installed = True


def _imp(name: str):
    install_shims()
    return importlib.import_module(name)


def unet_module():
    return _imp("flowdiffusion.flowdiffusion.guided_diffusion.guided_diffusion.unet")


def UNetModel():
    return unet_module().UNetModel


def Unet_Libero():
    return _imp("flowdiffusion.flowdiffusion.unet").Unet_Libero


def GoalGaussianDiffusion():
    return _imp("flowdiffusion.flowdiffusion.goal_diffusion").GoalGaussianDiffusion


def ConditionalUnet1D():
    return _imp("diffuser.diffusion_policy.model.conditional_unet1d").ConditionalUnet1D


def replay_buffer_module():
    """diffuser/datasets/env_img_replay_buffer.py (row N4); its simulator import is an inert stub."""
    install_shims()
    _stub("environment")
    _stub("environment.libero")
    _stub("environment.libero.lb_env_v3", LiberoEnvList_V3=type("LiberoEnvList_V3", (), {}))
    return _imp("diffuser.datasets.env_img_replay_buffer")


def img_utils_module():
    return _imp("diffuser.datasets.img_utils")
