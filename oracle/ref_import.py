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

def _default_root() -> str:
    """/root/reference in the build container; on the GPU box the copy `oracle/make_ref.py` ships under
    baseline/_ref (git-ignored, travels with the gpurun snapshot)."""
    if os.path.isdir("/root/reference/flowdiffusion/flowdiffusion"):
        return "/root/reference"
    return os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "baseline", "_ref")


REF_ROOT = os.environ.get("V2A_REFERENCE_ROOT") or _default_root()


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


def build_reference_policy():
    """The reference's own `DiffusionUnetImagePolicy` with the Libero yaml's settings
    (config/diff_policy/lb_train_diffusion_unet_image_orn10.yaml).  `diffusers` is absent offline: the class
    receives the restated scheduler objects of v2a_b200.diffusion_policy (third-party boundary)."""
    install_shims()
    from v2a_b200 import diffusion_policy as DP
    pol = importlib.import_module("diffuser.diffusion_policy.diffusion_unet_image_policy")
    moe = importlib.import_module("diffuser.diffusion_policy.model.multi_image_obs_encoder")
    vn = importlib.import_module("diffuser.diffusion_policy.common.vision_nets")
    meta = DP.libero_shape_meta()
    core = vn.VisualCore(input_shape=[3, 128, 128], backbone_class="ResNet18Conv",
                         backbone_kwargs=dict(pretrained=None, input_coord_conv=False), pool_class="SpatialSoftmax",
                         pool_kwargs=dict(num_kp=32, learnable_temperature=False, temperature=1.0, noise_std=0.0,
                                          output_variance=False), flatten=True, feature_dimension=64)
    enc = moe.MultiImageObsEncoder(meta, core, use_group_norm=True)
    sched = dict(num_train_timesteps=100, beta_start=0.0001, beta_end=0.02, beta_schedule="squaredcos_cap_v2",
                 clip_sample=True, prediction_type="epsilon")
    return pol.DiffusionUnetImagePolicy(meta, DP.DDPMScheduler(**sched), DP.DDIMScheduler(**sched), enc, horizon=16,
                                        n_action_steps=8, n_obs_steps=1, num_inference_steps=100,
                                        diffusion_step_embed_dim=128, down_dims=[256, 512, 1024], kernel_size=5,
                                        n_groups=8, cond_predict_scale=True)


def build_reference_video_diffusion(timesteps: int = 100, sampling_timesteps: int = 100, frames: int = 7,
                                     image_size=(128, 128)):
    """The reference's own GoalGaussianDiffusion(Unet_Libero()) at the shipped Libero settings
    (diffuser/models/video_model.py:15-40: pred_v, cosine schedule, guidance off)."""
    UL, G = Unet_Libero(), GoalGaussianDiffusion()
    return G(UL(), image_size=tuple(image_size), channels=3 * frames, timesteps=timesteps,
             sampling_timesteps=sampling_timesteps, loss_type="l2", objective="pred_v", beta_schedule="cosine",
             min_snr_loss_weight=True, guidance_weight=0)
