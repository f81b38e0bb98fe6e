"""Drop the B200 paths into an UNMODIFIED reference checkout.

``install()`` rebinds, inside the reference's own modules, the classes on the two hot paths:

  flowdiffusion.flowdiffusion.goal_diffusion.GoalGaussianDiffusion   -> v2a_b200.GoalGaussianDiffusion
  flowdiffusion.flowdiffusion.unet.Unet_Libero                       -> v2a_b200.Unet_Libero
  diffuser.diffusion_policy.model.conditional_unet1d.ConditionalUnet1D
  (+ the name imported into diffusion_unet_image_policy)             -> v2a_b200.ConditionalUnet1D
  diffuser.diffusion_policy.common.vision_nets.VisualCore            -> v2a_b200.diffusion_policy.VisualCore
  diffuser.diffusion_policy.model.multi_image_obs_encoder.MultiImageObsEncoder
                                                                     -> v2a_b200.diffusion_policy.MultiImageObsEncoder
  diffuser.datasets.env_img_replay_buffer.Global_EnvReplayBuffer_Img
  (+ the name imported into lb_online_trainer_v7)                    -> v2a_b200.replay.Global_EnvReplayBuffer_Img
                                                                        (episodes in HBM, row N4; ``replay=False`` skips it)

so ``lb_get_video_model_gcp_v2`` (diffuser/libero/lb_video_model_utils.py:13-66) and
``Init_Diffusion_Policy`` (diffuser/diffusion_policy/get_dp.py:27-89) build the CUDA-backed modules
while ``scripts/train_libero_dp.py`` and the trainer stay untouched.  The replacements keep the
reference's constructor signatures and ``state_dict`` layout, so its checkpoints load with
``strict=True`` and ``ema_pytorch.EMA`` can deep-copy them.
"""
from __future__ import annotations

import importlib
import sys
from typing import Dict, List, Tuple

_TARGETS: List[Tuple[str, str, str]] = [
    # (reference module, attribute, our module)
    ("flowdiffusion.flowdiffusion.goal_diffusion", "GoalGaussianDiffusion", "goal_diffusion"),
    ("flowdiffusion.flowdiffusion.unet", "Unet_Libero", "unet"),
    ("diffuser.libero.lb_video_model_utils", "GoalGaussianDiffusion", "goal_diffusion"),
    ("diffuser.libero.lb_video_model_utils", "Unet_Libero", "unet"),
    ("diffuser.diffusion_policy.model.conditional_unet1d", "ConditionalUnet1D", "policy_unet1d"),
    ("diffuser.diffusion_policy.diffusion_unet_image_policy", "ConditionalUnet1D", "policy_unet1d"),
    # observation encoder (row P6 / N1): the yaml instantiates both by `_target_` path, i.e. by these names
    ("diffuser.diffusion_policy.common.vision_nets", "VisualCore", "diffusion_policy"),
    ("diffuser.diffusion_policy.model.multi_image_obs_encoder", "MultiImageObsEncoder", "diffusion_policy"),
]
# replay buffer with HBM-resident episodes (row N4): the trainer touches only add_one_episode /
# sample_random_batch_seq / len (lb_online_trainer_v7.py:774,797-837,927,975-977)
_REPLAY_TARGETS: List[Tuple[str, str, str]] = [
    ("diffuser.datasets.env_img_replay_buffer", "Global_EnvReplayBuffer_Img", "replay"),
    ("diffuser.libero.lb_online_trainer_v7", "Global_EnvReplayBuffer_Img", "replay"),
]
_saved: Dict[Tuple[str, str], object] = {}


def install(import_missing: bool = True, replay: bool = True) -> List[str]:
    """Rebind the hot-path classes; returns the ``module.attr`` names that were patched.

    Modules the reference has not imported yet are imported first when ``import_missing`` (so call this
    after ``sys.path`` contains the reference checkout and before the models are constructed); modules that
    cannot be imported (e.g. the Libero helpers without the simulator installed) are skipped.
    """
    patched = []
    for mod_name, attr, ours in _TARGETS + (_REPLAY_TARGETS if replay else []):
        mod = sys.modules.get(mod_name)
        if mod is None and import_missing:
            try:
                mod = importlib.import_module(mod_name)
            except Exception:
                mod = None
        if mod is None or not hasattr(mod, attr):
            continue
        new = getattr(importlib.import_module(f"{__package__}.{ours}"), attr)
        cur = getattr(mod, attr)
        if cur is new:
            continue
        _saved.setdefault((mod_name, attr), cur)
        setattr(mod, attr, new)
        patched.append(f"{mod_name}.{attr}")
    return patched


def uninstall() -> None:
    """Restore the reference's own classes."""
    for (mod_name, attr), old in list(_saved.items()):
        mod = sys.modules.get(mod_name)
        if mod is not None:
            setattr(mod, attr, old)
        del _saved[(mod_name, attr)]
