"""Generate the committed golden vectors by running the UNMODIFIED reference modules.

Run in the build container only (needs /root/reference):
    python tests/golden/make_golden.py
Weights are synthetic but deterministic (oracle.video_oracle.seeded_state_dict from the
reference's own state_dict shapes), inputs are seeded, everything fp32 on CPU with a
fixed thread count is not required: results are compared with tolerances, except the
integer / schedule-buffer fixtures which are bit-exact.
"""
import json
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle import ref_import as R  # noqa: E402
from oracle.video_oracle import seeded_state_dict  # noqa: E402
from tests.golden.configs import TINY_UNET, tiny_inputs, config1_inputs  # noqa: E402


def shapes_of(module):
    return {k: tuple(v.shape) for k, v in module.state_dict().items()}


def build_ref_diffusion(unet_kwargs, channels, image_size, timesteps, sampling_timesteps, seed):
    UL, U, G = R.Unet_Libero(), R.UNetModel(), R.GoalGaussianDiffusion()
    net = UL.__new__(UL)
    torch.nn.Module.__init__(net)
    net.unet = U(**unet_kwargs)
    net.load_state_dict(seeded_state_dict(shapes_of(net), seed))
    diff = G(net, image_size=image_size, channels=channels, timesteps=timesteps,
             sampling_timesteps=sampling_timesteps, loss_type="l2", objective="pred_v",
             beta_schedule="cosine", min_snr_loss_weight=True, guidance_weight=0)
    return diff.eval()


def main():
    torch.set_grad_enabled(False)
    out = {}

    # ---- tiny UNet: every block kind, 2 samples, 3 frames, 16x16 ----
    diff = build_ref_diffusion(TINY_UNET, channels=9, image_size=(16, 16), timesteps=4, sampling_timesteps=4, seed=1)
    with open(os.path.join(HERE, "tiny_unet_state_dict_layout.json"), "w") as f:
        json.dump({k: list(v) for k, v in shapes_of(diff.model).items()}, f, indent=0)
    x, t, x_cond, te = tiny_inputs()
    out["tiny_forward"] = diff.model(torch.cat([x, x_cond], 1), t, te)
    torch.manual_seed(77)
    out["tiny_ddpm4"] = diff.sample(x_cond, te, batch_size=2)
    diff10 = build_ref_diffusion(TINY_UNET, channels=9, image_size=(16, 16), timesteps=10, sampling_timesteps=3, seed=1)
    assert diff10.is_ddim_sampling
    torch.manual_seed(78)
    out["tiny_ddim3of10"] = diff10.sample(x_cond, te, batch_size=2)

    # ---- config 1: the real Unet_Libero, 64x64, 4 frames, batch 1 ----
    UL, G = R.Unet_Libero(), R.GoalGaussianDiffusion()
    net = UL()
    full_shapes = shapes_of(net)
    net.load_state_dict(seeded_state_dict(full_shapes, 2))
    x, t, x_cond, te = config1_inputs()
    out["config1_forward"] = net.eval()(torch.cat([x, x_cond], 1), t, te)
    d1 = G(net, image_size=(64, 64), channels=12, timesteps=100, sampling_timesteps=1, loss_type="l2",
           objective="pred_v", beta_schedule="cosine", min_snr_loss_weight=True, guidance_weight=0).eval()
    torch.manual_seed(123)
    out["config1_ddim1"] = d1.sample(x_cond, te, batch_size=1)
    d2 = G(net, image_size=(64, 64), channels=12, timesteps=2, sampling_timesteps=2, loss_type="l2",
           objective="pred_v", beta_schedule="cosine", min_snr_loss_weight=True, guidance_weight=0).eval()
    torch.manual_seed(124)
    out["config1_ddpm2"] = d2.sample(x_cond, te, batch_size=1)

    # ---- schedule buffers (bit-exact) + state-dict layout of the shipped config ----
    d100 = G(net, image_size=(128, 128), channels=21, timesteps=100, sampling_timesteps=100, loss_type="l2",
             objective="pred_v", beta_schedule="cosine", min_snr_loss_weight=True, guidance_weight=0)
    for k, v in d100.state_dict().items():
        if not k.startswith("model."):
            out["buf100." + k] = v.clone()
    layout = {k: list(v.shape) for k, v in d100.state_dict().items()}
    with open(os.path.join(HERE, "goal_diffusion_state_dict_layout.json"), "w") as f:
        json.dump(layout, f, indent=0)
    torch.save({k: v.contiguous() for k, v in out.items()}, os.path.join(HERE, "video_golden.pt"))
    for k, v in out.items():
        print(f"{k:28s} {tuple(v.shape)} mean {v.float().mean():+.6f} std {v.float().std():.6f}")
    print("state_dict entries:", len(layout), "params:", sum(torch.tensor(s).prod().item() if s else 1 for s in layout.values()))


if __name__ == "__main__":
    main()
