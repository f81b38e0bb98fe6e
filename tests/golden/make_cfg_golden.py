"""Golden vectors for classifier-free guidance (guidance_weight > 0) from the UNMODIFIED
reference GoalGaussianDiffusion (run in the build container only: python tests/golden/make_cfg_golden.py)."""
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from tests.golden.configs import TINY_UNET, tiny_inputs  # noqa: E402
from tests.golden.make_golden import build_ref_diffusion  # noqa: E402

GW = 1.5


def main():
    torch.set_grad_enabled(False)
    _, _, x_cond, te = tiny_inputs()
    out = {}
    d = build_ref_diffusion(TINY_UNET, channels=9, image_size=(16, 16), timesteps=4, sampling_timesteps=4, seed=1)
    d.guidance_weight = GW                       # poked from outside, as diffuser/models/train_utils.py:23-30 does
    torch.manual_seed(81)
    out["tiny_ddpm4_cfg"] = d.sample(x_cond, te, batch_size=2)
    d10 = build_ref_diffusion(TINY_UNET, channels=9, image_size=(16, 16), timesteps=10, sampling_timesteps=3, seed=1)
    d10.guidance_weight = GW
    torch.manual_seed(82)
    out["tiny_ddim3of10_cfg"] = d10.sample(x_cond, te, batch_size=2)
    torch.save({k: v.contiguous() for k, v in out.items()}, os.path.join(HERE, "video_cfg_golden.pt"))
    for k, v in out.items():
        print(f"{k:24s} {tuple(v.shape)} mean {v.mean():+.6f} std {v.std():.6f}")


if __name__ == "__main__":
    main()
