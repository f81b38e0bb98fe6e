"""Golden vectors for (a) trajectory drift over the FULL 100-step DDPM loop the bench times and (b)
`return_all_timesteps`, from the UNMODIFIED reference GoalGaussianDiffusion (build container only:
python tests/golden/make_drift_golden.py).  Tiny UNet (every block kind), seeded weights and inputs, CPU RNG stream.

  tiny_ddpm100      sample() with timesteps = sampling_timesteps = 100 (goal_diffusion.py:582-599), [2, 9, 16, 16]
  tiny_ddpm4_all    sample(return_all_timesteps=True), 4 DDPM steps: [2, 5, 9, 16, 16]
  tiny_ddim3_all    same for the 3-of-10 DDIM plan (goal_diffusion.py:601-641): [2, 4, 9, 16, 16]
"""
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from tests.golden.configs import TINY_UNET, tiny_inputs  # noqa: E402
from tests.golden.make_golden import build_ref_diffusion  # noqa: E402


def main():
    torch.set_grad_enabled(False)
    _, _, x_cond, te = tiny_inputs()
    out = {}
    d100 = build_ref_diffusion(TINY_UNET, channels=9, image_size=(16, 16), timesteps=100, sampling_timesteps=100, seed=1)
    assert not d100.is_ddim_sampling
    torch.manual_seed(91)
    out["tiny_ddpm100"] = d100.sample(x_cond, te, batch_size=2)
    d4 = build_ref_diffusion(TINY_UNET, channels=9, image_size=(16, 16), timesteps=4, sampling_timesteps=4, seed=1)
    torch.manual_seed(92)
    out["tiny_ddpm4_all"] = d4.sample(x_cond, te, batch_size=2, return_all_timesteps=True)
    d10 = build_ref_diffusion(TINY_UNET, channels=9, image_size=(16, 16), timesteps=10, sampling_timesteps=3, seed=1)
    torch.manual_seed(93)
    out["tiny_ddim3_all"] = d10.sample(x_cond, te, batch_size=2, return_all_timesteps=True)
    torch.save({k: v.contiguous() for k, v in out.items()}, os.path.join(HERE, "video_drift_golden.pt"))
    for k, v in out.items():
        print(f"{k:18s} {tuple(v.shape)} mean {v.mean():+.6f} std {v.std():.6f}")


if __name__ == "__main__":
    main()
