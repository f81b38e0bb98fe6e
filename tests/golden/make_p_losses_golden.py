"""Golden values for the denoising loss (`GoalGaussianDiffusion.p_losses` / `.forward`, goal_diffusion.py:689-724) from
the UNMODIFIED reference (build container only: python tests/golden/make_p_losses_golden.py).  Tiny UNet, seeded weights
and inputs, explicit timesteps and noise so no RNG stream is involved.

  inputs   img in [0, 1] [2, 9, 16, 16] (normalised to [-1, 1] by the caller as forward() does), noise, t = [37, 4]
  losses   reference p_losses(normalize(img), t, x_cond, task_embed, noise) for timesteps = 100, objective pred_v,
           l2, min-SNR weighting (the shipped configuration), and the same with the loss weight gathered at t = [0, 99]
"""
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from tests.golden.configs import TINY_UNET, p_losses_inputs, tiny_inputs  # noqa: E402
from tests.golden.make_golden import build_ref_diffusion  # noqa: E402


def main():
    torch.set_grad_enabled(False)
    _, _, x_cond, te = tiny_inputs()
    img, noise = p_losses_inputs()
    d = build_ref_diffusion(TINY_UNET, channels=9, image_size=(16, 16), timesteps=100, sampling_timesteps=100, seed=1)
    out = {}
    for name, t in (("t_37_4", [37, 4]), ("t_0_99", [0, 99])):
        tt = torch.tensor(t, dtype=torch.long)
        out[name] = d.p_losses(d.normalize(img), tt, x_cond, te, noise=noise).reshape(1).clone()
        print(name, float(out[name]))
    torch.save(out, os.path.join(HERE, "video_p_losses_golden.pt"))


if __name__ == "__main__":
    main()
