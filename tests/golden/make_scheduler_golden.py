"""Pin for the policy's noise-scheduler restatement (build container only: python tests/golden/make_scheduler_golden.py).

`diffusers` (the library the reference imports, unpinned in its requirements.txt) is not installed and not vendored,
but the reference DOES vendor an independent implementation of the same published algorithms (improved-DDPM):
flowdiffusion/flowdiffusion/guided_diffusion/guided_diffusion/gaussian_diffusion.py (+ respace.py).  This script runs
that UNMODIFIED in-repo twin in float64 on seeded inputs and stores:
  * the cosine ("squaredcos_cap_v2") betas / alphas_cumprod for T = 100        (gaussian_diffusion.py:18-62)
  * q_sample = add_noise                                                        (:188-204)
  * one ancestral p_sample step per t with FIXED_SMALL variance and x0 clipping (:256-342, :404-440)
  * an 8-step eta = 0 DDIM trajectory over the respaced steps [84 .. 0]         (:537-585, respace.py:63-122),
    with a model whose x0 stays inside [-1, 1] (the twin re-derives eps from the clipped x0, diffusers keeps the
    model's eps -- the two only agree while the clip is inactive; with the clip active diffusers' published rule is
    what diffusion_policy.DDIMScheduler follows, and that branch stays pinned by its own formula test only).
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import ref_import as R  # noqa: E402

GD = "flowdiffusion.flowdiffusion.guided_diffusion.guided_diffusion."


def toy_eps_model(x, t, **kw):
    """Deterministic stand-in for the network: eps prediction as a smooth function of (x, t)."""
    return 0.25 * x + 0.1 * torch.sin(0.37 * t.to(x.dtype))[:, None, None]


def main():
    gd = R._imp(GD + "gaussian_diffusion")
    rs = R._imp(GD + "respace")
    T = 100
    betas = gd.get_named_beta_schedule("cosine", T)
    kw = dict(model_mean_type=gd.ModelMeanType.EPSILON, model_var_type=gd.ModelVarType.FIXED_SMALL,
              loss_type=gd.LossType.MSE)
    d = gd.GaussianDiffusion(betas=betas, **kw)
    g = torch.Generator().manual_seed(5)
    x0 = (torch.rand(4, 16, 7, generator=g, dtype=torch.float64) * 2 - 1)
    noise = torch.randn(4, 16, 7, generator=g, dtype=torch.float64)
    tt = torch.tensor([0, 17, 50, 99])
    out = {"betas": torch.from_numpy(betas), "alphas_cumprod": torch.from_numpy(d.alphas_cumprod),
           "x0": x0, "noise": noise, "t": tt, "q_sample": d.q_sample(x0, tt, noise=noise)}
    # ancestral steps: the twin draws th.randn_like(x) inside p_sample -> replay the same draw in the test
    xs = torch.randn(4, 16, 7, generator=g, dtype=torch.float64) * 1.5       # large enough that the x0 clip bites
    steps = []
    for t in (99, 50, 1, 0):
        torch.manual_seed(1000 + t)
        r = d.p_sample(toy_eps_model, xs, torch.full((4,), t), clip_denoised=True)
        steps.append({"t": t, "seed": 1000 + t, "sample": r["sample"], "pred_xstart": r["pred_xstart"]})
    out["ddpm_x"] = xs
    out["ddpm_steps"] = steps
    # DDIM, 8 of 100 steps ("leading" spacing 0, 12, ..., 84), eta = 0
    use = list(range(0, 96, 12))
    sp = rs.SpacedDiffusion(use_timesteps=set(use), betas=betas, **kw)
    x = torch.randn(4, 16, 7, generator=g, dtype=torch.float64) * 0.05       # keeps |x0| < 1: clip inactive
    traj = [x]
    for i in reversed(range(len(use))):
        r = sp.ddim_sample(toy_eps_model, x, torch.full((4,), i), clip_denoised=True, eta=0.0)
        assert float(r["pred_xstart"].abs().max()) < 1.0
        x = r["sample"]
        traj.append(x)
    out["ddim_timesteps"] = torch.tensor(list(reversed(use)))
    out["ddim_traj"] = torch.stack(traj)
    torch.save(out, os.path.join(HERE, "scheduler_golden.pt"))
    print("saved", {k: (tuple(v.shape) if torch.is_tensor(v) else type(v).__name__) for k, v in out.items()})


if __name__ == "__main__":
    main()
