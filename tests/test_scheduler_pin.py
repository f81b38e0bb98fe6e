"""The policy's noise schedulers (restatement of diffusers DDPMScheduler / DDIMScheduler for the yaml's settings,
v2a_b200/diffusion_policy.py) against the reference's VENDORED twin of the same algorithms
(guided_diffusion/gaussian_diffusion.py + respace.py, run unmodified in float64 by
tests/golden/make_scheduler_golden.py).  diffusers itself is neither installed nor vendored, so this is the
strongest pin available offline: schedule, add_noise, the clipped ancestral step and the eta = 0 DDIM trajectory."""
import os

import pytest
import torch

from v2a_b200 import diffusion_policy as DP

GOLD = os.path.join(os.path.dirname(__file__), "golden", "scheduler_golden.pt")
SCHED = dict(num_train_timesteps=100, beta_start=0.0001, beta_end=0.02, beta_schedule="squaredcos_cap_v2",
             clip_sample=True, prediction_type="epsilon")


def _toy_eps_model(x, t):
    return 0.25 * x + 0.1 * torch.sin(0.37 * t.to(x.dtype))[:, None, None]


def _rel(a, b):
    return ((a.double() - b.double()).norm() / b.double().norm()).item()


def test_schedule_and_add_noise():
    g = torch.load(GOLD)
    s = DP.DDPMScheduler(variance_type="fixed_small", **SCHED)
    assert torch.equal(s.betas, g["betas"].float())                     # fp32 rounding of the same float64 values
    assert _rel(s.alphas_cumprod, g["alphas_cumprod"]) < 2e-7
    got = s.add_noise(g["x0"], g["noise"], g["t"])
    assert _rel(got, g["q_sample"]) < 1e-6
    # the oracle's twin of the same table (oracle/policy_oracle.py) agrees too
    from oracle import policy_oracle as PO
    assert torch.equal(PO.ddpm_alphas_cumprod(100), s.alphas_cumprod)


def test_ddpm_ancestral_step_with_clipping():
    g = torch.load(GOLD)
    s = DP.DDPMScheduler(variance_type="fixed_small", **SCHED)
    x = g["ddpm_x"]
    clipped = 0
    for st in g["ddpm_steps"]:
        t = st["t"]
        eps = _toy_eps_model(x, torch.full((4,), t))
        torch.manual_seed(st["seed"])                                   # the twin draws randn_like(x) inside p_sample
        r = s.step(eps, t, x)
        # diffusers (and this restatement) hold the schedule in fp32, the twin in float64: at t = 99, where
        # alphas_cumprod = 2.4e-7 is divided by, the fp32 cumprod's 1e-6 shows up as ~5e-6; a wrong coefficient
        # would be off by 1e-2 or more
        assert _rel(r.pred_original_sample, st["pred_xstart"]) < 2e-5
        assert _rel(r.prev_sample, st["sample"]) < 2e-5, t
        clipped += int((st["pred_xstart"].abs() == 1.0).sum())
    assert clipped > 0                                                  # the clip_sample branch was exercised


def test_ddim_trajectory_eta0():
    g = torch.load(GOLD)
    s = DP.DDIMScheduler(set_alpha_to_one=True, steps_offset=0, **SCHED)
    s.set_timesteps(8)
    assert s.timesteps.tolist() == g["ddim_timesteps"].tolist() == [84, 72, 60, 48, 36, 24, 12, 0]
    x = g["ddim_traj"][0]
    for i, t in enumerate(s.timesteps):
        x = s.step(_toy_eps_model(x, torch.full((4,), int(t))), t, x).prev_sample
        assert _rel(x, g["ddim_traj"][i + 1]) < 1e-6, int(t)
