"""Pin oracle/policy_oracle.py against golden vectors from the unmodified reference
ConditionalUnet1D (forward, loss, input gradients, per-parameter gradient fingerprints)
and, when the reference checkout is mounted, against the live module.  CPU only."""
import json
import os

import pytest
import torch

from oracle import policy_oracle as PO
from oracle import ref_import
from tests.golden.configs import POLICY_LIBERO, POLICY_TINY, grad_fingerprint, policy_inputs

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = torch.load(os.path.join(HERE, "golden", "policy_golden.pt"))
with open(os.path.join(HERE, "golden", "policy_golden_meta.json")) as f:
    META = json.load(f)


def _oracle_run(name, cfg):
    m = META[name]
    sd = PO.seeded_policy_state_dict({k: tuple(v) for k, v in m["layout"].items()}, m["seed"])
    sd = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    traj, noise, t, gc = policy_inputs(m["B"], cfg, m["seed"])
    gc = gc.clone().requires_grad_(True)
    acp = PO.ddpm_alphas_cumprod(100)
    noisy = PO.add_noise(acp, traj, noise, t).requires_grad_(True)
    pred = PO.unet1d_forward(sd, noisy, t, gc, n_groups=cfg["n_groups"])
    loss = torch.nn.functional.mse_loss(pred, noise, reduction="none").reshape(m["B"], -1).mean(1).mean()
    loss.backward()
    return sd, pred, loss, gc, noisy


@pytest.mark.parametrize("name,cfg", [("tiny", POLICY_TINY), ("libero", POLICY_LIBERO)])
def test_policy_oracle_matches_reference_golden(name, cfg):
    sd, pred, loss, gc, noisy = _oracle_run(name, cfg)
    close = lambda a, b: torch.testing.assert_close(a, b, rtol=1e-4, atol=1e-6)
    close(pred.detach(), GOLD[f"{name}.pred"])
    close(loss.detach(), GOLD[f"{name}.loss"])
    close(gc.grad, GOLD[f"{name}.d_global_cond"])
    close(noisy.grad, GOLD[f"{name}.d_sample"])
    for k, (norm, proj) in META[name]["grad_fingerprints"].items():
        n2, p2 = grad_fingerprint(k, sd[k].grad)
        assert abs(n2 - norm) <= 1e-4 * max(norm, 1e-8), k
        assert abs(p2 - proj) <= 1e-4 * max(norm, 1e-8) * (sd[k].numel() ** 0.5), k


def test_ddpm_schedule_restatement():
    acp = PO.ddpm_alphas_cumprod(100)
    assert acp.shape == (100,) and acp[0] > 0.99 and acp[-1] < 1e-3 and (acp[1:] < acp[:-1]).all()


@pytest.mark.skipif(not ref_import.available(), reason="reference checkout not mounted")
def test_policy_oracle_matches_live_reference():
    net = ref_import.ConditionalUnet1D()(**POLICY_TINY)
    shapes = {k: tuple(v.shape) for k, v in net.state_dict().items()}
    sd = PO.seeded_policy_state_dict(shapes, 3)
    net.load_state_dict(sd)
    traj, noise, t, gc = policy_inputs(2, POLICY_TINY, 5)
    out_ref = net(traj, t, global_cond=gc)
    out = PO.unet1d_forward(sd, traj, t, gc)
    torch.testing.assert_close(out, out_ref, rtol=1e-5, atol=1e-6)
    # int / 0-d timestep forms accepted by the reference forward
    torch.testing.assert_close(PO.unet1d_forward(sd, traj, torch.tensor(7), gc), net(traj, 7, global_cond=gc),
                               rtol=1e-5, atol=1e-6)
