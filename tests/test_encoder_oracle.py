"""Pins oracle/encoder_oracle.py (the CPU restatement of the observation encoder):
against the golden vectors generated from the UNMODIFIED reference VisualCore (tests/golden/make_encoder_golden.py)
and, when /root/reference is mounted, against the live reference classes."""
import json
import os

import pytest
import torch

from oracle import encoder_oracle as EO
from oracle import policy_oracle as PO
from oracle import ref_import as R
from tests.golden.configs import encoder_inputs, grad_fingerprint

HERE = os.path.dirname(os.path.abspath(__file__))


def _meta():
    with open(os.path.join(HERE, "golden", "encoder_golden_meta.json")) as f:
        meta = json.load(f)
    with open(os.path.join(HERE, "golden", "policy_loss_golden_meta.json")) as f:
        layout = json.load(f)["layout"]
    return meta, layout


def _oracle_sd(layout, seed):
    """seeded weights + the SpatialSoftmax buffers the reference constructs (pos grid, temperature)."""
    from v2a_b200 import diffusion_policy as DP
    sd = DP.build_libero_policy().state_dict()
    sd.update(PO.seeded_full_policy_state_dict(layout, seed))
    return sd


def test_encoder_oracle_matches_reference_golden():
    meta, layout = _meta()
    sd = {k: v.clone().requires_grad_(v.dtype.is_floating_point and v.numel() > 0 and "pos_" not in k and "temperature" not in k)
          for k, v in _oracle_sd(layout, meta["seed"]).items()}
    x, w = encoder_inputs(meta["B"], meta["seed"])
    feat = EO.visual_core_forward(sd, meta["key"], x)
    gold = torch.load(os.path.join(HERE, "golden", "encoder_golden.pt"))
    torch.testing.assert_close(feat, gold["feat"], rtol=1e-5, atol=1e-6)
    (feat * w).sum().backward()
    assert len(meta["grad_fingerprints"]) == 64
    for n, (norm, proj) in meta["grad_fingerprints"].items():
        # the Sequential alias 'nets.0.*' / 'nets.1.*' of the reference maps onto backbone.* / pool.*
        k = meta["key"] + n
        g = sd[k].grad
        assert g is not None, k
        n2, p2 = grad_fingerprint(k, g)
        assert abs(n2 - norm) <= 1e-4 * max(norm, 1e-8), (n, n2, norm)
        assert abs(p2 - proj) <= 1e-4 * max(norm, 1e-8) * g.numel() ** 0.5, (n, p2, proj)


@pytest.mark.skipif(not R.available(), reason="reference checkout not mounted")
def test_encoder_oracle_matches_live_reference_encoder():
    from tests.golden.make_policy_loss_golden import build_reference_policy
    meta, layout = _meta()
    ref = build_reference_policy()
    sd = ref.state_dict()
    sd.update(PO.seeded_full_policy_state_dict(layout, 5))
    ref.load_state_dict(sd, strict=True)
    ref.eval()
    g = torch.Generator().manual_seed(9)
    obs = {"img_goal_1": torch.rand(2, 3, 128, 128, generator=g) * 2 - 1,
           "img_obs_1": torch.rand(2, 3, 128, 128, generator=g) * 2 - 1}
    with torch.no_grad():
        want = ref.obs_encoder(obs)
        got = EO.obs_encoder_forward(ref.state_dict(), "obs_encoder.", obs)
    torch.testing.assert_close(got, want, rtol=1e-5, atol=1e-6)
