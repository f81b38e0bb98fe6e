"""Pin the CPU oracle (oracle/video_oracle.py) against the golden vectors produced by the
UNMODIFIED reference (tests/golden/make_golden.py) and, when the reference checkout is
present, against the live reference modules.  CPU only."""
import json
import os

import pytest
import torch

from oracle import ref_import
from oracle import video_oracle as VO
from tests.golden.configs import TINY_UNET, config1_inputs, tiny_inputs

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = torch.load(os.path.join(HERE, "golden", "video_golden.pt"))
with open(os.path.join(HERE, "golden", "goal_diffusion_state_dict_layout.json")) as f:
    LAYOUT = json.load(f)


def rel_l2(a, b):
    return ((a.double() - b.double()).norm() / b.double().norm()).item()


def tiny_shapes():
    """Shapes of the tiny UNet = the matching subset logic of the full layout is not
    available, so they are stored with the fixture file."""
    with open(os.path.join(HERE, "golden", "tiny_unet_state_dict_layout.json")) as f:
        return {k: tuple(v) for k, v in json.load(f).items()}


def test_schedule_buffers_bit_exact():
    buf = VO.cosine_schedule_buffers(100)
    for k, v in buf.items():
        assert torch.equal(v, GOLD["buf100." + k]), k


def test_tiny_forward_and_samplers_match_reference_golden():
    sd = VO.seeded_state_dict(tiny_shapes(), 1)
    x, t, x_cond, te = tiny_inputs()
    with torch.no_grad():
        out = VO.unet_libero_forward(sd, torch.cat([x, x_cond], 1), t, te)
        assert rel_l2(out, GOLD["tiny_forward"]) < 1e-5
        torch.manual_seed(77)
        s = VO.ddpm_sample(sd, VO.cosine_schedule_buffers(4), x_cond, te, (2, 9, 16, 16))
        assert rel_l2(s, GOLD["tiny_ddpm4"]) < 1e-5
        torch.manual_seed(78)
        s = VO.ddim_sample(sd, VO.cosine_schedule_buffers(10), x_cond, te, (2, 9, 16, 16), 3)
        assert rel_l2(s, GOLD["tiny_ddim3of10"]) < 1e-5


def test_oracle_100_step_trajectory_matches_reference_golden():
    """The whole ancestral loop the bench times (100 clamped posterior steps, goal_diffusion.py:582-599):
    pins the oracle's sampler over the full trajectory, not only 2-4 steps (tests/golden/make_drift_golden.py)."""
    gold = torch.load(os.path.join(HERE, "golden", "video_drift_golden.pt"))
    sd = VO.seeded_state_dict(tiny_shapes(), 1)
    _, _, x_cond, te = tiny_inputs()
    with torch.no_grad():
        torch.manual_seed(91)
        s = VO.ddpm_sample(sd, VO.cosine_schedule_buffers(100), x_cond, te, (2, 9, 16, 16))
    assert rel_l2(s, gold["tiny_ddpm100"]) < 1e-4
    assert tuple(gold["tiny_ddpm4_all"].shape) == (2, 5, 9, 16, 16) and tuple(gold["tiny_ddim3_all"].shape) == (2, 4, 9, 16, 16)


def test_classifier_free_guidance_samplers_match_reference_golden():
    """guidance_weight > 0: doubled batch, zeroed task tokens, noise-space mixing."""
    gold = torch.load(os.path.join(HERE, "golden", "video_cfg_golden.pt"))
    sd = VO.seeded_state_dict(tiny_shapes(), 1)
    _, _, x_cond, te = tiny_inputs()
    with torch.no_grad():
        torch.manual_seed(81)
        s = VO.ddpm_sample(sd, VO.cosine_schedule_buffers(4), x_cond, te, (2, 9, 16, 16), guidance_weight=1.5)
        assert rel_l2(s, gold["tiny_ddpm4_cfg"]) < 1e-5
        torch.manual_seed(82)
        s = VO.ddim_sample(sd, VO.cosine_schedule_buffers(10), x_cond, te, (2, 9, 16, 16), 3, guidance_weight=1.5)
        assert rel_l2(s, gold["tiny_ddim3of10_cfg"]) < 1e-5
        # guidance really changes the result (the fixture is not vacuous)
        assert rel_l2(s, GOLD["tiny_ddim3of10"]) > 1e-2


def test_config1_forward_matches_reference_golden():
    shapes = {k[len("model."):]: tuple(v) for k, v in LAYOUT.items() if k.startswith("model.")}
    sd = VO.seeded_state_dict(shapes, 2)
    x, t, x_cond, te = config1_inputs()
    with torch.no_grad():
        out = VO.unet_libero_forward(sd, torch.cat([x, x_cond], 1), t, te)
    assert rel_l2(out, GOLD["config1_forward"]) < 1e-5


@pytest.mark.skipif(not ref_import.available(), reason="reference checkout not mounted")
def test_oracle_matches_live_reference_modules():
    U, UL = ref_import.UNetModel(), ref_import.Unet_Libero()
    net = UL.__new__(UL)
    torch.nn.Module.__init__(net)
    net.unet = U(**TINY_UNET)
    shapes = {k: tuple(v.shape) for k, v in net.state_dict().items()}
    assert shapes == tiny_shapes()
    sd = VO.seeded_state_dict(shapes, 5)
    net.load_state_dict(sd)
    g = torch.Generator().manual_seed(9)
    x = torch.randn(1, 15, 16, 16, generator=g)  # 4 frames + cond
    t = torch.tensor([2])
    te = torch.randn(1, 5, 512, generator=g)
    with torch.no_grad():
        assert rel_l2(VO.unet_libero_forward(sd, x, t, te), net.eval()(x, t, te)) < 1e-5
