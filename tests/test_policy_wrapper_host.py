"""CPU tests of the DiffusionUnetImagePolicy drop-in: state_dict layout of the whole policy equals the
reference's (414 keys), scheduler / normaliser restatements, the observation encoder against the live
reference modules (when /root/reference is mounted), install() rebinding."""
import copy
import importlib
import sys
import json
import os

import numpy as np
import pytest
import torch

from tests.stock_twins import global_cond_stock

from oracle import policy_oracle as PO
from oracle import ref_import as R
from tests.golden.configs import policy_loss_batch
from v2a_b200 import diffusion_policy as DP

HERE = os.path.dirname(os.path.abspath(__file__))
with open(os.path.join(HERE, "golden", "policy_loss_golden_meta.json")) as f:
    META = json.load(f)


@pytest.fixture(scope="module")
def policy():
    return DP.build_libero_policy()


def test_full_policy_state_dict_layout_equals_reference(policy):
    lay = {k: list(v.shape) for k, v in policy.state_dict().items()}
    assert len(lay) == 414
    assert list(lay.keys()) == list(META["layout"].keys())
    assert lay == META["layout"]
    assert sum(p.numel() for p in policy.parameters()) == 87_219_143
    # zero-size parameters of ModuleAttrMixin come first (optimiser / EMA must tolerate them)
    assert list(lay)[:2] == ["_dummy_variable", "obs_encoder._dummy_variable"]
    clone = copy.deepcopy(policy)                                        # ema_pytorch.EMA deep-copies
    clone.load_state_dict(policy.state_dict(), strict=True)
    assert policy.obs_encoder.rgb_keys == ["img_goal_1", "img_obs_1"]   # sorted: goal features first


def test_schedulers_restate_diffusers_for_the_yaml_settings():
    d = DP.DDPMScheduler(num_train_timesteps=100, prediction_type="epsilon")
    assert torch.equal(d.alphas_cumprod, PO.ddpm_alphas_cumprod(100))
    g = torch.Generator().manual_seed(0)
    x, n = torch.randn(5, 16, 7, generator=g), torch.randn(5, 16, 7, generator=g)
    t = torch.tensor([0, 3, 50, 98, 99])
    assert torch.equal(d.add_noise(x, n, t), PO.add_noise(d.alphas_cumprod, x, n, t))
    assert d.config.num_train_timesteps == 100 and d.config.prediction_type == "epsilon"
    i = DP.DDIMScheduler(num_train_timesteps=100, set_alpha_to_one=True, steps_offset=0)
    i.set_timesteps(8)
    assert i.timesteps.tolist() == [84, 72, 60, 48, 36, 24, 12, 0]
    # one DDIM step in closed form (eta = 0, clip_sample): x0 clipped, direction keeps the raw epsilon
    eps, xt = torch.randn(2, 16, 7, generator=g), 3 * torch.randn(2, 16, 7, generator=g)
    a_t, a_p = d.alphas_cumprod[84].double(), d.alphas_cumprod[72].double()
    x0 = ((xt.double() - (1 - a_t).sqrt() * eps.double()) / a_t.sqrt()).clamp(-1, 1)
    want = a_p.sqrt() * x0 + (1 - a_p).sqrt() * eps.double()
    got = i.step(eps, torch.tensor(84), xt).prev_sample
    assert (got.double() - want).abs().max() < 1e-5
    # last step lands on alpha = 1: prev_sample == clipped x0
    a0 = d.alphas_cumprod[0].double()
    x0 = ((xt.double() - (1 - a0).sqrt() * eps.double()) / a0.sqrt()).clamp(-1, 1)
    assert (i.step(eps, torch.tensor(0), xt).prev_sample.double() - x0).abs().max() < 1e-5
    # DDPM ancestral step: posterior mean in closed form at t = 0 (no noise is drawn there)
    d.set_timesteps(100)
    st = torch.get_rng_state()
    out = d.step(eps, torch.tensor(0), xt).prev_sample
    assert torch.equal(st, torch.get_rng_state())
    assert (out.double() - x0).abs().max() < 1e-4
    with pytest.raises(NotImplementedError):
        DP.DDPMScheduler(beta_schedule="linear")


def test_normaliser_is_2x_minus_1_for_images_and_identity_for_libero_actions(policy):
    b = policy_loss_batch(2, 3)
    n = policy.normalizer.normalize_d(b["obs"])
    assert torch.allclose(n["img_obs_1"], 2 * b["obs"]["img_obs_1"] - 1)
    a = policy.normalizer["action"].normalize(b["action"])
    assert torch.allclose(a, b["action"], atol=1e-7)
    assert torch.allclose(policy.normalizer["action"].unnormalize(a * 3), b["action"].mul(3).clamp(-1, 1), atol=1e-6)


def test_compute_loss_fails_loudly_without_cuda(policy):
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    with pytest.raises(RuntimeError, match="CUDA"):
        policy.compute_loss(policy_loss_batch(1, 0))


@pytest.mark.skipif(not R.available(), reason="reference checkout not mounted")
def test_encoder_and_install_against_live_reference(policy, monkeypatch):
    from tests.golden.make_policy_loss_golden import build_reference_policy
    from v2a_b200 import install
    ref = build_reference_policy()
    assert list(ref.state_dict().keys()) == list(policy.state_dict().keys())
    mine = copy.deepcopy(policy)
    mine.load_state_dict(ref.state_dict(), strict=True)
    b = policy_loss_batch(2, 5)
    torch.manual_seed(1)
    got, B = global_cond_stock(mine, b["obs"])     # the parameter-holding torch modules, evaluated on the host
    nobs = ref.normalizer.normalize_d(b["obs"])
    torch.manual_seed(1)
    want = ref.obs_encoder({k: v[:, :1].reshape(-1, *v.shape[2:]) for k, v in nobs.items()}).reshape(2, -1)
    assert torch.equal(got, want) and B == 2                              # same torch ops: bitwise
    try:
        done = install.install()
        assert "diffuser.diffusion_policy.diffusion_unet_image_policy.ConditionalUnet1D" in done
        swapped = build_reference_policy()                                 # the REFERENCE class, our UNet1D inside
        assert type(swapped.model).__module__.startswith("v2a_b200")
        assert list(swapped.state_dict().keys()) == list(ref.state_dict().keys())
        gd = importlib.import_module("flowdiffusion.flowdiffusion.goal_diffusion")
        assert gd.GoalGaussianDiffusion.__module__.startswith("v2a_b200")
        if "diffuser.datasets.env_img_replay_buffer" in sys.modules:        # row N4 (needs the simulator stub)
            rb = sys.modules["diffuser.datasets.env_img_replay_buffer"]
            assert rb.Global_EnvReplayBuffer_Img.__module__.startswith("v2a_b200")
    finally:
        install.uninstall()
    pol = importlib.import_module("diffuser.diffusion_policy.diffusion_unet_image_policy")
    assert not pol.ConditionalUnet1D.__module__.startswith("v2a_b200")
