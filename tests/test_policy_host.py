"""CPU tests of the policy host module: reference-identical state_dict layout (names, shapes,
ORDER), loud failure without CUDA, deepcopy (EMA) mechanics."""
import copy
import json
import os

import pytest
import torch

from tests.golden.configs import POLICY_LIBERO, POLICY_TINY
from v2a_b200.policy_unet1d import ConditionalUnet1D

HERE = os.path.dirname(os.path.abspath(__file__))
with open(os.path.join(HERE, "golden", "policy_golden_meta.json")) as f:
    META = json.load(f)


@pytest.mark.parametrize("name,cfg", [("tiny", POLICY_TINY), ("libero", POLICY_LIBERO)])
def test_policy_state_dict_layout_equals_reference(name, cfg):
    net = ConditionalUnet1D(**cfg)
    lay = {k: list(v.shape) for k, v in net.state_dict().items()}
    assert list(lay.keys()) == list(META[name]["layout"].keys())
    assert lay == META[name]["layout"]
    copy.deepcopy(net).load_state_dict(net.state_dict(), strict=True)


def test_policy_forward_fails_loudly_without_cuda():
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    net = ConditionalUnet1D(**POLICY_TINY)
    with pytest.raises(RuntimeError, match="CUDA"):
        net(torch.zeros(2, 16, 7), torch.tensor([1, 2]), global_cond=torch.zeros(2, 32))


def test_unsupported_configurations_are_rejected():
    with pytest.raises(NotImplementedError):
        ConditionalUnet1D(input_dim=7, local_cond_dim=4, global_cond_dim=32, cond_predict_scale=True)


def test_train_step_checkpoint_formats_roundtrip_with_stock_adamw():
    """PolicyTrainStep's optimiser / EMA state in the reference trainer's checkpoint formats
    (lb_online_trainer_v7.py:367-407): slab state -> AdamW.state_dict() -> stock AdamW, and back."""
    import copy
    from v2a_b200 import train_step as TS
    torch.manual_seed(0)
    net = torch.nn.Sequential(torch.nn.Linear(5, 4), torch.nn.GroupNorm(2, 4), torch.nn.Linear(4, 3))
    net.register_parameter("_dummy_variable", torch.nn.Parameter(torch.empty(0)))   # ModuleAttrMixin's empty param
    net.register_buffer("table", torch.arange(3.0))
    ref = copy.deepcopy(net)
    opt = torch.optim.AdamW(ref.parameters(), lr=1e-4, betas=(0.95, 0.999), eps=1e-8, weight_decay=1e-6)
    for _ in range(3):                                             # a real optimiser history to import
        opt.zero_grad()
        ref(torch.randn(7, 5)).square().mean().backward()
        opt.step()
    params = list(net.parameters())
    segs = [TS._Segment(params[:3], "cpu", own_grad=True, ema=True), TS._Segment(params[3:], "cpu", own_grad=True, ema=True)]
    steps = TS.import_optimizer_state(net, segs, opt.state_dict())
    assert steps == 3
    exported = TS.export_optimizer_state(net, segs, steps, lr=1e-4, betas=(0.95, 0.999), eps=1e-8, weight_decay=1e-6)
    want = opt.state_dict()
    assert sorted(exported["state"]) == sorted(want["state"])      # the empty parameter has no state on either side
    for i, st in want["state"].items():
        assert torch.equal(exported["state"][i]["exp_avg"], st["exp_avg"])
        assert torch.equal(exported["state"][i]["exp_avg_sq"], st["exp_avg_sq"])
        assert float(exported["state"][i]["step"]) == float(st["step"]) == 3.0
    fresh = torch.optim.AdamW(copy.deepcopy(net).parameters(), lr=1.0)
    fresh.load_state_dict(exported)                                # accepted by the stock class
    assert fresh.param_groups[0]["betas"] == (0.95, 0.999) and fresh.param_groups[0]["lr"] == 1e-4
    # EMA: ema_pytorch layout, round trip through the slabs
    for seg in segs:
        seg.ema.copy_(torch.randn_like(seg.ema))
    ema_sd = TS.export_ema_state(net, segs, steps)
    assert set(ema_sd) == {"ema_model." + k for k in net.state_dict()} | {"initted", "step"}
    assert bool(ema_sd["initted"]) and int(ema_sd["step"]) == 3
    # ema_pytorch 0.2.3 registers both bookkeeping buffers with shape [1] (float32 flag, int64 counter); the flag
    # is set by the SECOND update() call
    assert tuple(ema_sd["initted"].shape) == (1,) and ema_sd["initted"].dtype == torch.float32
    assert tuple(ema_sd["step"].shape) == (1,) and ema_sd["step"].dtype == torch.int64
    kept = [seg.ema.clone() for seg in segs]
    for seg in segs:
        seg.ema.zero_()
    TS.import_ema_state(net, segs, ema_sd)
    assert all(torch.equal(a.ema, b) for a, b in zip(segs, kept))
    assert TS.export_optimizer_state(net, segs, 0, lr=1e-4, betas=(0.95, 0.999), eps=1e-8, weight_decay=0)["state"] == {}


def test_launch_list_tail_keeps_tags_and_lanes():
    """`_Steps.tail(k)` (the single-graph predict_action starts each forward after the timestep MLP): same callables,
    tags and lanes from index k on; lane 1 = side stream, lane 2 = main-lane step that joins the side stream first."""
    from v2a_b200.policy_unet1d import _Steps
    s = _Steps()
    calls = []
    for i, lane in enumerate([0, 0, 1, 0, 2, 1]):
        s.add(f"step{i}", (lambda i=i: calls.append(i)), lane=lane)
    t = s.tail(2)
    assert len(t) == 4 and t.tags == ["step2", "step3", "step4", "step5"] and t.lanes == [1, 0, 2, 1]
    for fn in t:
        fn()
    assert calls == [2, 3, 4, 5]
    assert len(s) == 6 and s.lanes == [0, 0, 1, 0, 2, 1]          # the source list is untouched
    s.lane = 1
    s.add("default-lane", lambda: None)
    assert s.lanes[-1] == 1
