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
