"""Row N4 on the GPU: batches assembled by the gather kernels from HBM-resident uint8 episodes equal the
reference replay buffer's CPU batches bit for bit (golden from the unmodified reference class)."""
import os
import random

import numpy as np
import pytest
import torch

from tests.golden.configs import REPLAY
from tests.test_replay_host import GOLD, build_buffer, host_gather

pytestmark = pytest.mark.gpu


def test_batches_equal_reference_bit_for_bit():
    gold = torch.load(GOLD)
    buf = build_buffer("cuda")
    np.random.seed(REPLAY["np_seed"])
    random.seed(REPLAY["py_seed"])
    for draw in gold["draws"]:
        st, gl, acts, tasks, info = buf.sample_random_batch_seq(REPLAY["batch"])
        assert st.is_cuda and st.dtype == torch.float32 and st.shape == (REPLAY["batch"], 3, REPLAY["H"], REPLAY["W"])
        assert torch.equal(st.cpu(), draw["imgs_start_u8"].float() / 255.0)
        assert torch.equal(gl.cpu(), draw["imgs_goal_u8"].float() / 255.0)
        assert torch.equal(st.cpu()[:, :, 3, 5], draw["start_f32_sample"])
        assert torch.equal(acts.cpu(), draw["acts"])
        assert tasks == draw["tasks"] and info["cams_str"] == draw["cams"]
        assert info["env_idxs"].tolist() == draw["env_idxs"].tolist()


@pytest.mark.parametrize("H,W", [(128, 128), (5, 7), (3, 128), (64, 33)])
def test_gather_every_byte_value_and_ragged_shapes(H, W):
    """All 256 byte values through the IEEE division; shapes that exercise the 16-byte and the byte-wise
    staging paths and the last partial row block."""
    from v2a_b200.replay import Global_EnvReplayBuffer_Img
    rng = np.random.default_rng(H * 1000 + W)
    buf = Global_EnvReplayBuffer_Img(["t"], 8, 64, 2, None, (H, W), env_buf_config={"sample_act_seq_len": 3})
    for e in range(3):
        T = 9 + e
        frames = rng.integers(0, 256, size=(T, H, W, 3), dtype=np.uint8)
        frames.reshape(-1)[:256] = np.arange(256, dtype=np.uint8)
        buf.add_one_episode("t", "c", e, frames, rng.uniform(-1, 1, size=(T - 1, 7)).astype(np.float32))
    np.random.seed(1)
    random.seed(2)
    plan = buf.plan_batch(37)
    st, gl, acts = buf.gather(plan)
    est, egl, eacts = host_gather(buf, plan)
    assert torch.equal(st.cpu(), est.cpu().permute(0, 3, 1, 2).float() / 255.0)
    assert torch.equal(gl.cpu(), egl.cpu().permute(0, 3, 1, 2).float() / 255.0)
    assert torch.equal(acts.cpu(), eacts.cpu())


def test_full_size_batch_feeds_compute_loss_shapes():
    """B = 256 at 128 x 128 (configs[2]): the trainer's `to_batch_dict` views (lb_online_trainer_v7.py:1296-1310)."""
    from v2a_b200.replay import Global_EnvReplayBuffer_Img
    rng = np.random.default_rng(5)
    buf = Global_EnvReplayBuffer_Img(["t"], 64, 128, 20, None, (128, 128), env_buf_config={"sample_act_seq_len": 16})
    for e in range(6):
        buf.add_one_episode("t", "c", e, rng.integers(0, 256, size=(40, 128, 128, 3), dtype=np.uint8),
                            rng.uniform(-1, 1, size=(39, 7)).astype(np.float32))
    np.random.seed(0)
    random.seed(0)
    plan = buf.plan_batch(256)
    st, gl, acts = buf.gather(plan)
    est, egl, eacts = host_gather(buf, plan)
    assert st.shape == (256, 3, 128, 128) and acts.shape == (256, 16, 7)
    # expectation on the CPU, where the reference divides (torch's CUDA `x / scalar` multiplies by 1 / scalar)
    assert torch.equal(st.cpu(), est.cpu().permute(0, 3, 1, 2).float() / 255.0)
    assert torch.equal(gl.cpu(), egl.cpu().permute(0, 3, 1, 2).float() / 255.0)
    assert torch.equal(acts, eacts)
    assert st[:, None].shape == (256, 1, 3, 128, 128)
