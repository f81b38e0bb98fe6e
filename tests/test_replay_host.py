"""Row N4 host logic (no GPU): the HBM-resident replay buffer makes the reference's random draws in the
reference's order and stores frames losslessly.  Golden = the UNMODIFIED reference class
(tests/golden/make_replay_golden.py)."""
import os
import random

import numpy as np
import pytest
import torch

from tests.golden.configs import REPLAY, replay_episodes

GOLD = os.path.join(os.path.dirname(__file__), "golden", "replay_golden.pt")


def build_buffer(device, float_frames=False):
    from v2a_b200.replay import Global_EnvReplayBuffer_Img
    c = REPLAY
    buf = Global_EnvReplayBuffer_Img(["task_0", "task_1", "task_2"], c["max_num_unitBufs"], c["max_len_uB"],
                                     c["min_len_uB"], None, (c["H"], c["W"]),
                                     env_buf_config={"sample_act_seq_len": c["act_seq_len"]}, device=device)
    for tk, cam, env_idx, frames, acts in replay_episodes():
        if float_frames:   # the reference's producer format: list of float [3, H, W] = u8 / 255
            imgs = list(torch.unbind(torch.from_numpy(frames.copy()).permute(0, 3, 1, 2).float() / 255.0, dim=0))
            buf.add_one_episode(tk, cam, env_idx, imgs, list(torch.unbind(torch.from_numpy(acts), dim=0)))
        else:
            buf.add_one_episode(tk, cam, env_idx, frames, acts)
    return buf


def host_gather(buf, plan):
    """Test-side restatement of what the gather kernels must produce."""
    T = buf.sample_act_seq_len
    st = torch.stack([buf.buffers[b].frames[s] for b, s in zip(plan.buf_idxs, plan.start_idxs)])
    gl = torch.stack([buf.buffers[b].frames[g] for b, g in zip(plan.buf_idxs, plan.goal_idxs)])
    acts = torch.stack([buf.buffers[b].acts[s:s + T] for b, s in zip(plan.buf_idxs, plan.start_idxs)])
    return st, gl, acts


@pytest.mark.parametrize("float_frames", [False, True])
def test_draws_and_eviction_match_reference(float_frames):
    gold = torch.load(GOLD)
    buf = build_buffer("cpu", float_frames)
    assert len(buf) == gold["len"] and buf.cnt_all_history_episodes == gold["cnt"]
    assert [len(b) for b in buf.buffers] == gold["unit_lens"]          # deque eviction + max_len truncation
    assert buf.is_full()
    np.random.seed(REPLAY["np_seed"])
    random.seed(REPLAY["py_seed"])
    for draw in gold["draws"]:
        plan = buf.plan_batch(REPLAY["batch"])
        st, gl, acts = host_gather(buf, plan)
        assert torch.equal(st.permute(0, 3, 1, 2), draw["imgs_start_u8"])
        assert torch.equal(gl.permute(0, 3, 1, 2), draw["imgs_goal_u8"])
        assert torch.equal(acts, draw["acts"])
        assert [buf.buffers[i].task_name for i in plan.buf_idxs] == draw["tasks"]
        assert [buf.buffers[i].env_idx for i in plan.buf_idxs] == draw["env_idxs"].tolist()
        assert [buf.buffers[i].cam_name for i in plan.buf_idxs] == draw["cams"]
        # u8 / 255 in fp32 is the reference's float frame, bit for bit
        assert torch.equal(st.permute(0, 3, 1, 2)[:, :, 3, 5].float() / 255.0, draw["start_f32_sample"])


def test_address_table_points_at_the_planned_rows():
    buf = build_buffer("cpu")
    np.random.seed(3)
    random.seed(4)
    plan = buf.plan_batch(8)
    table = buf.address_table(plan)
    B = 8
    for i in range(B):
        b = buf.buffers[plan.buf_idxs[i]]
        fb = REPLAY["H"] * REPLAY["W"] * 3
        assert table[i] == b.frames.data_ptr() + plan.start_idxs[i] * fb
        assert table[B + i] == b.frames.data_ptr() + plan.goal_idxs[i] * fb
        assert table[2 * B + i] == b.acts.data_ptr() + plan.start_idxs[i] * REPLAY["A"] * 4
        assert plan.goal_idxs[i] == plan.start_idxs[i] + REPLAY["act_seq_len"] < len(b)


def test_lossy_float_frames_are_refused():
    from v2a_b200.replay import frames_to_u8
    x = torch.rand(3, 3, 8, 8)
    with pytest.raises(ValueError):
        frames_to_u8(list(torch.unbind(x, 0)))
    u8 = torch.randint(0, 256, (3, 8, 8, 3), dtype=torch.uint8)
    assert torch.equal(frames_to_u8(u8.permute(0, 3, 1, 2).float() / 255.0), u8)
    assert torch.equal(frames_to_u8(u8.numpy()), u8)
    with pytest.raises(ValueError):
        frames_to_u8(torch.zeros(3, 3, 8, 8, dtype=torch.uint8))      # CHW uint8 is not the wire format


def test_push_seq_continues_an_episode_like_the_deques():
    """Second push drops its first frame (it repeats the last stored one) and both deques stay bounded
    (env_img_replay_buffer.py:253-276)."""
    from collections import deque
    from v2a_b200.replay import EnvImg_UnitBuffer
    rng = np.random.default_rng(0)
    a = rng.integers(0, 256, size=(6, 4, 4, 3), dtype=np.uint8)
    b = rng.integers(0, 256, size=(5, 4, 4, 3), dtype=np.uint8)
    b[0] = a[-1]
    aa, ab = rng.random((5, 2), dtype=np.float32), rng.random((4, 2), dtype=np.float32)
    ub = EnvImg_UnitBuffer(8, "t", "c", 0, device="cpu")
    ub.push_seq(a, aa)
    ub.push_seq(b, ab)
    imgs, acts = deque(maxlen=8), deque(maxlen=7)
    imgs.extend(a); acts.extend(aa); imgs.extend(b[1:]); acts.extend(ab)
    assert np.array_equal(ub.frames.numpy(), np.stack(imgs)) and np.array_equal(ub.acts.numpy(), np.stack(acts))


def test_unit_buffer_reference_views():
    """`imgs_buf` / `sample_seq` hand back what the reference's unit buffer holds (CPU float CHW = u8 / 255)."""
    buf = build_buffer("cpu")
    ub = buf[1]
    frames = ub.frames
    imgs = ub.imgs_buf
    assert len(imgs) == len(ub) and imgs[0].shape == (3, REPLAY["H"], REPLAY["W"])
    assert torch.equal(imgs[3], frames[3].permute(2, 0, 1).float() / 255.0)
    random.seed(5)
    st, gl, acts, tk, env_idx = ub.sample_seq(REPLAY["act_seq_len"])
    random.seed(5)
    s = random.randint(0, len(ub) - REPLAY["act_seq_len"] - 1)
    assert torch.equal(st, imgs[s]) and torch.equal(gl, imgs[s + REPLAY["act_seq_len"]])
    assert torch.equal(acts, ub.acts[s:s + REPLAY["act_seq_len"]]) and (tk, env_idx) == (ub.task_name, ub.env_idx)
    with pytest.raises(IndexError):
        buf[len(buf)]


def test_no_cpu_batch_assembly():
    buf = build_buffer("cpu")
    with pytest.raises(RuntimeError, match="GPU only"):
        buf.sample_random_batch_seq(4)


def test_against_live_reference_class():
    """When /root/reference is mounted: a longer seeded scenario (evictions between draws) against the real class."""
    from oracle import ref_import as R
    if not R.available():
        pytest.skip("reference checkout not mounted")
    import types
    from v2a_b200 import install
    from v2a_b200.replay import Global_EnvReplayBuffer_Img
    mod = R.replay_buffer_module()
    ref = mod.Global_EnvReplayBuffer_Img(["a"], 5, 20, 6, types.SimpleNamespace(camera_list=["c"]), (8, 8),
                                         env_buf_config={"sample_act_seq_len": 5})
    mine = Global_EnvReplayBuffer_Img(["a"], 5, 20, 6, None, (8, 8), env_buf_config={"sample_act_seq_len": 5},
                                      device="cpu")
    rng = np.random.default_rng(9)
    np.random.seed(21)
    random.seed(22)
    for round_ in range(4):
        for e in range(3):
            T = int(rng.integers(6, 30))
            frames = rng.integers(0, 256, size=(T, 8, 8, 3), dtype=np.uint8)
            acts = rng.uniform(-1, 1, size=(T - 1, 7)).astype(np.float32)
            imgs = list(torch.unbind(torch.from_numpy(frames.copy()).permute(0, 3, 1, 2).float() / 255.0, dim=0))
            ref.add_one_episode("a", "c", round_ * 3 + e, imgs, list(torch.unbind(torch.from_numpy(acts), dim=0)))
            mine.add_one_episode("a", "c", round_ * 3 + e, frames, acts)
        st_np, st_py = np.random.get_state(), random.getstate()
        rst, rgl, racts, rtasks, rinfo = ref.sample_random_batch_seq(16)
        np.random.set_state(st_np)
        random.setstate(st_py)
        plan = mine.plan_batch(16)
        st, gl, acts = host_gather(mine, plan)
        assert torch.equal(st.permute(0, 3, 1, 2).float() / 255.0, rst)
        assert torch.equal(gl.permute(0, 3, 1, 2).float() / 255.0, rgl)
        assert torch.equal(acts, racts)
        assert [mine.buffers[i].env_idx for i in plan.buf_idxs] == rinfo["env_idxs"].tolist()
        assert len(mine) == len(ref) and mine.cnt_all_history_episodes == ref.cnt_all_history_episodes
    try:
        done = install.install()
        assert "diffuser.datasets.env_img_replay_buffer.Global_EnvReplayBuffer_Img" in done
        assert mod.Global_EnvReplayBuffer_Img is Global_EnvReplayBuffer_Img
    finally:
        install.uninstall()
    assert mod.Global_EnvReplayBuffer_Img is not Global_EnvReplayBuffer_Img
