"""Replay buffer with the episodes resident in HBM.

Mirrors ``Global_EnvReplayBuffer_Img`` / ``EnvImg_UnitBuffer`` of the reference
(diffuser/datasets/env_img_replay_buffer.py:10-116,214-309): same constructor arguments, same
``add_one_episode`` / ``sample_random_batch_seq`` / ``__len__`` / ``__getitem__`` / ``is_full`` surface, the same
bounded-deque eviction and — draw for draw — the same use of the global ``numpy.random`` and ``random``
generators, so a seeded run picks exactly the (episode, start frame) pairs the reference picks.

What changes is where the bytes live.  The reference keeps every frame as a CPU float tensor [3, H, W]
(``img_np_toTensor``, diffuser/datasets/img_utils.py:27-37: uint8 HWC -> float CHW / 255), stacks 2 x B of them in a
Python loop per optimisation step and copies 2 x B x 3 x H x W floats to the GPU (lb_online_trainer_v7.py:586).
Here an episode is ONE uint8 [T, H, W, 3] tensor in HBM (49 KB per 128 x 128 frame) plus its fp32 actions, and
a step's batch is two kernel launches over a 6 KB table of device addresses (``v2a_replay_gather_images`` /
``v2a_replay_gather_actions``): ``float(u8) / 255`` with IEEE division reproduces the reference's float frames
bit for bit, the host->device traffic of a step drops from 100.7 MB (B = 256) to the address table.

Frames given as float tensors are accepted only when they are exactly ``u8 / 255`` (what every producer in the
reference emits, img_utils.py:37); anything else raises instead of being stored lossily.  The host-side planning
(``plan_batch``) runs without CUDA (tests); assembling the batch requires it — there is no CPU path.
"""
from __future__ import annotations

import random
from collections import deque
from typing import NamedTuple, Optional

import numpy as np
import torch

from . import _lib


def frames_to_u8(imgs) -> torch.Tensor:
    """Frames of one episode -> uint8 [T, H, W, 3] on the CPU, lossless or ValueError.

    Accepts a uint8 array / tensor [T, H, W, 3] (what the simulator and the H5 files hold,
    lb_online_trainer_v7.py:741) or the reference's list of float [3, H, W] tensors in [0, 1].
    """
    if isinstance(imgs, np.ndarray):
        imgs = torch.from_numpy(np.ascontiguousarray(imgs))
    if torch.is_tensor(imgs) and imgs.dtype == torch.uint8:
        if imgs.ndim != 4 or imgs.shape[-1] != 3:
            raise ValueError(f"uint8 frames must be [T, H, W, 3], got {tuple(imgs.shape)}")
        return imgs.cpu().contiguous()
    if isinstance(imgs, (list, tuple)):
        imgs = torch.stack([torch.as_tensor(i) for i in imgs], dim=0)
    if not torch.is_tensor(imgs) or imgs.ndim != 4 or imgs.shape[1] != 3 or not imgs.is_floating_point():
        raise ValueError("frames must be uint8 [T, H, W, 3] or float [T, 3, H, W] in [0, 1]")
    x = imgs.detach().cpu().to(torch.float32)
    u8 = torch.round(x * 255.0).clamp_(0, 255).to(torch.uint8)
    if not torch.equal(u8.to(torch.float32) / 255.0, x):
        raise ValueError("float frames are not exactly uint8 / 255 (img_np_toTensor); refusing a lossy uint8 store")
    return u8.permute(0, 2, 3, 1).contiguous()


class BatchPlan(NamedTuple):
    """Host-side result of one ``sample_random_batch_seq`` draw (what the RNGs decided)."""
    buf_idxs: np.ndarray     # [B] index into the deque of unit buffers
    start_idxs: np.ndarray   # [B] frame index of the start image inside its unit buffer
    goal_idxs: np.ndarray    # [B] = start + act_seq_len


class EnvImg_UnitBuffer:
    """One episode of one environment (env_img_replay_buffer.py:214-309), frames uint8 HWC on ``device``."""

    def __init__(self, max_len, task_name, cam_name, env_idx, per_sample_gap=1, max_start_id=None, device="cuda"):
        self.max_len = max_len
        self.per_sample_gap = per_sample_gap
        assert self.per_sample_gap == 1
        assert type(task_name) == str
        self.task_name = task_name
        self.cam_name = cam_name
        self.env_idx = env_idx
        self.device = torch.device(device)
        self.frames: Optional[torch.Tensor] = None   # uint8 [T, H, W, 3]
        self.acts: Optional[torch.Tensor] = None     # fp32 [T - 1, A]

    def push_seq(self, new_imgs, new_acts):
        """Append frames / actions; on a non-empty buffer the first new frame repeats the last stored one
        (env_img_replay_buffer.py:253-276).  Both deques are bounded: the oldest entries fall off the left."""
        u8 = frames_to_u8(new_imgs)
        acts = new_acts
        if isinstance(acts, (list, tuple)):
            acts = torch.stack([torch.as_tensor(a) for a in acts], dim=0)
        acts = torch.as_tensor(acts).detach().cpu().to(torch.float32)
        assert acts.ndim == 2 and len(acts) == len(u8) - 1
        assert 2 <= len(u8) <= 800, 'can be large when video rollout'
        u8 = u8.to(self.device, non_blocking=False)
        acts = acts.contiguous().to(self.device)
        if self.frames is None:
            frames, all_acts = u8, acts
        else:
            assert self.frames.shape[1:] == u8.shape[1:]
            frames = torch.cat([self.frames, u8[1:]], dim=0)
            all_acts = torch.cat([self.acts, acts], dim=0)
        # deque(maxlen=max_len) for the frames, deque(maxlen=max_len - 1) for the actions
        self.frames = frames[-self.max_len:].contiguous()
        self.acts = all_acts[-(self.max_len - 1):].contiguous() if self.max_len > 1 else all_acts[:0]
        assert len(self.frames) <= self.max_len

    def sample_start(self, act_seq_len: int) -> int:
        """The one RNG draw of ``sample_seq`` (env_img_replay_buffer.py:287-291)."""
        cur_len = len(self)
        assert act_seq_len < cur_len
        return random.randint(0, cur_len - act_seq_len - 1)   # [a, b] inclusive

    def _float_frames(self, idx) -> torch.Tensor:
        """Frames as the reference stores them: CPU float [.., 3, H, W] = u8 / 255 (divided on the CPU, like
        img_np_toTensor; torch's CUDA division by a scalar multiplies by the reciprocal instead)."""
        f = self.frames[idx].cpu()
        return f.permute(*range(f.dim() - 3), -1, -3, -2).float() / 255.0

    @property
    def imgs_buf(self):
        """The reference's attribute (a deque of CPU float [3, H, W] tensors), materialised on demand — the trainer
        only reads it in its debug image dumps (lb_online_trainer_v7.py:541-546)."""
        return [] if self.frames is None else list(torch.unbind(self._float_frames(slice(None)), dim=0))

    def sample_seq(self, act_seq_len):
        """Reference-format single sample (env_img_replay_buffer.py:278-302): CPU tensors, same RNG draw."""
        start_idx = self.sample_start(act_seq_len)
        goal_idx = start_idx + act_seq_len
        ret_acts = self.acts[start_idx:goal_idx].cpu()
        assert len(ret_acts) == act_seq_len
        return self._float_frames(start_idx), self._float_frames(goal_idx), ret_acts, self.task_name, self.env_idx

    def __len__(self):
        return 0 if self.frames is None else int(self.frames.shape[0])


class Global_EnvReplayBuffer_Img:
    """Deque of unit buffers with device-side batch assembly (env_img_replay_buffer.py:10-116)."""

    def __init__(self, task_list, max_num_unitBufs, max_len_uB, min_len_uB, env_list=None, render_img_size=None,
                 env_buf_config={}, device="cuda"):
        self.buffers: deque = deque(maxlen=max_num_unitBufs)
        self.task_list = task_list
        self.env_list = env_list
        self.camera_list = getattr(env_list, "camera_list", None)
        self.bufs_task = deque(maxlen=max_num_unitBufs)
        self.bufs_cam = deque(maxlen=max_num_unitBufs)
        self.num_cams = len(self.camera_list) if self.camera_list is not None else 0
        self.render_img_size = render_img_size
        self.sample_act_seq_len = env_buf_config['sample_act_seq_len']
        self.per_sample_gap = 1
        self.max_num_unitBuf = max_num_unitBufs
        self.max_len_uB = max_len_uB
        self.min_len_uB = min_len_uB
        self.cnt_all_history_episodes = 0
        assert max_num_unitBufs <= 1e4
        self.device = torch.device(device)

    # -- filling ---------------------------------------------------------------------------------------------
    def add_one_episode(self, tk: str, cam_name, env_idx, imgs, acts, is_suc=False):
        """Add a whole episode (env_img_replay_buffer.py:44-63); ``imgs`` may also be uint8 [T, H, W, 3]."""
        assert len(self.buffers) <= self.max_num_unitBuf
        assert len(imgs) == len(acts) + 1
        tmp_buf = EnvImg_UnitBuffer(self.max_len_uB, tk, cam_name=cam_name, env_idx=env_idx,
                                    per_sample_gap=self.per_sample_gap, device=self.device)
        tmp_buf.push_seq(new_imgs=imgs, new_acts=acts)
        assert self.min_len_uB <= len(tmp_buf)
        self.buffers.append(tmp_buf)
        self.bufs_task.append(tk)
        self.bufs_cam.append(cam_name)
        self.cnt_all_history_episodes += 1

    # -- sampling --------------------------------------------------------------------------------------------
    def plan_batch(self, batch_size: int) -> BatchPlan:
        """The random draws of ``sample_random_batch_seq`` in the reference's order: one
        ``np.random.randint(0, n_buffers, size=B)``, then one ``random.randint`` per sample
        (env_img_replay_buffer.py:80,91 -> :289).  Pure host code."""
        assert len(self.buffers) == len(self.bufs_task) == len(self.bufs_cam)
        cur_len = len(self.buffers)
        buf_idxs = np.random.randint(0, cur_len, size=batch_size)
        starts = np.empty(batch_size, dtype=np.int64)
        bufs = list(self.buffers)                 # deque indexing walks from an end: O(n) per access
        for i, bf_idx in enumerate(buf_idxs.tolist()):
            starts[i] = bufs[bf_idx].sample_start(self.sample_act_seq_len)
        return BatchPlan(buf_idxs, starts, starts + self.sample_act_seq_len)

    def address_table(self, plan: BatchPlan) -> np.ndarray:
        """int64 [3B]: device addresses of the start frames, the goal frames and the first action rows."""
        n = len(self.buffers)
        fbase = np.fromiter((b.frames.data_ptr() for b in self.buffers), dtype=np.int64, count=n)
        fstep = np.fromiter((b.frames.stride(0) for b in self.buffers), dtype=np.int64, count=n)   # uint8: bytes
        abase = np.fromiter((b.acts.data_ptr() for b in self.buffers), dtype=np.int64, count=n)
        astep = np.fromiter((b.acts.stride(0) * 4 for b in self.buffers), dtype=np.int64, count=n)
        idx = np.asarray(plan.buf_idxs, dtype=np.int64)
        start = np.asarray(plan.start_idxs, dtype=np.int64)
        goal = np.asarray(plan.goal_idxs, dtype=np.int64)
        return np.concatenate([fbase[idx] + start * fstep[idx], fbase[idx] + goal * fstep[idx],
                               abase[idx] + start * astep[idx]])

    def gather(self, plan: BatchPlan, out_imgs: Optional[torch.Tensor] = None,
               out_acts: Optional[torch.Tensor] = None):
        """Assemble the planned batch on the device: (imgs_start, imgs_goal, acts) as CUDA tensors."""
        if self.device.type != "cuda":
            raise RuntimeError("v2a_b200.replay: batch assembly runs on the GPU only (buffer device is "
                               f"{self.device}); there is no CPU path")
        B = len(plan.buf_idxs)
        T = self.sample_act_seq_len
        first = self.buffers[plan.buf_idxs[0]] if B else self.buffers[0]
        H, W = int(first.frames.shape[1]), int(first.frames.shape[2])
        A = int(first.acts.shape[1])
        for i in np.unique(plan.buf_idxs):
            f = self.buffers[i].frames
            assert tuple(f.shape[1:]) == (H, W, 3) and self.buffers[i].acts.shape[1] == A
        if out_imgs is None:
            out_imgs = torch.empty(2 * B, 3, H, W, dtype=torch.float32, device=self.device)
        if out_acts is None:
            out_acts = torch.empty(B, T, A, dtype=torch.float32, device=self.device)
        assert out_imgs.is_contiguous() and tuple(out_imgs.shape) == (2 * B, 3, H, W)
        assert out_acts.is_contiguous() and tuple(out_acts.shape) == (B, T, A)
        if B == 0:
            return out_imgs[:0], out_imgs[0:], out_acts
        # 24 B per sample from pageable memory: staged by the driver before the call returns, so the numpy
        # array may die right away; the device copy must outlive the (asynchronous) launches below, which the
        # caching allocator guarantees — it reuses the block only for work queued on this stream after them
        table = torch.from_numpy(self.address_table(plan)).to(self.device)
        lib = _lib.load()
        stream = torch.cuda.current_stream(self.device).cuda_stream
        _lib.check(lib.v2a_replay_gather_images(table.data_ptr(), 2 * B, H, W, out_imgs.data_ptr(), stream),
                   "replay_gather_images")
        _lib.check(lib.v2a_replay_gather_actions(table.data_ptr() + 2 * B * 8, B, T, A, out_acts.data_ptr(), stream),
                   "replay_gather_actions")
        return out_imgs[:B], out_imgs[B:], out_acts

    def sample_random_batch_seq(self, batch_size):
        """Same return tuple as the reference (env_img_replay_buffer.py:68-116) —
        ``imgs_start, imgs_goal [B, 3, H, W]``, ``acts [B, act_len, A]``, ``tasks_str``, ``tmp_info`` — with the
        tensors already on the GPU (the trainer's ``to_device_tp`` becomes a no-op)."""
        plan = self.plan_batch(batch_size)
        imgs_start, imgs_goal, acts_seq = self.gather(plan)
        tasks_str = [self.buffers[i].task_name for i in plan.buf_idxs]
        env_idxs = [self.buffers[i].env_idx for i in plan.buf_idxs]
        cams_str = [self.buffers[i].cam_name for i in plan.buf_idxs]
        tmp_info = dict(env_idxs=np.array(env_idxs), cams_str=cams_str)
        assert imgs_start.ndim == 4 and acts_seq.ndim == 3
        return imgs_start, imgs_goal, acts_seq, tasks_str, tmp_info

    # -- container surface -----------------------------------------------------------------------------------
    def __len__(self):
        return len(self.buffers)

    def __getitem__(self, idx):
        if idx >= len(self.buffers):
            raise IndexError("index out of range")
        return self.buffers[idx]

    def is_full(self):
        return len(self.buffers) == self.max_num_unitBuf

    def nbytes(self) -> int:
        """Bytes of HBM (or host memory for a CPU-planned buffer) the stored episodes occupy."""
        return sum(b.frames.numel() + b.acts.numel() * 4 for b in self.buffers)
