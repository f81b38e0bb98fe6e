"""Golden batches of the UNMODIFIED reference replay buffer (row N4; build container only:
python tests/golden/make_replay_golden.py).  Seeded synthetic episodes go through the reference's own
`img_np_toTensor` and `Global_EnvReplayBuffer_Img`; two consecutive `sample_random_batch_seq` draws are stored
(frames back as uint8 — exact, they are u8 / 255 — plus the float bit pattern of a few pixels)."""
import os
import random
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import ref_import  # noqa: E402
from tests.golden.configs import REPLAY, replay_episodes  # noqa: E402


def main():
    mod = ref_import.replay_buffer_module()
    to_tensor = ref_import.img_utils_module().img_np_toTensor
    c = REPLAY
    env_list = types.SimpleNamespace(camera_list=["cam_0", "cam_1"])
    buf = mod.Global_EnvReplayBuffer_Img(["task_0", "task_1", "task_2"], c["max_num_unitBufs"], c["max_len_uB"],
                                         c["min_len_uB"], env_list, (c["H"], c["W"]),
                                         env_buf_config={"sample_act_seq_len": c["act_seq_len"]})
    for tk, cam, env_idx, frames, acts in replay_episodes():
        imgs = list(torch.unbind(to_tensor(frames), dim=0))          # what rendered_imgs_preproc_fn returns
        buf.add_one_episode(tk, cam, env_idx, imgs, list(torch.unbind(torch.from_numpy(acts), dim=0)))
    np.random.seed(c["np_seed"])
    random.seed(c["py_seed"])
    draws = []
    for _ in range(2):
        st, gl, acts, tasks, info = buf.sample_random_batch_seq(c["batch"])
        draws.append({
            "imgs_start_u8": (st * 255).round().to(torch.uint8), "imgs_goal_u8": (gl * 255).round().to(torch.uint8),
            "start_f32_sample": st[:, :, 3, 5].clone(), "acts": acts.clone(), "tasks": list(tasks),
            "env_idxs": torch.from_numpy(info["env_idxs"].copy()), "cams": list(info["cams_str"]),
        })
        assert torch.equal(draws[-1]["imgs_start_u8"].float() / 255.0, st)
    torch.save({"draws": draws, "len": len(buf), "cnt": buf.cnt_all_history_episodes,
                "unit_lens": [len(b) for b in buf.buffers]}, os.path.join(HERE, "replay_golden.pt"))
    print("buffers", len(buf), [len(b) for b in buf.buffers], "batch", tuple(draws[0]["imgs_start_u8"].shape))


if __name__ == "__main__":
    main()
