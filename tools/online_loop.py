"""BASELINE.json configs[4] in miniature: the online loop of `LB_Online_Trainer_V7`
(diffuser/libero/lb_online_trainer_v7.py:540-640 train, :862-960 video_guided_explore) with the simulator and CLIP
replaced by stubs — everything between them runs on the v2a_b200 paths:

  per task (tasks shard round-robin over the ranks, no collective):
      sub-goal video   GoalGaussianDiffusion.sample(x_cond, task_embed, batch_size=1)          (rows V1-V14)
      rollout          per sub-goal frame: policy.predict_action({obs, goal}, use_ddim=True)   (row N2)
                       -> stub environment "executes" the actions and renders uint8 frames
      replay           Global_EnvReplayBuffer_Img.add_one_episode(...)  (episodes in HBM)       (row N4)
  then K optimisation steps on every rank:
      batch            replay.sample_random_batch_seq(B) -> to_batch_dict                       (row N4)
      step             PolicyTrainStep.step(compute_loss): fwd + bwd + NCCL all-reduce of the gradient slabs
                       + clip + AdamW + EMA                                                     (rows P1-P8, N3)

A developer probe, not the contract bench: one JSON line from rank 0, phases timed with CUDA events, max over ranks.

usage:  python tools/online_loop.py [--tasks 8] [--iters 1] [--policy-steps 10] [--batch 256] [--denoise-steps 100]
        python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/online_loop.py ...
"""
import argparse
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

H = W = 128
FRAMES = 7


class StubEnv:
    """Stands in for one Libero environment: renders seeded uint8 frames [H, W, 3], ignores the physics."""

    def __init__(self, seed):
        self.rng = np.random.default_rng(seed)

    def render(self):
        return self.rng.integers(0, 256, size=(H, W, 3), dtype=np.uint8)

    def step(self, action):
        return self.render()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--tasks", type=int, default=8)
    ap.add_argument("--iters", type=int, default=1)
    ap.add_argument("--policy-steps", type=int, default=10)
    ap.add_argument("--batch", type=int, default=256)
    ap.add_argument("--denoise-steps", type=int, default=100)
    ap.add_argument("--exec-steps", type=int, default=4, help="actions executed per sub-goal frame")
    ap.add_argument("--batch-videos", action="store_true",
                    help="sample the sub-goal videos of all of a rank's tasks in one call instead of one by one")
    args = ap.parse_args()
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("online_loop.py needs a CUDA device: the v2a_b200 paths have no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    out = run_loop(args)
    if out is not None:
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


def run_loop(args):
    """The loop itself on an initialised process group (bench.py calls this in-process so that the N-GPU run of
    configs[4] is part of the driver-run line).  ``args``: namespace with tasks, iters, policy_steps, batch,
    denoise_steps, exec_steps, batch_videos.  Returns the result dict on rank 0, None elsewhere."""
    import torch.distributed as dist
    from v2a_b200.diffusion_policy import build_libero_policy
    from v2a_b200.goal_diffusion import GoalGaussianDiffusion
    from v2a_b200.replay import Global_EnvReplayBuffer_Img
    from v2a_b200.train_step import PolicyTrainStep
    from v2a_b200.unet import Unet_Libero

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))

    torch.manual_seed(0)                                     # same initial weights on every rank
    net = Unet_Libero()
    with torch.no_grad():
        for p in net.parameters():                           # no checkpoints offline: perturbed random init
            p.add_(0.02 * torch.randn_like(p))
    diff = GoalGaussianDiffusion(net, image_size=(H, W), channels=3 * FRAMES, timesteps=100,
                                 sampling_timesteps=args.denoise_steps, loss_type="l2", objective="pred_v",
                                 beta_schedule="cosine", min_snr_loss_weight=True, guidance_weight=0).cuda()
    policy = build_libero_policy().to("cuda")
    step = PolicyTrainStep(policy)
    T = policy.horizon
    tasks = [f"task_{i}" for i in range(args.tasks)]
    task_embed = {tk: torch.randn(1, 12, 512, generator=torch.Generator().manual_seed(i)).cuda()
                  for i, tk in enumerate(tasks)}             # stands in for the CLIP text encoder
    replay = Global_EnvReplayBuffer_Img(tasks, 1000, 800, T + 1, None, (H, W), env_buf_config={"sample_act_seq_len": T})
    np.random.seed(100 + rank)
    import random
    random.seed(200 + rank)
    # initial random-exploration episodes (h5_add_rand_act_episodes_to_Buf, trainer :718-780)
    rng = np.random.default_rng(300 + rank)
    for i, tk in enumerate(tasks):
        replay.add_one_episode(tk, "agentview", i, rng.integers(0, 256, size=(40, H, W, 3), dtype=np.uint8),
                               rng.uniform(-1, 1, size=(39, 7)).astype(np.float32))

    def ev():
        return torch.cuda.Event(enable_timing=True)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(v):
        if world == 1:
            return v
        t = torch.tensor([v], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def explore(only_first=False):
        """video_guided_explore for this rank's tasks: returns (#videos, #predict_action calls, #env frames).
        The reference generates the sub-goal videos one task at a time (bs = 1, trainer :880-891); --batch-videos
        samples all of a rank's tasks in ONE call (they are independent), which fills the GPU far better."""
        n_vid = n_act = n_frames = 0
        policy.eval()
        mine = [(i, tk) for i, tk in enumerate(tasks) if i % world == rank]
        if only_first:
            mine = mine[:1]
        envs = {i: StubEnv(1000 * (i + 1)) for i, _ in mine}
        first = {i: envs[i].render() for i, _ in mine}
        to_cond = lambda fr: torch.from_numpy(fr).permute(2, 0, 1).float() / 255.0
        videos = {}
        if args.batch_videos and len(mine) > 1:
            x_cond = torch.stack([to_cond(first[i]) for i, _ in mine]).cuda()
            te = torch.cat([task_embed[tk] for _, tk in mine], dim=0)
            out = diff.sample(x_cond, te, batch_size=len(mine))                    # [n, 21, H, W] in [0, 1]
            videos = {i: out[j] for j, (i, _) in enumerate(mine)}
        for i, tk in mine:
            env, frame = envs[i], first[i]
            if i not in videos:
                videos[i] = diff.sample(to_cond(frame)[None].cuda(), task_embed[tk], batch_size=1)[0]
            goals = videos[i].reshape(FRAMES, 3, H, W)
            n_vid += 1
            frames, acts = [frame], []
            with torch.no_grad():
                for f in range(FRAMES):
                    obs = {"img_obs_1": to_cond(frames[-1]).cuda()[None, None], "img_goal_1": goals[f][None, None]}
                    a = policy.predict_action(obs, use_ddim=True)["action"][0].cpu().numpy()   # [n_action_steps, 7]
                    n_act += 1
                    for k in range(min(args.exec_steps, len(a))):
                        frames.append(env.step(a[k]))
                        acts.append(np.clip(a[k], -1, 1).astype(np.float32))
            replay.add_one_episode(tk, "agentview", i, np.stack(frames), np.stack(acts))
            n_frames += len(frames)
        policy.train()
        return n_vid, n_act, n_frames

    loss_host = torch.zeros(1).pin_memory()

    def train(k):
        for _ in range(k):
            st, gl, acts, _, _ = replay.sample_random_batch_seq(args.batch)
            batch = {"obs": {"img_obs_1": st[:, None], "img_goal_1": gl[:, None]}, "action": acts}   # to_batch_dict
            loss = step.step(lambda: policy.compute_loss(batch))
        loss_host.copy_(loss.reshape(1), non_blocking=True)

    # warm-up: builds the plans / CUDA graphs of every engine at its batch size (B = 1 video + policy, B = batch train)
    explore(only_first=not args.batch_videos)     # batched: the timed call must find its B = len(tasks) engine built
    train(3)
    barrier()
    ms_explore = ms_train = 0.0
    counts = (0, 0, 0)
    for _ in range(args.iters):
        e0, e1, e2 = ev(), ev(), ev()
        barrier()
        e0.record()
        counts = explore()
        e1.record()
        train(args.policy_steps)
        e2.record()
        barrier()
        ms_explore += max_over_ranks(e0.elapsed_time(e1))
        ms_train += max_over_ranks(e1.elapsed_time(e2))
    n_vid = args.tasks * args.iters
    step.close()
    if rank == 0:
        return ({
            "what": "configs[4] in miniature: per task sample() at B=1 + predict_action rollout on a stub simulator + "
                    "replay insert, then policy optimisation steps fed by the HBM replay buffer (all-reduce for N>1)",
            "n_gpus": world, "tasks": args.tasks, "batch_videos": bool(args.batch_videos), "iters": args.iters, "denoise_steps": args.denoise_steps,
            "explore_ms_per_iter": ms_explore / args.iters,
            "videos_per_s": n_vid / (ms_explore * 1e-3), "video_frames_per_s": n_vid * FRAMES / (ms_explore * 1e-3),
            "predict_action_calls_per_task": counts[1] // max(1, counts[0]), "env_frames_per_task": counts[2] // max(1, counts[0]),
            "train_ms_per_step": ms_train / (args.iters * args.policy_steps),
            "policy_samples_per_s": args.batch * world * args.iters * args.policy_steps / (ms_train * 1e-3),
            "batch_per_gpu": args.batch, "policy_steps_per_iter": args.policy_steps,
            "replay_episodes": len(replay), "replay_bytes_in_hbm": replay.nbytes(), "loss": float(loss_host.item()),
            "stubs": "simulator (random uint8 frames), CLIP text encoder (random task embeddings), random-init weights"})
    return None


if __name__ == "__main__":
    main()
