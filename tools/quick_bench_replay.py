"""Developer timing probe for row N4 (not the contract bench): the HBM-resident replay buffer.

usage: python tools/quick_bench_replay.py [B] [--step]
Times the device-side batch assembly alone (two gather kernels over uint8 episodes) against the reference's way
(stack B float frames on the host + host->device copy), and with --step the whole compute_loss optimisation step
fed either way.
"""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from v2a_b200.replay import Global_EnvReplayBuffer_Img  # noqa: E402


def timed(fn, n=20, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n, (time.perf_counter() - t0) * 1e3 / n


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 and sys.argv[1].isdigit() else 256
    T, A = 16, 7
    rng = np.random.default_rng(0)
    rb = Global_EnvReplayBuffer_Img(["t"], 64, 128, T + 1, None, (128, 128), env_buf_config={"sample_act_seq_len": T})
    host_eps = []
    for e in range(32):
        frames = rng.integers(0, 256, size=(64, 128, 128, 3), dtype=np.uint8)
        acts = rng.uniform(-1, 1, size=(63, A)).astype(np.float32)
        rb.add_one_episode("t", "c", e, frames, acts)
        host_eps.append(torch.from_numpy(frames).permute(0, 3, 1, 2).float() / 255.0)   # the reference's storage
    print(f"replay buffer: {len(rb)} episodes, {rb.nbytes() / 1e6:.1f} MB of HBM "
          f"(the reference's float frames: {sum(e.numel() * 4 for e in host_eps) / 1e6:.1f} MB of host memory)")
    plan = rb.plan_batch(B)
    out_imgs = torch.empty(2 * B, 3, 128, 128, device="cuda")
    out_acts = torch.empty(B, T, A, device="cuda")
    ms_k, _ = timed(lambda: rb.gather(plan, out_imgs, out_acts))
    moved = 2 * B * 128 * 128 * 3 * (1 + 4)
    print(f"B={B} gather kernels + address table: {ms_k * 1e3:.1f} us per batch -> {moved / ms_k / 1e6:.0f} GB/s "
          f"(3 B read + 12 B written per pixel)")
    ms_s, wall_s = timed(lambda: rb.sample_random_batch_seq(B))
    print(f"  sample_random_batch_seq (draws + table + kernels): device {ms_s * 1e3:.1f} us, host wall {wall_s * 1e3:.1f} us")

    def reference_way():
        st = torch.stack([host_eps[b][s] for b, s in zip(plan.buf_idxs, plan.start_idxs)])
        gl = torch.stack([host_eps[b][g] for b, g in zip(plan.buf_idxs, plan.goal_idxs)])
        return st.cuda(), gl.cuda()
    _, wall_r = timed(reference_way, n=5, warm=1)
    print(f"  the reference's way (torch.stack of {2 * B} float frames + .cuda()): host wall {wall_r:.2f} ms per batch")
    st, gl = reference_way()
    a, b, _ = rb.gather(plan)
    print("  bit-identical to the reference's batch:", bool(torch.equal(a, st) and torch.equal(b, gl)))

    if "--step" in sys.argv:
        from v2a_b200.diffusion_policy import build_libero_policy
        from v2a_b200.train_step import PolicyTrainStep
        torch.manual_seed(0)
        policy = build_libero_policy().to("cuda")
        policy.train()
        step = PolicyTrainStep(policy)
        loss_host = torch.zeros(1).pin_memory()

        def replay_step():
            s, g, acts, _, _ = rb.sample_random_batch_seq(B)
            batch = {"obs": {"img_obs_1": s[:, None], "img_goal_1": g[:, None]}, "action": acts}
            loss_host.copy_(step.step(lambda: policy.compute_loss(batch)).reshape(1), non_blocking=True)
        ms, wall = timed(replay_step, n=10, warm=3)
        print(f"  compute_loss optimisation step fed by the device replay buffer: {ms:.2f} ms (host wall {wall:.2f}) "
              f"-> {B / ms * 1e3:.0f} samples/s, loss {loss_host.item():.4f}")


if __name__ == "__main__":
    main()
