"""Developer timing probe: DiffusionUnetImagePolicy.predict_action (8-step DDIM) latency."""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from v2a_b200 import ops  # noqa: E402
from v2a_b200.diffusion_policy import build_libero_policy  # noqa: E402

torch.manual_seed(0)
pol = build_libero_policy().to("cuda").eval()
for B in (1, 8):
    obs = {"img_obs_1": torch.rand(B, 1, 3, 128, 128, device="cuda"), "img_goal_1": torch.rand(B, 1, 3, 128, 128, device="cuda")}
    with torch.no_grad():
        for _ in range(4):
            out = pol.predict_action(obs, use_ddim=True)
        torch.cuda.synchronize()
        n0 = ops.launch_count()
        t0 = time.perf_counter()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        n = 10
        for _ in range(n):
            out = pol.predict_action(obs, use_ddim=True)
        e1.record()
        torch.cuda.synchronize()
        wall = (time.perf_counter() - t0) / n * 1e3
    print(f"B={B}: predict_action {e0.elapsed_time(e1) / n:.3f} ms GPU, {wall:.3f} ms wall, "
          f"{(ops.launch_count() - n0) / n:.0f} eager v2a launches per call, action {tuple(out['action'].shape)}")

    # where the time goes: the captured graph alone (GPU time), and its pieces launched eagerly one by one
    from v2a_b200 import diffusion_policy as DP
    plan = list(DP._PREDICT_PLANS[pol].values())[-1]
    if plan.graph is not None:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(20):
            plan.graph.replay()
        e1.record()
        torch.cuda.synchronize()
        print(f"  graph replay alone: {e0.elapsed_time(e1) / 20:.3f} ms")

    def timed(fn, n=20):
        fn()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(n):
            fn()
        b.record()
        torch.cuda.synchronize()
        return a.elapsed_time(b) / n
    u = plan.unet
    print(f"  UNet1D forward graph (B={B}): {timed(lambda: u._run('fwd', u.fwd)):.3f} ms for {len(u.fwd)} launches; "
          f"encoder forward graph: {timed(lambda: plan.enc[0]._run('fwd', plan.enc[0].fwd, pre=(plan.enc[0].stats_arena,))):.3f} ms "
          f"for {len(plan.enc[0].fwd)} launches")
    rows = sorted(((timed(fn, 10), tag) for fn, tag in zip(u.fwd, u.fwd.tags)), reverse=True)
    print("  UNet1D forward, slowest eager launches: " + ", ".join(f"{t * 1e3:.0f}us {tag}" for t, tag in rows[:8]),
          f"| sum {sum(t for t, _ in rows):.3f} ms")
