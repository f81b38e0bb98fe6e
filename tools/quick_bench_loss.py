"""Developer probe: the whole policy optimisation step (compute_loss -> backward -> clip/AdamW/EMA) at B = 256 with the
batch resident in HBM, CUDA-event timed.  usage: python tools/quick_bench_loss.py [steps]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from v2a_b200.diffusion_policy import build_libero_policy  # noqa: E402
from v2a_b200.train_step import PolicyTrainStep  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 30
B = 256
torch.manual_seed(77)
policy = build_libero_policy().to("cuda")
policy.train()
g = torch.Generator().manual_seed(2000)
batch = {"obs": {"img_obs_1": torch.rand(B, 1, 3, 128, 128, generator=g).cuda(),
                 "img_goal_1": torch.rand(B, 1, 3, 128, 128, generator=g).cuda()},
         "action": (torch.rand(B, 16, 7, generator=g) * 2 - 1).cuda()}
step = PolicyTrainStep(policy)
for _ in range(5):
    loss = step.step(lambda: policy.compute_loss(batch))
torch.cuda.synchronize()
best = 1e9
for rep in range(3):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        loss = step.step(lambda: policy.compute_loss(batch))
    e1.record()
    torch.cuda.synchronize()
    best = min(best, e0.elapsed_time(e1) / steps)
print(f"compute_loss step {best:.3f} ms  ({B / best * 1e3:.0f} samples/s)  loss {float(loss):.6f}")
