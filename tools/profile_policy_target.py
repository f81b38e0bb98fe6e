"""ncu target: steady-state DiffusionUnetImagePolicy.compute_loss optimisation steps at B=256 (configs[2]).
  ncu --profile-from-start off ... python tools/profile_policy_target.py
Only the steps between cudaProfilerStart/Stop are profiled (warm-up, plan building and graph capture are not)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from v2a_b200.diffusion_policy import build_libero_policy  # noqa: E402
from v2a_b200.train_step import PolicyTrainStep  # noqa: E402

B = int(os.environ.get("B", "256"))
STEPS = int(os.environ.get("STEPS", "2"))
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
torch.manual_seed(0)
policy = build_libero_policy().to("cuda")
policy.train()
batch = {"obs": {"img_obs_1": torch.rand(B, 1, 3, 128, 128, device="cuda"),
                 "img_goal_1": torch.rand(B, 1, 3, 128, 128, device="cuda")},
         "action": torch.rand(B, 16, 7, device="cuda") * 2 - 1}
step = PolicyTrainStep(policy)
for _ in range(4):
    step.step(lambda: policy.compute_loss(batch))
torch.cuda.synchronize()
torch.cuda.profiler.start()
for _ in range(STEPS):
    loss = step.step(lambda: policy.compute_loss(batch))
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("done", float(loss))
