"""Profiling target: ONE steady-state predict_action call at B = 1 (the single CUDA graph) between
cudaProfilerStart/Stop.  Run under `ncu --profile-from-start off` (kernel nodes of a graph replay are profiled one by
one); numbers printed under a profiler are not bench values."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from v2a_b200.diffusion_policy import build_libero_policy  # noqa: E402

torch.manual_seed(0)
pol = build_libero_policy().to("cuda").eval()
B = int(os.environ.get("B", "1"))
obs = {"img_obs_1": torch.rand(B, 1, 3, 128, 128, device="cuda"), "img_goal_1": torch.rand(B, 1, 3, 128, 128, device="cuda")}
with torch.no_grad():
    for _ in range(4):
        out = pol.predict_action(obs, use_ddim=True)
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    out = pol.predict_action(obs, use_ddim=True)
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
print("done", tuple(out["action"].shape))
