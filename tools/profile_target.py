"""Short profiling target: config-2 geometry (B=16, 128x128x7), a few denoise steps.
Run under ncu (see profiles/README.md); numbers printed under a profiler are not bench values."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from v2a_b200.goal_diffusion import GoalGaussianDiffusion  # noqa: E402
from v2a_b200.unet import Unet_Libero  # noqa: E402

B = int(os.environ.get("B", "16"))
STEPS = int(os.environ.get("STEPS", "3"))
torch.manual_seed(0)
net = Unet_Libero()
with torch.no_grad():
    for p in net.parameters():
        if p.dim() > 1:
            p.add_(0.02 * torch.randn_like(p))
d = GoalGaussianDiffusion(net, image_size=(128, 128), channels=21, timesteps=100, sampling_timesteps=100,
                          loss_type="l2", objective="pred_v", beta_schedule="cosine", min_snr_loss_weight=True,
                          guidance_weight=0).cuda()
d.sampling_timesteps, d.is_ddim_sampling = STEPS, True
out = d.sample(torch.rand(B, 3, 128, 128, device="cuda"), torch.randn(B, 12, 512, device="cuda"), batch_size=B)
torch.cuda.synchronize()
print("done", float(out.mean()))
