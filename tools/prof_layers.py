"""ncu target for single layers at the config-2 geometry (B=16, 128x128x7, C=128 full-resolution level):
  ncu --set full --import-source on --clock-control none --profile-from-start off -o gpurun_out/prof python tools/prof_layers.py
Profiles (between cudaProfilerStart/Stop): the GroupNorm+SiLU prep of in.1, its 3x3 spatial conv (K=1152, N=128),
its temporal conv (K=384, N=128, + emb add + GroupNorm sums) and the widest N=256 spatial conv."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from v2a_b200.unet import Unet_Libero  # noqa: E402

B = int(os.environ.get("B", "16"))
torch.manual_seed(0)
net = Unet_Libero().cuda()
with torch.no_grad():
    for p in net.parameters():
        if p.dim() > 1:
            p.add_(0.02 * torch.randn_like(p))
x = torch.randn(B, 24, 128, 128, device="cuda")
t = torch.full((B,), 50, device="cuda")
te = torch.randn(B, 12, 512, device="cuda")
net(x, t, te)
net(x, t, te)
torch.cuda.synchronize()
if os.environ.get("WARM", "1") == "1":      # bring clocks / power state to steady state before timing
    import time
    t0 = time.time()
    while time.time() - t0 < 4.0:
        net(x, t, te)
    torch.cuda.synchronize()
eng = net.unet.engine(B, 7, 128, 128, "cuda")
preps = [i for i, tag in enumerate(eng.tags) if tag == "prep_gn"]
picks = [("prep_gn C128", eng.steps[preps[0]]), ("igemm spatial K1152 N128", eng.igemms[2].run),
         ("igemm temporal K384 N128", eng.igemms[3].run), ("igemm spatial K2304 N256", eng.igemms[134].run)]
extra = os.environ.get("IGEMMS", "")
for s in [int(v) for v in extra.split(",") if v]:
    picks.append((f"igemm #{s}", eng.igemms[s].run))
for name, fn in picks:          # warm
    fn()
torch.cuda.synchronize()
torch.cuda.profiler.start()
for name, fn in picks:
    fn()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
for name, fn in picks:
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        fn()
    e1.record()
    torch.cuda.synchronize()
    print(f"{name}: {e0.elapsed_time(e1) / 3:.3f} ms")
