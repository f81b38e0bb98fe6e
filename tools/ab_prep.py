"""A/B probe: summed time of the prep launches of one B=16 UNet forward under V2A_PREP_ITERS settings."""
import os, subprocess, sys
CODE = r'''
import os, sys, torch
sys.path.insert(0, os.getcwd())
from v2a_b200.unet import Unet_Libero
torch.manual_seed(0)
net = Unet_Libero().cuda()
x = torch.randn(16, 24, 128, 128, device="cuda"); t = torch.full((16,), 50, device="cuda"); te = torch.randn(16, 12, 512, device="cuda")
for _ in range(2): net(x, t, te)
torch.cuda.synchronize()
eng = net.unet.engine(16, 7, 128, 128, "cuda")
tot = 0.0
for st, tag in zip(eng.steps, eng.tags):
    if not tag.startswith("prep"): continue
    st(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3): st()
    e1.record(); torch.cuda.synchronize()
    tot += e0.elapsed_time(e1) / 3
print(f"{tot:.3f}")
'''
for it in (sys.argv[1:] or ["2", "4", "8", "16", "32"]):
    env = dict(os.environ, V2A_PREP_ITERS=it)
    out = subprocess.run([sys.executable, "-c", CODE], env=env, capture_output=True, text=True)
    print("iters", it, out.stdout.strip().splitlines()[-1] if out.stdout.strip() else out.stderr[-200:])
