"""A/B probe: UNet forward time at B=16 under different V2A_* settings, interleaved in ONE process run per setting
(each setting in a fresh subprocess, repeated round-robin so box / clock drift hits all settings alike)."""
import os, subprocess, sys
SETTINGS = [s for s in sys.argv[1:]] or ["V2A_CTA2=1", "V2A_CTA2=3", "V2A_CTA2=0"]
CODE = r'''
import os, sys, torch
sys.path.insert(0, os.getcwd())
from v2a_b200.unet import Unet_Libero
torch.manual_seed(0)
net = Unet_Libero().cuda()
x = torch.randn(16, 24, 128, 128, device="cuda"); t = torch.full((16,), 50, device="cuda"); te = torch.randn(16, 12, 512, device="cuda")
for _ in range(3): net(x, t, te)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(8): net(x, t, te)
e1.record(); torch.cuda.synchronize()
print(f"{e0.elapsed_time(e1) / 8:.2f}")
'''
res = {s: [] for s in SETTINGS}
for rep in range(3):
    for s in SETTINGS:
        env = dict(os.environ)
        for kv in s.split(","):
            k, v = kv.split("=")
            env[k] = v
        out = subprocess.run([sys.executable, "-c", CODE], env=env, capture_output=True, text=True)
        try:
            res[s].append(float(out.stdout.strip().splitlines()[-1]))
        except Exception:
            print(s, "failed", out.stderr[-300:])
for s, v in res.items():
    print(s, " ".join(f"{x:.2f}" for x in v), "ms  min", min(v) if v else None)
