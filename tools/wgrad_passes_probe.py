"""Developer probe: how far do the policy's parameter gradients move when the weight-gradient GEMMs run ONE bf16 product
instead of the three-pass split (V2A_WGRAD_PASSES=1 / auto), at the benchmarked batch (B = 256)?  Each setting runs in
its own process (the switch is read when the engines are planned); gradients are compared per parameter tensor against
the three-pass run.  usage: python tools/wgrad_passes_probe.py [setting ...]   (setting: 1 | auto | auto:MINK)"""
import os
import subprocess
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import os, sys, torch
sys.path.insert(0, os.getcwd())
from v2a_b200.diffusion_policy import build_libero_policy
torch.manual_seed(77)
pol = build_libero_policy().to("cuda"); pol.train()
g = torch.Generator().manual_seed(2000)
B = 256
batch = {"obs": {"img_obs_1": torch.rand(B, 1, 3, 128, 128, generator=g).cuda(), "img_goal_1": torch.rand(B, 1, 3, 128, 128, generator=g).cuda()},
         "action": (torch.rand(B, 16, 7, generator=g) * 2 - 1).cuda()}
torch.manual_seed(5)
loss = pol.compute_loss(batch)
loss.backward()
torch.cuda.synchronize()
torch.save({"loss": float(loss), "grads": {n: p.grad.detach().cpu() for n, p in pol.named_parameters() if p.grad is not None and p.numel()}}, sys.argv[1])
'''


def run(setting, path):
    env = dict(os.environ)
    if setting.startswith("auto:"):
        env["V2A_WGRAD_PASSES"], env["V2A_WGRAD_MINK"] = "auto", setting.split(":")[1]
    else:
        env["V2A_WGRAD_PASSES"] = setting
    out = subprocess.run([sys.executable, "-c", CHILD, path], env=env, cwd=ROOT, capture_output=True, text=True)
    if out.returncode:
        print(setting, "failed", out.stderr[-400:])
        return None
    return torch.load(path)


def main():
    settings = sys.argv[1:] or ["1", "auto", "auto:1024"]
    ref = run("3", "/tmp/wg_ref.pt")
    again = run("3", "/tmp/wg_ref2.pt")       # run-to-run noise of the three-pass path (atomics reorder)
    for name, other in [("3 (repeat)", again)] + [(s, run(s, f"/tmp/wg_{i}.pt")) for i, s in enumerate(settings)]:
        if other is None:
            continue
        rows = []
        for n, gr in ref["grads"].items():
            go = other["grads"][n]
            den = gr.double().norm().item()
            rows.append(((go.double() - gr.double()).norm().item() / den if den > 0 else 0.0, n, tuple(gr.shape)))
        rows.sort(reverse=True)
        worst = ", ".join(f"{e:.2e} {n} {s}" for e, n, s in rows[:4])
        med = rows[len(rows) // 2][0]
        print(f"V2A_WGRAD_PASSES={name}: loss {other['loss']:.7f} (ref {ref['loss']:.7f}); per-tensor rel-L2 vs three passes: "
              f"max {rows[0][0]:.2e}, median {med:.2e}, tensors > 1e-3: {sum(r[0] > 1e-3 for r in rows)} of {len(rows)}\n    worst: {worst}")


if __name__ == "__main__":
    main()
