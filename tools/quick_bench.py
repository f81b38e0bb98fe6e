"""Developer timing probe (not the contract bench): UNet forward time + per-launch breakdown.

usage: python tools/quick_bench.py [B] [--layers]
"""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from v2a_b200 import ops  # noqa: E402
from v2a_b200.unet import Unet_Libero  # noqa: E402


def timed(fn, n=3, warm=1):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 and sys.argv[1].isdigit() else 2
    Fr, H, W = 7, 128, 128
    torch.manual_seed(0)
    net = Unet_Libero().cuda()
    with torch.no_grad():
        for p in net.parameters():
            if p.dim() > 1:
                p.add_(0.02 * torch.randn_like(p))
    x = torch.randn(B, 3 * Fr + 3, H, W, device="cuda")
    t = torch.full((B,), 50, device="cuda")
    te = torch.randn(B, 12, 512, device="cuda")
    t0 = time.time()
    net(x, t, te)
    torch.cuda.synchronize()
    print(f"first forward (plan build + pack) {time.time() - t0:.2f}s")
    ms = timed(lambda: net(x, t, te))
    eng = net.unet.engine(B, Fr, H, W, "cuda")
    print(f"B={B} forward {ms:.2f} ms  igemm-flops {eng.flops / 1e12:.3f} TF -> {eng.flops / ms / 1e9:.1f} TFLOP/s "
          f"(algorithmic; x{eng.passes} MMA passes)  launches/forward {len(eng.steps)}  arena {eng.pool.total / 1e9:.2f} GB")
    print(f"torch mem allocated {torch.cuda.memory_allocated() / 1e9:.2f} GB")
    if "--layers" in sys.argv:
        rows = []
        tot = 0.0
        for i, g in enumerate(eng.igemms):
            m = timed(g.run, n=3, warm=1)
            tot += m
            d = g.desc
            rows.append((m, i, g.rows, g.cout, g.ktot, d.block_n, g.flops / m / 1e9))
        print(f"sum of igemm launches {tot:.2f} ms ({len(eng.igemms)} launches)")
        ideal_tf = 1390.7 / eng.passes   # measured sustained bf16 peak / MMA passes
        lost = lambda m, tf: m * (1.0 - min(tf / ideal_tf, 1.0))
        print(f"  time above the tensor roofline ({ideal_tf:.0f} TFLOP/s algorithmic): "
              f"{sum(lost(m, tf) for m, _, _, _, _, _, tf in rows):.2f} ms; by (cout, kind):")
        cls = {}
        for m, i, r, c, k, bn, tf in rows:
            key = (c, "temporal/1x1" if k <= 6 * 64 * max(1, c // 64) and k < 9 * c else "spatial")
            a = cls.setdefault(key, [0.0, 0.0, 0])
            a[0] += m; a[1] += lost(m, tf); a[2] += 1
        for key, a in sorted(cls.items(), key=lambda kv: -kv[1][1]):
            print(f"    cout {key[0]:5d} {key[1]:12s} n={a[2]:3d}  {a[0]:7.2f} ms  lost {a[1]:6.2f} ms")
        for m, i, r, c, k, bn, tf in sorted(rows, reverse=True):
            print(f"  #{i:3d} rows {r:8d} cout {c:5d} K {k:6d} bn {bn:3d}  {m:7.3f} ms  {tf:7.1f} TFLOP/s  ks {eng.igemms[i].k_splits}")
        print(f"non-igemm share ~ {ms - tot:.2f} ms")
        # every planned step, eagerly, by kind (static I/O bound by the forward above)
        eng.stats_arena.zero_()
        by, rows2 = {}, []
        for i, (st, tag) in enumerate(zip(eng.steps, eng.tags)):
            m = timed(st, n=2, warm=1)
            by[tag.split()[0]] = (by.get(tag.split()[0], (0, 0))[0] + m, by.get(tag.split()[0], (0, 0))[1] + 1)
            if not tag.startswith("igemm"):
                rows2.append((m, i, tag))
        print("by kind: " + ", ".join(f"{k} {v[0]:.2f} ms/{v[1]}" for k, v in sorted(by.items(), key=lambda kv: -kv[1][0])))
        for m, i, tag in sorted(rows2, reverse=True)[:25]:
            print(f"  step {i:3d} {m:7.3f} ms {tag}")


if __name__ == "__main__":
    main()
