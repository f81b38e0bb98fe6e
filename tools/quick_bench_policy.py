"""Developer timing probe for the policy path (not the contract bench).

usage: python tools/quick_bench_policy.py [B] [--layers]
Times ConditionalUnet1D forward, backward, and the fused optimiser tail at the Libero config.
"""
import os
import sys
import time

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests.golden.configs import POLICY_LIBERO  # noqa: E402
from v2a_b200 import policy_unet1d as PU  # noqa: E402
from v2a_b200.train_step import PolicyTrainStep  # noqa: E402


def ev():
    return torch.cuda.Event(enable_timing=True)


def timed(fn, n=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = ev(), ev()
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 and sys.argv[1].isdigit() else 256
    T = 16
    torch.manual_seed(0)
    net = PU.ConditionalUnet1D(**POLICY_LIBERO).cuda()
    x = torch.randn(B, T, 7, device="cuda")
    t = torch.randint(0, 100, (B,), device="cuda")
    gc = torch.randn(B, 128, device="cuda")
    noise = torch.randn(B, T, 7, device="cuda")
    t0 = time.time()
    step = PolicyTrainStep(net)

    def loss_fn():
        return F.mse_loss(net(x, t, global_cond=gc), noise)
    step.step(loss_fn)
    torch.cuda.synchronize()
    print(f"first step (plan build + pack) {time.time() - t0:.2f}s")
    eng = PU.last_engine(net)
    flops_f = sum(g.flops for g in eng.igemms) + sum(g.flops for g in eng.wgrads)
    ms_step = timed(lambda: step.step(loss_fn))
    print(f"B={B} full step {ms_step:.3f} ms -> {B / ms_step * 1e3:.0f} samples/s; igemm flops fwd+bwd {flops_f / 1e9:.1f} GF "
          f"-> {flops_f / ms_step / 1e9:.1f} TFLOP/s; launches fwd {len(eng.fwd)} bwd {len(eng.bwd)}")
    xs, ts, gs = x.float(), t, gc.float()
    eng.refresh_weights()
    ms_f = timed(lambda: eng.forward(xs, ts, gs))
    go = torch.randn(B, T, 7, device="cuda")
    ms_b = timed(lambda: eng.backward(go, clone_param_grads=False))
    ms_t = timed(step.optimizer_tail)
    eng._wkey = None
    ms_r = timed(lambda: (setattr(eng, "_wkey", None), eng.refresh_weights()))
    print(f"  forward {ms_f:.3f} ms  backward {ms_b:.3f} ms  optimiser tail {ms_t:.3f} ms  weight repack {ms_r:.3f} ms")
    with torch.no_grad():
        ms_inf = timed(lambda: net(x, t, global_cond=gc))
    print(f"  no-grad forward through the module {ms_inf:.3f} ms")
    if "--layers" in sys.argv:
        for name, steps in (("fwd", eng.fwd), ("bwd", eng.bwd)):
            rows = []
            for i, s in enumerate(steps):
                rows.append((timed(s, n=5, warm=1), i, steps.tags[i]))
            tot = sum(r[0] for r in rows)
            print(f"{name}: sum of per-launch times {tot:.3f} ms over {len(rows)} launches")
            by = {}
            for m, i, tag in rows:
                k = tag.split()[0]
                by[k] = (by.get(k, (0, 0))[0] + m, by.get(k, (0, 0))[1] + 1)
            print("  by kind: " + ", ".join(f"{k} {v[0]:.3f} ms/{v[1]}" for k, v in sorted(by.items(), key=lambda kv: -kv[1][0])))
            for m, i, tag in sorted(rows, key=lambda r: -r[0])[:30]:
                print(f"  #{i:3d} {m:7.4f} ms {tag}")


if __name__ == "__main__":
    main()
