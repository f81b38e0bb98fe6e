"""Developer timing probe for the observation-encoder engine (not the contract bench).
usage: python tools/quick_bench_encoder.py [B] [--layers]"""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from v2a_b200 import diffusion_policy as DP, obs_encoder as OE  # noqa: E402


def timed(fn, n=5, warm=2):
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
    B = int(sys.argv[1]) if len(sys.argv) > 1 and sys.argv[1].isdigit() else 256
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.manual_seed(0)
    pol = DP.build_libero_policy().cuda()
    pol.train()
    core = pol.obs_encoder.key_model_map["img_obs_1"]
    x = torch.rand(B, 3, 128, 128, device="cuda") * 2 - 1
    w = torch.randn(B, 64, device="cuda")
    t0 = time.time()
    out = core(x)
    (out * w).sum().backward()
    torch.cuda.synchronize()
    print(f"first fwd+bwd (plan build) {time.time() - t0:.2f}s  mem {torch.cuda.memory_allocated() / 1e9:.2f} GB")
    eng = OE.last_engine(core)
    fl_f = sum(g.flops for g in eng.igemms[:sum(1 for t in eng.fwd.tags if t.startswith('igemm'))])
    fl_all = sum(g.flops for g in eng.igemms) + sum(g.flops for g in eng.wgrads)
    ms_f = timed(lambda: eng.forward(x))
    ms_b = timed(lambda: eng.backward(w, clone_param_grads=False))
    print(f"B={B} engine forward {ms_f:.3f} ms  backward {ms_b:.3f} ms  (fwd {fl_f / 1e9:.0f} GF, all {fl_all / 1e9:.0f} GF -> "
          f"{fl_all / (ms_f + ms_b) / 1e9:.1f} TFLOP/s algorithmic)  launches fwd {len(eng.fwd)} bwd {len(eng.bwd)}")
    def torch_step():
        core.zero_grad(set_to_none=True)
        o = core.nets(x)          # the stock-op twin on the parameter-holder modules
        (o * w).sum().backward()
    ms_t = timed(torch_step, n=3, warm=1)
    print(f"torch/cuDNN fp32 (TF32 off) fwd+bwd {ms_t:.3f} ms -> engine speed-up {ms_t / (ms_f + ms_b):.2f}x")
    if "--layers" in sys.argv:
        for name, steps in (("fwd", eng.fwd), ("bwd", eng.bwd)):
            rows = [(timed(s, n=3, warm=1), i, steps.tags[i]) for i, s in enumerate(steps)]
            tot = sum(r[0] for r in rows)
            print(f"{name}: sum of per-launch times {tot:.3f} ms over {len(rows)} steps")
            by = {}
            for m, i, tag in rows:
                k = tag.split()[0] + (" " + tag.split()[1] if tag.startswith("igemm") and len(tag.split()) > 1 else "")
                by[k] = (by.get(k, (0, 0))[0] + m, by.get(k, (0, 0))[1] + 1)
            print("  by kind: " + ", ".join(f"{k} {v[0]:.3f} ms/{v[1]}" for k, v in sorted(by.items(), key=lambda kv: -kv[1][0])))
            for m, i, tag in sorted(rows, key=lambda r: -r[0])[:22]:
                print(f"  #{i:3d} {m:7.4f} ms {tag}")


if __name__ == "__main__":
    main()
