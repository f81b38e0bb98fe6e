"""Experiment: can an HBM-bound prep launch co-run with a tensor-bound igemm launch (two streams)?"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from v2a_b200.unet import Unet_Libero
B = 16
torch.manual_seed(0)
net = Unet_Libero().cuda()
x = torch.randn(B, 24, 128, 128, device="cuda"); t = torch.full((B,), 50, device="cuda"); te = torch.randn(B, 12, 512, device="cuda")
net(x, t, te); net(x, t, te); torch.cuda.synchronize()
eng = net.unet.engine(B, 7, 128, 128, "cuda")
preps = [i for i, tag in enumerate(eng.tags) if tag == "prep_gn"]
prep = eng.steps[preps[0]]
for name, ig in (("spatial K2304 N256", eng.igemms[134].run), ("spatial K1152 N128", eng.igemms[2].run), ("temporal K384 N128", eng.igemms[3].run)):
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    def timed(fn, n=5):
        fn(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n): fn()
        e1.record(); torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n
    NP = 6
    def seq():
        ig()
        for _ in range(NP): prep()
    def par():
        cur = torch.cuda.current_stream()
        s1.wait_stream(cur); s2.wait_stream(cur)
        with torch.cuda.stream(s1): ig()
        with torch.cuda.stream(s2):
            for _ in range(NP): prep()
        cur.wait_stream(s1); cur.wait_stream(s2)
    t_ig = timed(ig); t_p = timed(prep)
    print(f"{name}: igemm {t_ig:.3f} ms, prep {t_p:.3f} ms x{NP}; sequential {timed(seq):.3f} ms, two streams {timed(par):.3f} ms")
