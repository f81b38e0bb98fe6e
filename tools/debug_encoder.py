import copy, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from tests.test_encoder_gpu import _seeded_core
from v2a_b200 import obs_encoder as OE
torch.backends.cudnn.allow_tf32 = False
def rel(a, b): return ((a.double() - b.double()).norm() / b.double().norm()).item()
B = 3
core = _seeded_core(); core.train()
torch.manual_seed(3)
x = torch.rand(B, 3, 128, 128, device="cuda") * 2 - 1
wout = torch.randn(B, 64, device="cuda")
c64 = copy.deepcopy(core).double()
outs = []
mods = [c64.backbone.nets[3]] + [b for i in range(4, 8) for b in c64.backbone.nets[i]]
for m in mods:
    def hook(mod, inp, out, store=outs):
        out.retain_grad(); store.append(out)
    m.register_forward_hook(hook)
ref = c64.nets(x.double()); (ref * wout.double()).sum().backward()
got = core(x); (got * wout).sum().backward()
eng = OE.last_engine(core)
for i, (a, t) in enumerate(zip(eng.acts, outs)):
    tf = t.detach().permute(0, 2, 3, 1).reshape(-1, t.shape[1])
    tg = t.grad.permute(0, 2, 3, 1).reshape(-1, t.shape[1])
    ef = rel(a.f32, tf) if a.f32 is not None else float("nan")
    eg = rel(a.grad, tg) if a.grad is not None else float("nan")
    d = (a.grad.double() - tg).abs().amax(1).view(a.N, a.H, a.W) if a.grad is not None else None
    nb = int((d > 1e-3 * tg.abs().max()).sum()) if d is not None else -1
    print(f"act {i}: {a.H}x{a.W}x{a.C} fwd rel {ef:.2e} grad rel {eg:.2e} bad pixels {nb}", (d > 1e-3 * tg.abs().max()).nonzero()[:5].tolist() if d is not None else "")
# block 0 internals: grads wrt conv2 output (draw2) and conv1 output (draw1)
blk = c64.backbone.nets[4][0]
# rerun torch with hooks on conv outputs
c64.zero_grad()
store = {}
def mk(name):
    def hook(mod, inp, out):
        out.retain_grad(); store[name] = out
    return hook
h1 = blk.conv1.register_forward_hook(mk("c1")); h2 = blk.conv2.register_forward_hook(mk("c2"))
ref = c64.nets(x.double()); (ref * wout.double()).sum().backward()
pr = eng.probes[-1]   # plans run in reverse: the last probe is block 0
for name, key in (("c2", "d2"), ("c1", "d1")):
    t = store[name].grad.permute(0, 2, 3, 1).reshape(-1, 64)
    mine = pr[key].float()
    d = (mine.double() - t).abs().amax(1).view(B, 32, 32)
    print(name, "rel", rel(mine, t), "bad", (d > 1e-3 * t.abs().max()).nonzero()[:8].tolist())
for name, key in (("c2", "raw2"), ("c1", "raw1")):
    t = store[name].detach().permute(0, 2, 3, 1).reshape(-1, 64)
    d = (pr[key].double() - t).abs().amax(1).view(B, 32, 32)
    print("fwd", name, "rel", rel(pr[key], t), "bad", (d > 1e-3 * t.abs().max()).nonzero()[:8].tolist())
c64.zero_grad(); store.clear()
h3 = blk.bn2.register_forward_hook(mk("b2"))
ref = c64.nets(x.double()); (ref * wout.double()).sum().backward()
tg = store["b2"].grad.permute(0, 2, 3, 1).reshape(-1, 64)
mine = pr["g"]
d = (mine.double() - tg).abs().amax(1).view(B, 32, 32)
print("g rel", rel(mine, tg), "bad", (d > 1e-3 * tg.abs().max()).nonzero()[:8].tolist())
row = (1 * 32 + 24) * 32 + 1
print("g mine ", mine[row, :8].tolist()); print("g truth", tg[row, :8].tolist())
t2 = store["c2"].grad.permute(0, 2, 3, 1).reshape(-1, 64)
print("d2 mine ", pr["d2"].float()[row, :8].tolist()); print("d2 truth", t2[row, :8].tolist())
print("d2 ratio", (pr["d2"].float()[row, :8].double() / t2[row, :8]).tolist())
dO = eng.acts[1].grad
print("dO mine", dO[row, :8].tolist()); print("dO truth", outs[1].grad.permute(0, 2, 3, 1).reshape(-1, 64)[row, :8].tolist())
print("out mine", eng.acts[1].f32[row, :8].tolist()); print("out truth", outs[1].detach().permute(0, 2, 3, 1).reshape(-1, 64)[row, :8].tolist())
print("---- worst channel at the bad pixel")
diff = (pr["d2"].float()[row].double() - t2[row]).abs()
c = int(diff.argmax())
print("channel", c, "d2 mine", pr["d2"].float()[row, c].item(), "truth", t2[row, c].item())
print("dO mine", dO[row, c].item(), "truth", outs[1].grad.permute(0, 2, 3, 1).reshape(-1, 64)[row, c].item())
print("out mine", eng.acts[1].f32[row, c].item(), "truth", outs[1].detach().permute(0, 2, 3, 1).reshape(-1, 64)[row, c].item())
print("raw2 mine", pr["raw2"][row, c].item(), "truth", store["c2"].detach().permute(0, 2, 3, 1).reshape(-1, 64)[row, c].item())
print("idn (P0) mine", eng.acts[0].f32[row, c].item(), "truth", outs[0].detach().permute(0, 2, 3, 1).reshape(-1, 64)[row, c].item())
# global: how many elements of d2 differ by more than 1e-3 * max
dd = (pr["d2"].float().double() - t2).abs()
print("elements off by > 1e-3 max:", int((dd > 1e-3 * t2.abs().max()).sum()), "of", dd.numel(), "max", dd.max().item(), "t2 max", t2.abs().max().item())
idx = (dd > 1e-3 * t2.abs().max()).nonzero()[:10]
for r_, c_ in idx.tolist():
    print(" row", r_, "ch", c_, "mine", pr["d2"].float()[r_, c_].item(), "truth", t2[r_, c_].item(), "out mine", eng.acts[1].f32[r_, c_].item(),
          "out truth", outs[1].detach().permute(0, 2, 3, 1).reshape(-1, 64)[r_, c_].item(), "dO", dO[r_, c_].item())
