import os, sys, torch
import torch.nn.functional as F
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, root)
from oracle import encoder_oracle as EO
from tests.test_encoder_gpu import _seeded_core
from v2a_b200 import obs_encoder as OE, ops
torch.backends.cudnn.allow_tf32 = False
def rel(a, b): return ((a.double() - b.double()).norm() / b.double().norm()).item()
B = 3
core = _seeded_core(); core.train()
torch.manual_seed(3)
x = torch.rand(B, 3, 128, 128, device="cuda") * 2 - 1
wout = torch.randn(B, 64, device="cuda")
got = core(x); (got * wout).sum().backward()
eng = OE.last_engine(core)
def nchw(t2d, a): return t2d.view(a.N, a.H, a.W, a.C).permute(0, 3, 1, 2)
masks = {}
for k, (a1, out) in enumerate(zip(eng.inner_acts, eng.acts[1:])):
    masks[2 * k + 1] = nchw(a1.float() > 0, out).double(); masks[2 * k + 2] = nchw(out.f32 > 0, out).double()
sp = eng.stem_probe; gn0 = core.backbone.nets[1]
raw0 = sp["raw0"].view(B, sp["H"], sp["W"], -1).permute(0, 3, 1, 2).double()
key = "k."
sd = {key + n: p.detach().double().clone().requires_grad_(True) for n, p in core.named_parameters()}
sd.update({key + n: b.detach().double() for n, b in core.named_buffers()})
relu = lambda t, i: t * masks[i] if i in masks else torch.relu(t)
p = key + "backbone.nets."
y0 = F.conv2d(x.double(), sd[p + "0.weight"], stride=2, padding=3); y0.retain_grad()
print("stem conv output: ours vs f64", rel(raw0, y0))
hp = EO._gn(sd, p + "1.", y0); hp.retain_grad()
h = torch.relu(hp); h.retain_grad()
h2 = F.max_pool2d(h, 3, 2, 1); h2.retain_grad()
z = h2; k = 0
for li in range(4, 8):
    for bi in range(2):
        z = EO.basic_block(sd, f"{p}{li}.{bi}.", z, 2 if (li > 4 and bi == 0) else 1, relu, 2 * k + 1); k += 1
kp = EO.spatial_softmax(sd, key + "pool.", z)
ref = F.linear(kp.flatten(1), sd[key + "nets.3.weight"], sd[key + "nets.3.bias"])
(ref * wout.double()).sum().backward()
print("dP0 (grad wrt pooled):", rel(eng.acts[0].grad, h2.grad.permute(0, 2, 3, 1).reshape(-1, 64)))
# our draw0 is not stored; recompute the weight gradient two ways from OUR planes
pr = None
d_y0 = y0.grad.permute(0, 2, 3, 1).reshape(-1, 64)          # truth gradient wrt stem conv output
w_grad_truth = sd[p + "0.weight"].grad
mine = core.backbone.nets[0].weight.grad
print("conv1.weight grad ours vs truth", rel(mine, w_grad_truth))
# wgrad kernel in isolation on the TRUTH d_y0 and our im2col planes
col = F.unfold(x.double(), 7, padding=3, stride=2).transpose(1, 2).reshape(-1, 147)   # [B*4096, 147] k = c*49+ky*7+kx
print("truth check: unfold^T d_y0", rel((d_y0.t() @ col).reshape(64, 3, 7, 7), w_grad_truth))
colp = ops.split_hl(F.pad(col.float(), (0, 45)).contiguous())
dyp = ops.split_hl(d_y0.float().contiguous())
sc = torch.zeros(192, 64, device="cuda")
ops.Wgrad(srcs=[(colp, 192, (4096, B, 1, 1))], units=[(0, (0, 0, 0, 0), c) for c in range(3)], dy=dyp, dy_channels=64,
          dy_dims=(4096, B, 1, 1), cout=64, out=sc).run()
dw = torch.zeros(64, 147, device="cuda"); ops.wgrad_scatter(sc, 64, 147, 1, dw)
print("wgrad kernel on truth d_y0:", rel(dw.view(64, 3, 7, 7), w_grad_truth), "k_splits")
print("fp32 matmul on truth d_y0:", rel((d_y0.float().t() @ col.float()).reshape(64, 3, 7, 7), w_grad_truth))

g0 = sp["g0"]; d0 = sp["d0"].float()
tg = hp.grad.permute(0, 2, 3, 1).reshape(-1, 64)
print("g0 (grad wrt GN0 output) ours vs truth", rel(g0, tg), "nonzero frac", (tg != 0).double().mean().item())
dd = (g0.double() - tg).abs().amax(1).view(B, 64, 64)
print("  bad pixels", int((dd > 1e-3 * tg.abs().max()).sum()), (dd > 1e-3 * tg.abs().max()).nonzero()[:6].tolist())
print("d0 (grad wrt conv1 output) ours vs truth", rel(d0, d_y0))
dd = (d0.double() - d_y0).abs().amax(1).view(B, 64, 64)
print("  bad pixels", int((dd > 1e-3 * d_y0.abs().max()).sum()), (dd > 1e-3 * d_y0.abs().max()).nonzero()[:6].tolist())
