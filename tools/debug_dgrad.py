import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from v2a_b200 import convs, ops, obs_encoder as OE
torch.backends.cudnn.allow_tf32 = False
def rel(a, b): return ((a.double() - b.double()).norm() / b.double().norm()).item()
for (N, H, W, Ci, Co) in [(3, 32, 32, 64, 64), (3, 16, 16, 128, 128), (8, 32, 32, 64, 64)]:
    torch.manual_seed(0)
    w = torch.randn(Co, Ci, 3, 3, device="cuda") / 10
    dy = torch.randn(N, H, W, Co, device="cuda")
    res = torch.randn(N * H * W, Ci, device="cuda")
    wd = ops.split_hl_torch(OE.dgrad3x3_weight(w))
    prog = convs.spatial3x3(Co, N, H, W)
    hl = ops.split_hl(dy.reshape(-1, Co).contiguous())
    want = torch.nn.grad.conv2d_input((N, Ci, H, W), w, dy.permute(0, 3, 1, 2), padding=1).permute(0, 2, 3, 1).reshape(-1, Ci)
    for use_res in (False, True):
        out = torch.zeros(N * H * W, Ci, device="cuda")
        g = ops.Igemm(srcs=[(hl, Co, prog.src_dims[0])], taps=prog.taps, w=wd, out_dims=prog.out_dims, cout=Ci, out_f32=out,
                      residual=res if use_res else None)
        g.run(); torch.cuda.synchronize()
        e = rel(out, want + (res if use_res else 0))
        # per-row error map
        d = (out - want - (res if use_res else 0)).abs().amax(1).view(N, H, W)
        bad = (d > 1e-3).nonzero()
        print(f"N{N} {H}x{W} C{Ci}->{Co} residual={use_res} ks={g.k_splits} rel={e:.2e} bad rows={len(bad)} first={bad[:6].tolist()}")
