"""Timing probe of single igemm shapes under V2A_IGEMM_DEBUG variants (developer tool)."""
import os, subprocess, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

def run():
    from v2a_b200 import convs, ops
    B, Fr, HW, c = 16, 7, 16384, 128
    rows = B * Fr * HW
    y = ops.HL.empty(rows, c, "cuda"); y.hi.normal_(); y.lo.zero_()
    wt = torch.randn(c, c, 3, device="cuda") / 20
    out = torch.empty(rows, c, device="cuda")
    res = torch.randn(rows, c, device="cuda")
    emb = torch.randn(B, c, device="cuda")
    bias = torch.randn(c, device="cuda")
    stats = torch.zeros(8, B * Fr, c, 2, dtype=torch.float64, device="cuda")
    prog = convs.temporal3(c, B, Fr, HW)
    w = ops.split_hl_torch(convs.temporal3_weight(wt))
    def t(g, n=5):
        g.run(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n): g.run()
        e1.record(); torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n
    base = dict(srcs=[(y, c, prog.src_dims[0])], taps=prog.taps, w=w, out_dims=prog.out_dims, cout=c)
    r = {}
    r["plain f32 out"] = t(ops.Igemm(out_f32=out, **base))
    r["+bias+rowvec"] = t(ops.Igemm(out_f32=out, bias=bias, rowvec=emb, rowvec_mul=(0, 0, 1, 0), **base))
    r["+stats"] = t(ops.Igemm(out_f32=out, bias=bias, rowvec=emb, rowvec_mul=(0, 0, 1, 0), stats=stats, stats_mul=(0, 1, Fr, 0), **base))
    r["+stats+residual"] = t(ops.Igemm(out_f32=out, bias=bias, residual=res, stats=stats, stats_mul=(0, 1, Fr, 0), **base))
    ohl = ops.HL.empty(rows, c, "cuda")
    r["hl out"] = t(ops.Igemm(out_hl=ohl, bias=bias, **base))
    progs = convs.spatial3x3(c, B * Fr, 128, 128)
    ws = ops.split_hl_torch(convs.spatial3x3_weight(torch.randn(c, c, 3, 3, device="cuda") / 30))
    r["spatial K=1152 hl out"] = t(ops.Igemm(srcs=[(y, c, progs.src_dims[0])], taps=progs.taps, w=ws, out_dims=progs.out_dims, cout=c, out_hl=ohl, bias=bias))
    print(os.environ.get("V2A_IGEMM_DEBUG", "0"), {k: round(v, 3) for k, v in r.items()})

if __name__ == "__main__":
    if len(sys.argv) > 1:
        run()
    else:
        for dbg in ("0", "1", "2", "3", "7"):
            subprocess.run([sys.executable, __file__, "x"], env=dict(os.environ, V2A_IGEMM_DEBUG=dbg))
