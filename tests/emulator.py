"""CPU/GPU-agnostic emulator of the implicit-GEMM kernel's *semantics*.

Test infrastructure only: it evaluates a `ConvProgram` (tap program + packed
K-major weights) with plain tensor indexing in the dtype of its inputs, so the
host-side logic (tap offsets, phase split, weight packing, padding) can be
pinned against torch convolutions on CPU, and the CUDA kernel can be checked
against it in float64 on the GPU.
"""
import torch


def _shifted(x, off, out_dims):
    """x [X3, X2, X1, X0, C] -> window [D3, D2, D1, D0, C] at coordinate offset `off`, zero filled."""
    X = [x.shape[3], x.shape[2], x.shape[1], x.shape[0]]  # X0..X3
    idx, msk = [], []
    for d in range(4):
        i = torch.arange(out_dims[d], device=x.device) + off[d]
        ok = (i >= 0) & (i < X[d])
        idx.append(i.clamp(0, X[d] - 1))
        msk.append(ok)
    i3, i2, i1, i0 = idx[3], idx[2], idx[1], idx[0]
    g = x[i3[:, None, None, None], i2[None, :, None, None], i1[None, None, :, None], i0[None, None, None, :]]
    m = (msk[3][:, None, None, None] & msk[2][None, :, None, None] & msk[1][None, None, :, None]
         & msk[0][None, None, None, :])
    return g * m[..., None].to(g.dtype)


def emulate(program, srcs, w, cout):
    """srcs: list of [X3, X2, X1, X0, C] tensors; w: [>=cout, ktot]; returns [rows, cout]."""
    D = program.out_dims
    out = None
    k0 = 0
    for (s, off, nch) in program.taps:
        x = srcs[s]
        C = x.shape[-1]
        xs = _shifted(x, off, D).reshape(-1, C)
        wk = w[:cout, k0:k0 + C].to(xs.dtype)
        y = xs @ wk.t()
        out = y if out is None else out + y
        k0 += nch * 64
    return out


def as5d(x2d, channels, dims):
    """[rows, C] channels-last -> [X3, X2, X1, X0, C]."""
    return x2d.reshape(dims[3], dims[2], dims[1], dims[0], channels)
