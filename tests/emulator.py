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


def emulate_fused3d(fp, srcs, w_spatial, b_spatial, w_temporal, cout):
    """Semantics of the planned one-kernel Conv3d (docs/FUSED_CONV3D_PLAN.md): stage 1 = `emulate` of the spatial
    program + spatial bias, evaluated tile by tile in the kernel's on-chip layout — rows = 16 pixels x 8 frame slots,
    slots >= F zero — then stage 2 = whole-frame shifts of that tile (16 zero rows on either side) times the temporal
    tap matrices, plus the extra taps from global memory.  srcs[0]: [B, F, H, W, Cin]; returns [B*F*H*W, cout]."""
    Wd, Hd, Fd, Bd = fp.out_dims
    mid = emulate(fp.spatial, [srcs[0]], w_spatial, cout) + b_spatial          # [(b f h w), cout]
    mid = mid.reshape(Bd, Fd, Hd, Wd, cout)
    pw, slots = 1 << fp.tile_log2[0], 1 << fp.tile_log2[2]
    nck = fp.mid_chunks * 64
    out = torch.zeros(Bd, Fd, Hd, Wd, cout, dtype=mid.dtype)
    for b in range(Bd):
        for h in range(Hd):
            for w0 in range(0, Wd, pw):
                tile = torch.zeros(pw * (slots + 2), cout, dtype=mid.dtype)     # [16 zero | 128 rows | 16 zero]
                for f in range(Fd):                                              # row = 16 * (f + 1) + w
                    tile[pw * (f + 1):pw * (f + 2)] = mid[b, f, h, w0:w0 + pw]
                acc = torch.zeros(pw * slots, cout, dtype=mid.dtype)
                for ti, t in enumerate(fp.frame_taps):                           # the SAME buffer at row offset 16 (t + 1)
                    view = tile[pw * (t + 1):pw * (t + 1) + pw * slots]
                    acc += view @ w_temporal[:cout, ti * nck:ti * nck + cout].to(mid.dtype).t()
                out[b, :, h, w0:w0 + pw] = acc.reshape(slots, pw, cout)[:Fd]
    k0 = len(fp.frame_taps) * nck
    for (s, off, nch), C in zip(fp.extra, fp.extra_channels):
        assert off == (0, 0, 0, 0)
        out += srcs[s] @ w_temporal[:cout, k0:k0 + C].to(mid.dtype).t()
        k0 += nch * 64
    return out.reshape(-1, cout)
