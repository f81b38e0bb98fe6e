"""Tap programs + weight packing for every convolution shape on the hot paths.

Pure host logic (no CUDA): each helper returns the implicit-GEMM description
(`ConvProgram`) that `ops.Igemm` turns into TMA descriptors, and the matching
K-major packed weight matrix.  tests/ replays the same programs through a CPU
emulator of the kernel's semantics to pin this logic against torch convs.

Layouts (channels-last, X0 fastest):
  spatial 3x3        src [N][H][W][C]            dims (W, H, N, 1)
  spatial 3x3 s2     src [N][4 phases][H/2][W/2][C]   dims (W/2, H/2, 4, N)
  temporal k3        src [B][F][H*W][C]          dims (H*W, F, B, 1)
  pointwise          src [rows][C]               dims (rows, 1, 1, 1) or (L, N, 1, 1)
  conv1d k (policy)  src [B][T][C]               dims (T, B, 1, 1)
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import List, Sequence, Tuple

import torch

from .ops import nchunks, pack_weight_taps


@dataclass
class ConvProgram:
    src_channels: List[int]
    src_dims: List[Tuple[int, int, int, int]]
    taps: List[Tuple[int, Tuple[int, int, int, int], int]]  # (src, offset, nchunks)
    out_dims: Tuple[int, int, int, int]
    ktot: int = field(init=False)

    def __post_init__(self):
        self.ktot = 64 * sum(t[2] for t in self.taps)


# ---- video: Conv3d = spatial Conv2d + temporal Conv1d (guided_diffusion/nn.py:30-87) ----
def spatial3x3(cin: int, N: int, H: int, W: int) -> ConvProgram:
    taps = [(0, (kw - 1, kh - 1, 0, 0), nchunks(cin)) for kh in range(3) for kw in range(3)]
    return ConvProgram([cin], [(W, H, N, 1)], taps, (W, H, N, 1))


def spatial3x3_weight(w: torch.Tensor) -> torch.Tensor:
    """[Cout, Cin, 3, 3] -> [Cout, 9 * pad64(Cin)], tap-major (kh, kw)."""
    return pack_weight_taps([w[:, :, kh, kw] for kh in range(3) for kw in range(3)])


def spatial3x3_s2(cin: int, N: int, H: int, W: int) -> ConvProgram:
    """Stride-2 conv on a phase-split input (H, W are the INPUT sizes).

    input row 2*oh + kh - 1: kh=0 -> odd phase, index oh-1; kh=1 -> even phase,
    index oh; kh=2 -> odd phase, index oh.  Same along W.
    """
    taps = []
    for kh in range(3):
        for kw in range(3):
            ph, dh = (1, -1) if kh == 0 else ((0, 0) if kh == 1 else (1, 0))
            pw, dw = (1, -1) if kw == 0 else ((0, 0) if kw == 1 else (1, 0))
            taps.append((0, (dw, dh, ph * 2 + pw, 0), nchunks(cin)))
    return ConvProgram([cin], [(W // 2, H // 2, 4, N)], taps, (W // 2, H // 2, 1, N))


def temporal3(c: int, B: int, F: int, HW: int, skip_channels: int = 0) -> ConvProgram:
    """Conv1d(C, C, 3) over frames, zero padded; optional 1x1 skip conv as extra K."""
    taps = [(0, (0, k - 1, 0, 0), nchunks(c)) for k in range(3)]
    chans, dims = [c], [(HW, F, B, 1)]
    if skip_channels:
        taps.append((1, (0, 0, 0, 0), nchunks(skip_channels)))
        chans.append(skip_channels)
        dims.append((HW, F, B, 1))
    return ConvProgram(chans, dims, taps, (HW, F, B, 1))


def temporal3_weight(wt: torch.Tensor, wskip: torch.Tensor | None = None) -> torch.Tensor:
    """[C, C, 3] (+ [C, Cx, 1, 1] or [C, Cx]) -> [C, 3*pad64(C) (+ pad64(Cx))]."""
    parts = [wt[:, :, k] for k in range(3)]
    if wskip is not None:
        parts.append(wskip.reshape(wskip.shape[0], wskip.shape[1]))
    return pack_weight_taps(parts)


def pointwise(cin: int, dims: Sequence[int]) -> ConvProgram:
    d = tuple(list(dims) + [1] * (4 - len(dims)))
    return ConvProgram([cin], [d], [(0, (0, 0, 0, 0), nchunks(cin))], d)


def pointwise_weight(w: torch.Tensor) -> torch.Tensor:
    return pack_weight_taps([w.reshape(w.shape[0], w.shape[1])])


def input_conv_weight(w: torch.Tensor) -> torch.Tensor:
    """First conv [Cout, 6, 3, 3] against the im2col'd 64-wide operand: k = tap*6 + c."""
    cout = w.shape[0]
    k = w.permute(0, 2, 3, 1).reshape(cout, 54)
    return torch.nn.functional.pad(k, (0, 10)).contiguous()


# ---- policy: Conv1d over the horizon axis (diffusion_policy/model/conv1d_components.py) ----
def conv1d(cin: int, B: int, T: int, k: int, pad: int) -> ConvProgram:
    taps = [(0, (j - pad, 0, 0, 0), nchunks(cin)) for j in range(k)]
    return ConvProgram([cin], [(T, B, 1, 1)], taps, (T, B, 1, 1))


def conv1d_weight(w: torch.Tensor) -> torch.Tensor:
    """[Cout, Cin, k] -> [Cout, k*pad64(Cin)]."""
    return pack_weight_taps([w[:, :, j] for j in range(w.shape[2])])


def conv1d_dgrad_weight(w: torch.Tensor) -> torch.Tensor:
    """Weights of the data-gradient conv: dx[t] = sum_j W[:, :, j]^T dy[t + pad - j].

    Run as conv1d over dy with taps offset (pad - j), i.e. kernel flipped and
    in/out channels swapped: [Cin, k*pad64(Cout)], tap order j = k-1 .. 0 so the
    tap offsets ascend like the forward program's.
    """
    k = w.shape[2]
    return pack_weight_taps([w[:, :, j].t() for j in reversed(range(k))])
