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


# ---- Upsample = nearest x2 on (H, W) then 3x3 conv (guided_diffusion/unet.py:107-114), as four sub-pixel phases ----
# Output pixel (2i + py, 2j + px) of the fine grid reads fine rows 2i + py + kh - 1, i.e. COARSE rows
#   py = 0: i - 1 (kh = 0), i (kh = 1, 2)        py = 1: i (kh = 0, 1), i + 1 (kh = 2)
# and the same along W.  Taps that hit the same coarse pixel share its value, so their weights add: each phase is a
# 2x2-tap conv over the coarse grid (4 instead of 9 taps: 2.25x fewer MACs than convolving the upsampled tensor, which
# is never materialised).  The fine grid's zero padding (rows -1 and 2H) falls on coarse rows -1 and H: still the TMA
# out-of-bounds zero fill.
_UP_ROWS = {0: ((-1, (0,)), (0, (1, 2))), 1: ((0, (0, 1)), (1, (2,)))}     # phase -> ((coarse offset, kernel taps), ...)


def upsample3x3_phase(cin: int, N: int, H: int, W: int, py: int, px: int) -> ConvProgram:
    """Phase (py, px) over the COARSE grid (H, W = input sizes); tap order (row offset, col offset) row-major."""
    taps = [(0, (dw, dh, 0, 0), nchunks(cin)) for dh, _ in _UP_ROWS[py] for dw, _ in _UP_ROWS[px]]
    return ConvProgram([cin], [(W, H, N, 1)], taps, (W, H, N, 1))


def upsample3x3_phase_weight(w: torch.Tensor, py: int, px: int) -> torch.Tensor:
    """[Cout, Cin, 3, 3] -> [Cout, 4 * pad64(Cin)]: per coarse tap the sum of the kernel taps that land on it."""
    return pack_weight_taps([sum(w[:, :, kh, kw] for kh in khs for kw in kws)
                             for _, khs in _UP_ROWS[py] for _, kws in _UP_ROWS[px]])


def upsample3x3_out_pix(H: int, W: int, py: int, px: int):
    """(multipliers, offset) of the fine-grid row written for coarse grid point (j, i, n): ((n*2H + 2i+py)*2W + 2j+px)."""
    return (2, 4 * W, 4 * H * W, 0), py * 2 * W + px


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


@dataclass
class FusedConv3dProgram:
    """Two-stage program of ONE Conv3d launch (docs/FUSED_CONV3D_PLAN.md; host description only — the kernel that
    consumes it is next round's work, nothing on the product path builds one yet).

    stage 1 = ``spatial``: 3x3 taps over the grid (W, H, F, B) into an on-chip tile of 16 pixels x 8 frame slots
    (``tile_log2``; slot F..7 is out of bounds = zero filled); stage 2 = ``frame_taps`` shifts of that tile by whole
    frames (zero padded) times the temporal tap matrices, plus ``extra`` taps read from global memory (the
    ResBlock's 1x1 skip conv).  Weights: ``spatial3x3_weight`` and ``temporal3_weight``, unchanged."""
    spatial: ConvProgram
    tile_log2: Tuple[int, int, int, int]
    frame_taps: Tuple[int, ...]
    mid_chunks: int
    extra: List[Tuple[int, Tuple[int, int, int, int], int]]
    extra_channels: List[int]
    out_dims: Tuple[int, int, int, int]
    ktot2: int = field(init=False)

    def __post_init__(self):
        self.ktot2 = 64 * (len(self.frame_taps) * self.mid_chunks + sum(t[2] for t in self.extra))


def fused3d(cin: int, cout: int, B: int, F: int, H: int, W: int, skip_channels: int = 0) -> FusedConv3dProgram:
    if W % 16 != 0 or F > 8:
        raise ValueError("fused Conv3d tiles are 16 pixels along W x 8 frame slots")
    taps = [(0, (kw - 1, kh - 1, 0, 0), nchunks(cin)) for kh in range(3) for kw in range(3)]
    spatial = ConvProgram([cin], [(W, H, F, B)], taps, (W, H, F, B))
    extra, extra_ch = [], []
    if skip_channels:
        extra.append((1, (0, 0, 0, 0), nchunks(skip_channels)))
        extra_ch.append(skip_channels)
    return FusedConv3dProgram(spatial, (4, 0, 3, 0), (-1, 0, 1), nchunks(cout), extra, extra_ch, (W, H, F, B))


def pointwise(cin: int, dims: Sequence[int]) -> ConvProgram:
    d = tuple(list(dims) + [1] * (4 - len(dims)))
    return ConvProgram([cin], [d], [(0, (0, 0, 0, 0), nchunks(cin))], d)


def pointwise_weight(w: torch.Tensor) -> torch.Tensor:
    return pack_weight_taps([w.reshape(w.shape[0], w.shape[1])])


def taps_as_columns_weight(w: torch.Tensor) -> torch.Tensor:
    """Out head: [Cout, Cin, 3, 3] (Cout tiny) -> [9 * Cout, pad64(Cin)], row = (kh*3 + kw) * Cout + co: the 3x3 conv as a
    1x1 conv to 9*Cout columns whose taps `ops.stencil9` gathers afterwards."""
    cout, cin = w.shape[:2]
    return pack_weight_taps([w.permute(2, 3, 0, 1).reshape(9 * cout, cin)])


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


# ---- policy: stride-2 Downsample1d = Conv1d(C, C, 3, 2, 1) on the phase view [B][T/2][2][C] ----
def down1d(c: int, B: int, T: int) -> ConvProgram:
    """input t = 2*o + k - 1:  k=0 -> (phase 1, o-1); k=1 -> (0, o); k=2 -> (1, o).  No copy:
    [B][T][C] memory IS [B][T/2][2][C]."""
    taps = [(0, (1, -1, 0, 0), nchunks(c)), (0, (0, 0, 0, 0), nchunks(c)), (0, (1, 0, 0, 0), nchunks(c))]
    return ConvProgram([c], [(2, T // 2, B, 1)], taps, (1, T // 2, B, 1))


def down1d_dgrad(c: int, B: int, T: int) -> ConvProgram:
    """dx[2j] = W1^T dy[j]; dx[2j+1] = W2^T dy[j] + W0^T dy[j+1]: one GEMM with N = 2C whose
    output row (b, j) holds [phase 0 | phase 1] = the [B][T][C] layout of dx.  T = INPUT length."""
    taps = [(0, (0, 0, 0, 0), nchunks(c)), (0, (1, 0, 0, 0), nchunks(c))]
    return ConvProgram([c], [(T // 2, B, 1, 1)], taps, (T // 2, B, 1, 1))


def down1d_dgrad_weight(w: torch.Tensor) -> torch.Tensor:
    """[Cout, Cin, 3] -> [2*Cin, 2*pad64(Cout)]: rows p=0: [W1^T | 0]; rows p=1: [W2^T | W0^T]."""
    z = torch.zeros_like(w[:, :, 0].t())
    return torch.cat([pack_weight_taps([w[:, :, 1].t(), z]), pack_weight_taps([w[:, :, 2].t(), w[:, :, 0].t()])], 0)


# ---- policy: Upsample1d = ConvTranspose1d(C, C, 4, 2, 1), weight [Cin, Cout, 4] ----
def up1d(c: int, B: int, T: int) -> ConvProgram:
    """out[2j] = Wt1^T x[j] + Wt3^T x[j-1]; out[2j+1] = Wt2^T x[j] + Wt0^T x[j+1] (+bias): one
    GEMM with N = 2C over taps x[j-1], x[j], x[j+1]; output row (b, j) = [phase 0 | phase 1]."""
    taps = [(0, (d, 0, 0, 0), nchunks(c)) for d in (-1, 0, 1)]
    return ConvProgram([c], [(T, B, 1, 1)], taps, (T, B, 1, 1))


def up1d_weight(wt: torch.Tensor) -> torch.Tensor:
    """[Cin, Cout, 4] -> [2*Cout, 3*pad64(Cin)]."""
    z = torch.zeros_like(wt[:, :, 0].t())
    p0 = pack_weight_taps([wt[:, :, 3].t(), wt[:, :, 1].t(), z])
    p1 = pack_weight_taps([z, wt[:, :, 2].t(), wt[:, :, 0].t()])
    return torch.cat([p0, p1], 0)


def up1d_dgrad(c: int, B: int, T: int) -> ConvProgram:
    """dx[i] = sum_k Wt[:, :, k] dy[2i - 1 + k] on the phase view of dy [B][T][2][C] (T = INPUT length):
    k=0 -> (1, i-1); k=1 -> (0, i); k=2 -> (1, i); k=3 -> (0, i+1)."""
    taps = [(0, (1, -1, 0, 0), nchunks(c)), (0, (0, 0, 0, 0), nchunks(c)), (0, (1, 0, 0, 0), nchunks(c)),
            (0, (0, 1, 0, 0), nchunks(c))]
    return ConvProgram([c], [(2, T, B, 1)], taps, (1, T, B, 1))


def up1d_dgrad_weight(wt: torch.Tensor) -> torch.Tensor:
    """[Cin, Cout, 4] -> [Cin, 4*pad64(Cout)]."""
    return pack_weight_taps([wt[:, :, k] for k in range(4)])


def conv1d_cat(cins, B: int, T: int, k: int, pad: int) -> ConvProgram:
    """Conv1d over a channel concat of up to two sources without materialising the concat."""
    taps, dims = [], []
    for s, c in enumerate(cins):
        taps += [(s, (j - pad, 0, 0, 0), nchunks(c)) for j in range(k)]
        dims.append((T, B, 1, 1))
    return ConvProgram(list(cins), dims, taps, (T, B, 1, 1))


def conv1d_cat_weight(w: torch.Tensor, cins) -> torch.Tensor:
    parts, off = [], 0
    for c in cins:
        parts += [w[:, off:off + c, j] for j in range(w.shape[2])]
        off += c
    return pack_weight_taps(parts)
