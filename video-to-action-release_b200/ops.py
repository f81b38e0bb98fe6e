"""Thin Python wrappers over the C ABI: torch tensors in, raw pointers out.

PyTorch is used for device memory and streams only.  Every function here
launches hand-written CUDA through libv2a_b200.so and raises if it cannot.
"""
from __future__ import annotations

import ctypes as C
import itertools
import math
from dataclasses import dataclass
from typing import Optional, Sequence

import torch

from . import _lib

CHUNK_K = 64
ACT_NONE, ACT_SILU, ACT_MISH = 0, 1, 2


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def _require_cuda(*ts: torch.Tensor) -> None:
    for t in ts:
        if t is not None and not t.is_cuda:
            raise RuntimeError("v2a_b200 ops run on CUDA tensors only (no CPU fallback)")


@dataclass
class HL:
    """bf16 (hi, lo) planes of an fp32 tensor: x ~= hi + lo."""

    hi: torch.Tensor
    lo: torch.Tensor
    fp16: bool = False   # planes hold fp16 (hi, lo) bit patterns (in bf16-typed storage) instead of bf16 ones

    @staticmethod
    def empty(rows: int, cols: int, device) -> "HL":
        return HL(torch.empty(rows, cols, dtype=torch.bfloat16, device=device),
                  torch.empty(rows, cols, dtype=torch.bfloat16, device=device))

    def float(self) -> torch.Tensor:
        if self.fp16:
            return self.hi.view(torch.float16).float() + self.lo.view(torch.float16).float()
        return self.hi.float() + self.lo.float()


def split_hl_torch(x: torch.Tensor) -> HL:
    """Weight-time split with torch ops (one-off packing; same RN rounding as the kernels)."""
    hi = x.to(torch.bfloat16)
    lo = (x - hi.float()).to(torch.bfloat16)
    return HL(hi.contiguous(), lo.contiguous())


def split_hl(x: torch.Tensor, ld: Optional[int] = None) -> HL:
    """fp32 [rows, cols] -> HL [rows, ld] (zero padded) on the GPU kernel."""
    _require_cuda(x)
    x = x.contiguous()
    rows, cols = x.shape
    ld = ld or cols
    out = HL.empty(rows, ld, x.device)
    _lib.check(_lib.load().v2a_split_hl(x.data_ptr(), rows, cols, ld, out.hi.data_ptr(),
                                        out.lo.data_ptr(), _stream()), "split_hl")
    return out


# ---------------------------------------------------------------------------
# tile / block heuristics (host logic, unit-tested on CPU)
# ---------------------------------------------------------------------------
def choose_tile(out_dims: Sequence[int], total_log2: int = 7) -> tuple[int, int, int, int]:
    """log2 of the 2^total_log2-row tile box along D0..D3 minimising padded rows (128-row output tiles of
    the implicit GEMM; 64-pixel reduction boxes of the weight-gradient GEMM).

    Ties prefer boxes that are long in the fastest dims (contiguous TMA rows).
    """
    best = None
    for l in itertools.product(range(total_log2 + 1), repeat=4):
        if sum(l) != total_log2:
            continue
        rows = 1
        over = 0
        for d, ld in zip(out_dims, l):
            t = 1 << ld
            rows *= -(-d // t) * t
            if t > d:
                over += 1
        key = (rows, over, tuple(-x for x in l))
        if best is None or key < best[0]:
            best = (key, l)
    return best[1]


def choose_block_n(cout: int, rows: Optional[int] = None, sms: int = 148) -> int:
    """N tile (multiple of 16, <= 256): least padding, then the widest.

    ``rows`` (the launch's output rows) switches on the small-grid rule: when the widest tile leaves fewer than
    ~0.7 x ``sms`` (M tile, N tile) pairs -- the deep levels of the video UNet at batch 1-8: 1792 or 448 rows against
    512-640 channels and K ~ 10 k, which ran as 16-28 CTAs walking 160 k-steps each -- the N tile narrows (multiples of
    32, >= 64, so CTA pairs stay possible) until the grid fills the machine or the floor is reached.  Outputs that
    are bf16 planes or carry GroupNorm sums cannot use split-K (the epilogue needs the finished sum), so the N
    dimension is the parallelism there is."""
    best = None
    for n in range(16, 257, 16):
        tiles = -(-cout // n)
        key = (tiles * n - cout, -n)
        if best is None or key < best[0]:
            best = (key, n)
    bn = best[1]
    if rows is not None:
        import os
        m_tiles = -(-rows // 128)
        want = int(float(os.environ.get("V2A_BN_FILL", "0.7")) * sms)       # env: tuning probes
        floor = int(os.environ.get("V2A_BN_FLOOR", "64"))
        if m_tiles * -(-cout // bn) < want:
            # floor 64: below it every CTA re-reads the activation tile for too few columns (N = 32: 20 x the A
            # traffic at 640 channels)
            cands = [n for n in range(floor, bn + 1, 32) if (-(-cout // n)) * n - cout <= best[0][0] + 31]
            filled = [n for n in cands if m_tiles * -(-cout // n) >= want]
            if filled:
                bn = max(filled)
            elif cands:
                bn = min(cands)
    return bn


def nchunks(c: int) -> int:
    return -(-c // CHUNK_K)


def pack_weight_taps(per_tap: Sequence[torch.Tensor]) -> torch.Tensor:
    """[Cout, Cin_t] per tap -> [Cout, sum 64*ceil(Cin_t/64)] K-major, zero padded per tap."""
    cols = []
    for w in per_tap:
        cout, cin = w.shape
        pad = nchunks(cin) * CHUNK_K - cin
        cols.append(torch.nn.functional.pad(w, (0, pad)) if pad else w)
    return torch.cat(cols, dim=1).contiguous()


# ---------------------------------------------------------------------------
# implicit GEMM plan
# ---------------------------------------------------------------------------
def _igemm_desc(*, srcs, taps, w: HL, out_dims, cout, ldc=None, out_f32=None, out_hl=None,
                bias=None, rowvec=None, rowvec_mul=(0, 0, 0, 0), residual=None, stats=None,
                stats_mul=(0, 0, 0, 0), block_n=None, passes=3, tile_log2=None, out_pix=None, algo_flops_scale=1.0,
                fill_sms=False, split_stride=0):
    """Fill a `v2a_igemm_desc` from torch tensors.  Returns (desc, keep-alive list, meta dict).

    out_pix = ((m0, m1, m2, m3), off): the output row of grid point c is off + sum c[d] * m[d] instead of
    the dense grid order (sub-pixel phases writing into a finer grid); the output tensors are then only
    checked for covering the largest row.  algo_flops_scale: algorithmic / executed FLOPs of this launch
    (9/4 for a sub-pixel phase, which does 4 of the reference's 9 taps' worth of work)."""
    d = _lib.IgemmDesc()
    keep = [srcs, w, out_f32, out_hl, bias, rowvec, residual, stats]
    assert 1 <= len(srcs) <= _lib.V2A_MAX_SRC
    for i, (hl, channels, dims) in enumerate(srcs):
        _require_cuda(hl.hi, hl.lo)
        dims = list(dims) + [1] * (4 - len(dims))
        need = channels * dims[0] * dims[1] * dims[2] * dims[3]
        assert hl.hi.numel() == need, f"src {i}: {hl.hi.numel()} elements, dims say {need}"
        d.src[i].hi = hl.hi.data_ptr()
        d.src[i].lo = hl.lo.data_ptr()
        d.src[i].channels = channels
        for k in range(4):
            d.src[i].dims[k] = dims[k]
    d.nsrc = len(srcs)
    fmts = {bool(hl.fp16) for hl, _, _ in srcs}
    assert fmts == {bool(w.fp16)}, "sources and weights of one igemm share a plane format"
    d.a_fp16 = d.b_fp16 = int(bool(w.fp16))
    assert 1 <= len(taps) <= _lib.V2A_MAX_TAPS
    ktot = 0
    for i, (src, off, nch) in enumerate(taps):
        off = list(off) + [0] * (4 - len(off))
        d.taps[i].src = src
        for k in range(4):
            d.taps[i].d[k] = off[k]
        d.taps[i].nchunks = nch
        ktot += nch * CHUNK_K
    d.ntaps = len(taps)
    assert w.hi.shape[1] == ktot, f"weight K {w.hi.shape[1]} != tap program K {ktot}"
    d.w_hi, d.w_lo = w.hi.data_ptr(), w.lo.data_ptr()
    d.wrows = w.hi.shape[0]
    d.ktot = ktot
    out_dims = list(out_dims) + [1] * (4 - len(out_dims))
    tile_log2 = tile_log2 or choose_tile(out_dims)
    for k in range(4):
        d.out_dims[k] = out_dims[k]
        d.tile_log2[k] = tile_log2[k]
        d.rowvec_mul[k] = rowvec_mul[k]
        d.stats_mul[k] = stats_mul[k]
    d.block_n = block_n or choose_block_n(cout, (out_dims[0] * out_dims[1] * out_dims[2] * out_dims[3]) if fill_sms else None)
    d.split_stride = split_stride        # > 0: split-K partial sums go to their own slices (deterministic reduction)
    d.passes = passes
    d.cout = cout
    rows = out_dims[0] * out_dims[1] * out_dims[2] * out_dims[3]
    out_rows = rows
    if out_pix is not None:
        mul, off = out_pix
        out_rows = off + sum((out_dims[k] - 1) * mul[k] for k in range(4)) + 1     # largest row written + 1
        for k in range(4):
            d.out_pix_mul[k] = mul[k]
        d.out_pix_off = off
        assert any(mul), "out_pix multipliers must not all be zero"
    # outputs may be column windows of wider row-pitched matrices: the pitch is the view's stride(0)
    if ldc is None:
        if out_f32 is not None and out_f32.dim() == 2:
            ldc = out_f32.stride(0)
        elif out_hl is not None and out_hl.hi.dim() == 2:
            ldc = out_hl.hi.stride(0)
        else:
            ldc = -(-cout // 16) * 16
    d.ldc = ldc
    if out_f32 is not None:
        assert out_f32.dtype == torch.float32
        if out_f32.dim() == 2:
            assert out_f32.shape[0] >= out_rows and out_f32.stride(1) == 1 and out_f32.stride(0) == ldc
            assert out_pix is not None or out_f32.shape[0] == rows
        else:
            assert out_f32.numel() >= out_rows * ldc
        d.out_f32 = out_f32.data_ptr()
    if out_hl is not None:
        if out_hl.hi.dim() == 2:
            assert out_hl.hi.shape[0] >= out_rows and out_hl.hi.stride(0) == ldc == out_hl.lo.stride(0)
            assert out_pix is not None or out_hl.hi.shape[0] == rows
        else:
            assert out_hl.hi.numel() >= out_rows * ldc
        d.out_hi, d.out_lo = out_hl.hi.data_ptr(), out_hl.lo.data_ptr()
    if bias is not None:
        assert bias.dtype == torch.float32 and bias.numel() >= cout
        d.bias = bias.data_ptr()
    if rowvec is not None:
        assert rowvec.dtype == torch.float32 and rowvec.dim() == 2
        d.rowvec = rowvec.data_ptr()
        d.ld_rowvec = rowvec.stride(0)
    if residual is not None:
        assert residual.dtype == torch.float32
        d.residual = residual.data_ptr()
        d.ld_res = residual.stride(0) if residual.dim() == 2 else residual.shape[-1]
    if stats is not None:
        assert stats.dtype == torch.float64
        d.stats = stats.data_ptr()
        d.stats_ld = stats.shape[-2]
        # optional leading replica dim: [R, instances, C, 2]
        d.stats_replicas = stats.shape[0] if stats.dim() == 4 else 1
        d.stats_rep_stride = stats.stride(0) if stats.dim() == 4 else 0
    flops = 2.0 * rows * cout * ktot                    # executed by the tensor cores
    meta = dict(rows=rows, cout=cout, ldc=ldc, ktot=ktot, flops=flops, algo_flops=flops * algo_flops_scale,
                out_dims=tuple(out_dims), tile_log2=tuple(tile_log2))
    return d, keep, meta


class Igemm:
    """One planned tcgen05 implicit-GEMM launch (TMA descriptors built once); keyword arguments: `_igemm_desc`."""

    def __init__(self, **kw):
        lib = _lib.load()
        d, self._keep, meta = _igemm_desc(**kw)
        self.rows, self.cout, self.ldc, self.ktot = meta["rows"], meta["cout"], meta["ldc"], meta["ktot"]
        self.desc = d
        self.flops, self.algo_flops = meta["flops"], meta["algo_flops"]
        plan = C.c_void_p()
        _lib.check(lib.v2a_igemm_plan_create(C.byref(d), C.byref(plan)), "igemm_plan_create")
        self._plan = plan
        self._lib = lib
        self.k_splits = int(lib.v2a_igemm_plan_k_splits(plan))

    def run(self) -> None:
        _lib.check(self._lib.v2a_igemm_plan_run(self._plan, _stream()), "igemm_plan_run")

    def __del__(self):
        try:
            if getattr(self, "_plan", None):
                self._lib.v2a_igemm_plan_destroy(self._plan)
                self._plan = None
        except Exception:
            pass


def _contiguous_tiling(out_dims, tile_log2) -> bool:
    """True when the 128-row tile boxes cut the dense row space (D0 fastest) into consecutive 128-row blocks, in tile
    order: every box divides its dim, and once a box is narrower than its dim all higher dims have box 1."""
    partial = False
    for dim, l in zip(out_dims, tile_log2):
        box = 1 << l
        if dim % box != 0 or (partial and box != 1):
            return False
        if box != dim:
            partial = True
    return True


def dual_conv3d_ok(spatial_out_dims, temporal_out_dims, cout: int, passes: int, frames: int) -> bool:
    """Can a Conv3d (spatial program -> temporal program) run as ONE dual launch?  Narrow layers only (the fused
    split product needs block_n <= 128), and both programs must cut the same dense row space into the same blocks."""
    if passes != 3 or cout > 128 or cout % 32 != 0:
        return False
    sd = list(spatial_out_dims) + [1] * (4 - len(spatial_out_dims))
    td = list(temporal_out_dims) + [1] * (4 - len(temporal_out_dims))
    rows_s, rows_t = sd[0] * sd[1] * sd[2] * sd[3], td[0] * td[1] * td[2] * td[3]
    if rows_s != rows_t or td[1] != frames or td[0] % 128 != 0 or (rows_s // 128) < 8:
        return False
    return _contiguous_tiling(sd, choose_tile(sd)) and _contiguous_tiling(td, choose_tile(td))


class IgemmDual:
    """Conv3d as ONE launch: the spatial implicit GEMM and the temporal one that consumes its hi/lo output planes,
    interleaved tile by tile in one persistent kernel (`v2a_igemm_dual_plan_*`, csrc/igemm.cu `igemm_dual_kernel`)."""

    def __init__(self, spatial: dict, temporal: dict, frames: int):
        lib = _lib.load()
        ds, keep_s, ms = _igemm_desc(**spatial)
        dt, keep_t, mt = _igemm_desc(**temporal)
        assert ds.block_n == dt.block_n and ms["rows"] == mt["rows"]
        assert _contiguous_tiling(ms["out_dims"], ms["tile_log2"]) and _contiguous_tiling(mt["out_dims"], mt["tile_log2"])
        tiles = mt["rows"] // 128
        tpf = mt["out_dims"][0] // 128                   # temporal grid = (H*W, F, B): tiles of one (sample, frame)
        assert mt["out_dims"][1] == frames and tiles % (tpf * frames) == 0
        dev = spatial["w"].hi.device
        self.flags = torch.zeros(tiles, dtype=torch.int32, device=dev)
        self._keep = [keep_s, keep_t, self.flags]
        self.desc, self.desc_t = ds, dt
        self.rows, self.cout, self.ldc = mt["rows"], mt["cout"], mt["ldc"]
        self.ktot = ms["ktot"] + mt["ktot"]
        self.flops = ms["flops"] + mt["flops"]
        self.algo_flops = ms["algo_flops"] + mt["algo_flops"]
        self.k_splits = 1
        plan = C.c_void_p()
        _lib.check(lib.v2a_igemm_dual_plan_create(C.byref(ds), C.byref(dt), frames, tpf, self.flags.data_ptr(),
                                                  C.byref(plan)), "igemm_dual_plan_create")
        self._plan, self._lib = plan, lib

    def run(self) -> None:
        _lib.check(self._lib.v2a_igemm_dual_plan_run(self._plan, _stream()), "igemm_dual_plan_run")

    def __del__(self):
        try:
            if getattr(self, "_plan", None):
                self._lib.v2a_igemm_dual_plan_destroy(self._plan)
                self._plan = None
        except Exception:
            pass


def wgrad_passes(default_passes: int, reduction_len: int) -> int:
    """MMA passes of a weight-gradient GEMM whose reduction runs over ``reduction_len`` pixels / rows.

    V2A_WGRAD_PASSES: unset or "3" -> ``default_passes`` (the engine's class); "1" -> one bf16 product everywhere;
    "auto" -> one bf16 product where the reduction is at least V2A_WGRAD_MINK terms long (default 4096), the engine's
    class below that.  (Probe for DESIGN.md: the rounding errors of a long random-sign reduction average out.)"""
    import os
    mode = os.environ.get("V2A_WGRAD_PASSES", "3")
    if mode == "1":
        return 1
    if mode == "auto" and reduction_len >= int(os.environ.get("V2A_WGRAD_MINK", "4096")):
        return 1
    return default_passes


class Wgrad:
    """One planned weight-gradient GEMM (MN-major tcgen05, operands straight from channels-last planes):
    out[(unit, ci), co] += sum_pixels x[pixel + d(unit), 64 * chunk(unit) + ci] * dy[pixel, co]."""

    def __init__(self, *, srcs, units, dy: HL, dy_channels: int, dy_dims, cout: int, out: torch.Tensor,
                 passes: int = 3, box_log2=None):
        lib = _lib.load()
        d = _lib.WgradDesc()
        self._keep = [srcs, dy, out]
        assert 1 <= len(srcs) <= _lib.V2A_MAX_SRC and 1 <= len(units) <= _lib.V2A_WGRAD_MAX_UNITS
        for i, (hl, channels, dims) in enumerate(srcs):
            _require_cuda(hl.hi, hl.lo)
            dims = list(dims) + [1] * (4 - len(dims))
            assert hl.hi.numel() == channels * dims[0] * dims[1] * dims[2] * dims[3], f"src {i}: size mismatch"
            d.src[i].hi, d.src[i].lo, d.src[i].channels = hl.hi.data_ptr(), hl.lo.data_ptr(), channels
            for k in range(4):
                d.src[i].dims[k] = dims[k]
        d.nsrc = len(srcs)
        fmts = {bool(hl.fp16) for hl, _, _ in srcs}
        assert fmts == {bool(dy.fp16)}, "x and dy planes share one format (a tcgen05 MMA takes a single operand format)"
        d.x_fp16 = int(fmts.pop())
        for i, (src, off, chunk) in enumerate(units):
            off = list(off) + [0] * (4 - len(off))
            d.units[i].src, d.units[i].chunk = src, chunk
            for k in range(4):
                d.units[i].d[k] = off[k]
        d.nunits = len(units)
        dy_dims = list(dy_dims) + [1] * (4 - len(dy_dims))
        assert dy.hi.numel() == dy_channels * dy_dims[0] * dy_dims[1] * dy_dims[2] * dy_dims[3]
        d.dy.hi, d.dy.lo, d.dy.channels = dy.hi.data_ptr(), dy.lo.data_ptr(), dy_channels
        box_log2 = box_log2 or choose_tile(dy_dims, 6)
        for k in range(4):
            d.dy.dims[k] = dy_dims[k]
            d.box_log2[k] = box_log2[k]
        assert out.dtype == torch.float32 and out.dim() == 2 and out.shape[0] == 64 * len(units) and out.stride(1) == 1
        d.cout, d.passes, d.out, d.ld_out = cout, passes, out.data_ptr(), out.stride(0)
        self.desc = d
        self.rows, self.cout = 64 * len(units), cout
        pixels = dy_dims[0] * dy_dims[1] * dy_dims[2] * dy_dims[3]
        self.flops = 2.0 * pixels * cout * 64 * len(units)
        plan = C.c_void_p()
        _lib.check(lib.v2a_wgrad_plan_create(C.byref(d), C.byref(plan)), "wgrad_plan_create")
        self._plan, self._lib = plan, lib
        self.k_splits = int(lib.v2a_wgrad_plan_k_splits(plan))

    def run(self) -> None:
        _lib.check(self._lib.v2a_wgrad_plan_run(self._plan, _stream()), "wgrad_plan_run")

    def __del__(self):
        try:
            if getattr(self, "_plan", None):
                self._lib.v2a_wgrad_plan_destroy(self._plan)
                self._plan = None
        except Exception:
            pass


def wgrad_scatter(wt: torch.Tensor, cout: int, cin: int, ntaps: int, dw: torch.Tensor, *, ld_dw: int = 0,
                  accumulate: bool = True) -> None:
    """dw [cout, cin, taps...] (+)= wt[(tap * ceil(cin/64) + chunk) * 64 + ci % 64, co]; ``ld_dw``: row pitch of a
    column window inside a wider [cout, cin_total * taps] gradient (default: contiguous)."""
    _require_cuda(wt, dw)
    if not ld_dw:
        assert dw.is_contiguous() and dw.numel() == cout * cin * ntaps
        ld_dw = cin * ntaps
    _lib.check(_lib.load().v2a_wgrad_scatter(wt.data_ptr(), wt.stride(0), cout, cin, ntaps, dw.data_ptr(), ld_dw,
                                             int(accumulate), _stream()), "wgrad_scatter")


# ---------------------------------------------------------------------------
# elementwise wrappers
# ---------------------------------------------------------------------------
def channel_stats(x: torch.Tensor, instances: int, stats: torch.Tensor) -> None:
    """x fp32 [instances*ppi, C]; stats float64 [instances, C, 2] accumulated in place."""
    _require_cuda(x, stats)
    rows, Cc = x.shape
    _lib.check(_lib.load().v2a_channel_stats(x.data_ptr(), instances, rows // instances, Cc,
                                             stats.data_ptr(), _stream()), "channel_stats")


class Prep:
    """Planned GroupNorm-apply / activation / concat / resample / split launch."""

    def __init__(self, *, x0, x1=None, stats0=None, stats1=None, pixels_per_inst=1, inst_per_group=1,
                 groups=32, eps=1e-5, gamma=None, beta=None, act=ACT_NONE, film=None, pixels_per_film=1,
                 mode=0, H=0, W=0, out_hl=None, out_f32=None, raw_hl=None):
        _require_cuda(x0)
        d = _lib.PrepDesc()
        P, C0 = x0.shape
        C1 = 0 if x1 is None else x1.shape[1]
        d.x0, d.x1, d.C0, d.C1 = x0.data_ptr(), _ptr(x1), C0, C1
        d.stats0, d.stats1 = _ptr(stats0), _ptr(stats1)
        for i, st in enumerate((stats0, stats1)):
            if st is not None and st.dim() == 4:  # [R, instances, C, 2] replicated sums
                setattr(d, f"stats_rep{i}", st.shape[0])
                setattr(d, f"stats_rep_stride{i}", st.stride(0))
        d.pixels_per_inst, d.inst_per_group, d.groups, d.eps = pixels_per_inst, inst_per_group, groups, eps
        self.scratch = None
        if stats0 is not None:
            samples = P // (pixels_per_inst * inst_per_group)
            self.scratch = torch.empty(samples * groups * 2, dtype=torch.float32, device=x0.device)
            d.gn_scratch = self.scratch.data_ptr()
        d.gamma, d.beta, d.act = _ptr(gamma), _ptr(beta), act
        d.film, d.pixels_per_film = _ptr(film), pixels_per_film
        d.mode, d.H, d.W, d.P = mode, H, W, P
        if out_hl is not None:
            d.out_hi, d.out_lo = out_hl.hi.data_ptr(), out_hl.lo.data_ptr()
        d.out_f32 = _ptr(out_f32)
        if raw_hl is not None:
            d.raw_hi, d.raw_lo = raw_hl.hi.data_ptr(), raw_hl.lo.data_ptr()
        self.desc = d
        self._keep = [x0, x1, stats0, stats1, gamma, beta, film, out_hl, out_f32, raw_hl]
        self._lib = _lib.load()

    def run(self) -> None:
        _lib.check(self._lib.v2a_prep(C.byref(self.desc), _stream()), "prep")


def attention(qkv: torch.Tensor, N: int, L: int, heads: int, out: HL) -> None:
    _require_cuda(qkv)
    _lib.check(_lib.load().v2a_attention(qkv.data_ptr(), N, L, heads, out.hi.data_ptr(),
                                         out.lo.data_ptr(), _stream()), "attention")


def linear(x, W, bias, y, *, add=None, act_in=ACT_NONE, act_out=ACT_NONE) -> None:
    """y[b, o] = act_out(act_in(x[b]) . W[o] + bias[o]) (+ add[b, o]); fp32, small batch."""
    _require_cuda(x, W, y)
    B, IN = x.shape
    OUT = W.shape[0]
    assert W.shape[1] == IN and W.is_contiguous() and y.shape[0] == B and y.shape[1] >= OUT
    _lib.check(_lib.load().v2a_linear(x.data_ptr(), x.stride(0), W.data_ptr(), _ptr(bias), _ptr(add),
                                      0 if add is None else add.stride(0), y.data_ptr(), y.stride(0),
                                      B, IN, OUT, act_in, act_out, _stream()), "linear")


def timestep_embedding(t: torch.Tensor, dim: int, mode: int, out: torch.Tensor) -> None:
    _require_cuda(t, out)
    assert t.dtype == torch.int64
    _lib.check(_lib.load().v2a_timestep_embedding(t.data_ptr(), t.numel(), dim, mode, out.data_ptr(),
                                                  _stream()), "timestep_embedding")


def _i64x3(v):
    return (C.c_int64 * 3)(*[int(i) for i in v])


def unet_input_pack(x, x_strides, cond, cond_strides, B, F, H, W, out: HL) -> None:
    """x / cond are device pointers (ints) or tensors; strides are (batch, frame, channel) in elements."""
    xp = x if isinstance(x, int) else x.data_ptr()
    cp = cond if isinstance(cond, int) else cond.data_ptr()
    _lib.check(_lib.load().v2a_unet_input_pack(xp, _i64x3(x_strides), cp, _i64x3(cond_strides), B, F, H, W,
                                               out.hi.data_ptr(), out.lo.data_ptr(), _stream()),
               "unet_input_pack")


def stencil9(P, bias, N, H, W, cout, y) -> None:
    """y[pix][co] = bias[co] + sum over the 3x3 neighbourhood of P[pix + d(tap)][tap * cout + co] (zero padded per image)."""
    _require_cuda(P, y)
    assert P.dim() == 2 and y.dim() == 2 and P.shape[0] == y.shape[0] == N * H * W
    _lib.check(_lib.load().v2a_stencil9(P.data_ptr(), P.stride(0), _ptr(bias), N, H, W, cout, y.data_ptr(), y.stride(0),
                                        _stream()), "stencil9")


def unet_output_head(y, ldy, wt, bt, B, F, H, W, out, out_strides) -> None:
    _lib.check(_lib.load().v2a_unet_output_head(y.data_ptr(), ldy, wt.data_ptr(), bt.data_ptr(), B, F,
                                                H, W, out.data_ptr(), _i64x3(out_strides), _stream()),
               "unet_output_head")


def ddpm_step(x, v, noise, coef) -> None:
    _lib.check(_lib.load().v2a_ddpm_step(x.data_ptr(), v.data_ptr(), _ptr(noise), coef.data_ptr(),
                                         x.numel(), _stream()), "ddpm_step")


def ddim_step(x, v, noise, coef) -> None:
    _lib.check(_lib.load().v2a_ddim_step(x.data_ptr(), v.data_ptr(), _ptr(noise), coef.data_ptr(),
                                         x.numel(), _stream()), "ddim_step")


def cfg_ddpm_step(x, v, noise, coef) -> None:
    """x, v: doubled batch [cond | uncond]; noise: one half; coef: 9+ floats (coef[8] = guidance weight)."""
    _lib.check(_lib.load().v2a_cfg_step(x.data_ptr(), v.data_ptr(), _ptr(noise), coef.data_ptr(), x.numel() // 2, 0,
                                        _stream()), "cfg_step")


def cfg_ddim_step(x, v, noise, coef) -> None:
    _lib.check(_lib.load().v2a_cfg_step(x.data_ptr(), v.data_ptr(), _ptr(noise), coef.data_ptr(), x.numel() // 2, 1,
                                        _stream()), "cfg_step")


def unnormalize_clamp(x, out) -> None:
    _lib.check(_lib.load().v2a_unnormalize_clamp(x.data_ptr(), out.data_ptr(), x.numel(), _stream()),
               "unnormalize_clamp")


# ---- task-token conditioning (row V13): PerceiverResampler pieces, fp32 CUDA cores ----
ACT_GELU = 3


def _rows3(t: torch.Tensor):
    """[B, n, D] view (last dim contiguous) -> (ptr, batch stride, row stride, B, n, D)."""
    assert t.dim() == 3 and t.stride(2) == 1 and t.dtype == torch.float32
    return t.data_ptr(), t.stride(0), t.stride(1), t.shape[0], t.shape[1], t.shape[2]


def pr_layernorm(x, gamma, beta, out, *, pos=None, act=ACT_NONE, eps=1e-5) -> None:
    """out[b, i] = LayerNorm(act(x[b, i]) (+ pos[i])) * gamma (+ beta) over the last dim; x / out [B, n, D] views."""
    _require_cuda(x, gamma, out)
    xp, xb, xl, B, n, D = _rows3(x)
    op, ob, ol, B2, n2, D2 = _rows3(out)
    assert (B, n, D) == (B2, n2, D2) and gamma.numel() == D and gamma.is_contiguous()
    if pos is not None:
        assert pos.is_contiguous() and pos.shape[-1] == D and pos.shape[0] >= n
    _lib.check(_lib.load().v2a_pr_layernorm(xp, xb, xl, _ptr(pos), B, n, D, act, gamma.data_ptr(), _ptr(beta),
                                            float(eps), op, ob, ol, _stream()), "pr_layernorm")


def pr_l2norm_scale(x, heads: int, scale, out) -> None:
    """x / out [rows, heads * dh] (row-strided views): per (row, head) x / max(||x||, 1e-12) * scale[dh]."""
    _require_cuda(x, scale, out)
    rows, width = x.shape
    dh = width // heads
    assert x.stride(1) == 1 and out.stride(1) == 1 and out.shape == x.shape and scale.numel() == dh
    _lib.check(_lib.load().v2a_pr_l2norm_scale(x.data_ptr(), x.stride(0), rows, heads, dh, scale.data_ptr(),
                                               out.data_ptr(), out.stride(0), _stream()), "pr_l2norm_scale")


def pr_attention(q, k, v, B: int, heads: int, scale: float, out) -> None:
    """q [B*nq, heads*dh], k / v [B*nk, heads*dh] (row-strided views) -> out [B*nq, heads*dh]."""
    _require_cuda(q, k, v, out)
    nq, nk = q.shape[0] // B, k.shape[0] // B
    dh = q.shape[1] // heads
    assert all(t.stride(1) == 1 for t in (q, k, v, out)) and k.shape == v.shape and out.shape == q.shape
    _lib.check(_lib.load().v2a_pr_attention(q.data_ptr(), q.stride(0), k.data_ptr(), k.stride(0), v.data_ptr(),
                                            v.stride(0), B, heads, dh, nq, nk, float(scale), out.data_ptr(),
                                            out.stride(0), _stream()), "pr_attention")


def pr_token_mean(x, out) -> None:
    """out[b] = mean_i x[b, i]; x [B, n, D] view, out [B, D]."""
    _require_cuda(x, out)
    xp, xb, xl, B, n, D = _rows3(x)
    assert out.shape == (B, D) and out.stride(1) == 1
    _lib.check(_lib.load().v2a_pr_token_mean(xp, xb, xl, B, n, D, out.data_ptr(), out.stride(0), _stream()),
               "pr_token_mean")


def pr_broadcast_rows(src, out) -> None:
    """out[b, i] = src[i]; src [n, D] contiguous, out [B, n, D] view."""
    _require_cuda(src, out)
    op, ob, ol, B, n, D = _rows3(out)
    assert src.is_contiguous() and tuple(src.shape) == (n, D)
    _lib.check(_lib.load().v2a_pr_broadcast_rows(src.data_ptr(), n, D, B, op, ob, ol, _stream()), "pr_broadcast_rows")


def add_rows_(dst, src) -> None:
    """dst[r, c] += src[r, c] for row-strided fp32 [rows, C] views."""
    _require_cuda(dst, src)
    assert dst.shape == src.shape and dst.stride(1) == 1 and src.stride(1) == 1
    _lib.check(_lib.load().v2a_add_strided(dst.data_ptr(), dst.stride(0), src.data_ptr(), src.stride(0),
                                           dst.shape[0], dst.shape[1], 1, _stream()), "add_strided")


# ---- content fingerprint of a parameter list (guards the packed-weight caches, see packing.content_key) ----
_FP_TABLES: dict = {}
_FP_CHUNK = 1 << 16


def params_fingerprint(params: Sequence[torch.Tensor], key=None, launch_only: bool = False):
    """64-bit fingerprint of the VALUES of ``params`` (CUDA fp32 tensors): one kernel launch over a cached
    device table of (pointer, length, global offset) chunks, then an 8-byte read-back (this synchronises).
    ``key``: any hashable that changes whenever a parameter's storage does (callers that already hold the
    (data_ptr, _version) tuple pass it to save a second walk over the parameters).  ``launch_only``: enqueue the
    kernel on the current stream and return the int64[1] device tensor it writes (the caller reads it back later)."""
    ps = [p.detach() for p in params if p.numel()]
    if not ps:
        return torch.zeros(1, dtype=torch.int64) if launch_only else 0
    _require_cuda(*ps)
    dev = ps[0].device
    key = (str(dev), key) if key is not None else (str(dev),) + tuple((p.data_ptr(), p.numel()) for p in ps)
    ent = _FP_TABLES.get(key)
    if ent is None:
        rows, off = [], 0
        for p in ps:
            if p.dtype != torch.float32 or not p.is_contiguous():
                raise RuntimeError("v2a_b200: engine parameters must be contiguous fp32 tensors")
            n, base = p.numel(), p.data_ptr()
            for lo in range(0, n, _FP_CHUNK):
                rows.append((base + 4 * lo, min(_FP_CHUNK, n - lo), off + lo))
            off += n
        table = torch.tensor(rows, dtype=torch.int64).to(dev)
        if len(_FP_TABLES) >= 32:
            _FP_TABLES.pop(next(iter(_FP_TABLES)))
        ent = _FP_TABLES[key] = (table, torch.zeros(1, dtype=torch.int64, device=dev), len(rows))
    table, out, n = ent
    _lib.check(_lib.load().v2a_params_fingerprint(table.data_ptr(), n, out.data_ptr(), _stream()), "params_fingerprint")
    return out if launch_only else int(out.item())


def launch_count() -> int:
    return int(_lib.load().v2a_launch_count())
