"""CPU tests of the host-side conv programs (tap offsets, packing, phase split)
replayed through tests/emulator.py and compared with torch's own convolutions."""
import pytest
import torch
import torch.nn.functional as F

from tests.emulator import as5d, emulate
from v2a_b200 import convs, ops

torch.manual_seed(0)
D = torch.float64


def _nhwc(x):  # [N, C, H, W] -> [N*H*W, C]
    return x.permute(0, 2, 3, 1).reshape(-1, x.shape[1])


def test_spatial3x3_matches_conv2d():
    N, Ci, Co, H, W = 3, 72, 40, 8, 16
    x = torch.randn(N, Ci, H, W, dtype=D)
    w = torch.randn(Co, Ci, 3, 3, dtype=D)
    prog = convs.spatial3x3(Ci, N, H, W)
    out = emulate(prog, [as5d(_nhwc(x), Ci, prog.src_dims[0])], convs.spatial3x3_weight(w), Co)
    ref = _nhwc(F.conv2d(x, w, padding=1))
    assert prog.ktot == 9 * 128
    torch.testing.assert_close(out, ref, rtol=1e-12, atol=1e-12)


def test_spatial3x3_stride2_phase_split():
    N, Ci, Co, H, W = 2, 16, 24, 8, 12
    x = torch.randn(N, Ci, H, W, dtype=D)
    w = torch.randn(Co, Ci, 3, 3, dtype=D)
    # phase split exactly as prep mode 2 lays it out: [img][ph*2+pw][H/2][W/2][C]
    xp = x.permute(0, 2, 3, 1).reshape(N, H // 2, 2, W // 2, 2, Ci).permute(0, 2, 4, 1, 3, 5)
    xp = xp.reshape(N, 4, H // 2, W // 2, Ci)
    prog = convs.spatial3x3_s2(Ci, N, H, W)
    out = emulate(prog, [xp], convs.spatial3x3_weight(w), Co)
    ref = _nhwc(F.conv2d(x, w, padding=1, stride=2))
    torch.testing.assert_close(out, ref, rtol=1e-12, atol=1e-12)


def test_fused_conv3d_program_matches_conv2d_then_padded_conv1d():
    """The two-stage program of the planned one-kernel Conv3d (docs/FUSED_CONV3D_PLAN.md): 16-pixel x 8-frame-slot
    tiles, temporal taps as whole-frame shifts of the on-chip tile, spatial bias BEFORE the temporal zero padding
    (guided_diffusion/nn.py:72-85), optional 1x1 skip conv as extra K."""
    from tests.emulator import emulate_fused3d
    B, Fr, H, W, Ci, Co, Cx = 2, 7, 3, 32, 24, 40, 16
    x = torch.randn(B, Fr, H, W, Ci, dtype=D)
    xs = torch.randn(B, Fr, H, W, Cx, dtype=D)
    ws, bs = torch.randn(Co, Ci, 3, 3, dtype=D), torch.randn(Co, dtype=D)
    wt, wk = torch.randn(Co, Co, 3, dtype=D), torch.randn(Co, Cx, 1, 1, dtype=D)
    fp = convs.fused3d(Ci, Co, B, Fr, H, W, skip_channels=Cx)
    assert fp.spatial.ktot == 9 * 64 and fp.ktot2 == 3 * 64 + 64 and fp.tile_log2 == (4, 0, 3, 0)
    out = emulate_fused3d(fp, [x, xs], convs.spatial3x3_weight(ws), bs, convs.temporal3_weight(wt, wk), Co)
    y = F.conv2d(x.reshape(B * Fr, H, W, Ci).permute(0, 3, 1, 2), ws, bs, padding=1)          # [(b f), Co, H, W]
    yy = y.reshape(B, Fr, Co, H, W).permute(0, 3, 4, 2, 1).reshape(B * H * W, Co, Fr)         # '(b h w) c f'
    ref_t = F.conv1d(F.pad(yy, (1, 1)), wt).reshape(B, H, W, Co, Fr).permute(0, 4, 1, 2, 3)   # [B, F, H, W, Co]
    ref = ref_t + torch.einsum("bfhwc,oc->bfhwo", xs, wk[:, :, 0, 0])
    torch.testing.assert_close(out, ref.reshape(-1, Co), rtol=1e-11, atol=1e-11)
    with pytest.raises(ValueError):
        convs.fused3d(Ci, Co, B, Fr, H, 8)


def test_temporal_with_skip_matches_conv1d_plus_1x1():
    B, Fr, HW, C, Cx = 2, 7, 6, 24, 40
    y = torch.randn(B, Fr, HW, C, dtype=D)
    xs = torch.randn(B, Fr, HW, Cx, dtype=D)
    wt = torch.randn(C, C, 3, dtype=D)
    ws = torch.randn(C, Cx, 1, 1, dtype=D)
    prog = convs.temporal3(C, B, Fr, HW, skip_channels=Cx)
    # src layout [X3=1][X2=B][X1=F][X0=HW][C]
    out = emulate(prog, [y[None], xs[None]], convs.temporal3_weight(wt, ws), C)
    # reference: '(b h w) c f' conv1d with zero padding 1 (guided_diffusion/nn.py:76-85)
    yy = y.permute(0, 2, 3, 1).reshape(B * HW, C, Fr)
    ref_t = F.conv1d(F.pad(yy, (1, 1)), wt).reshape(B, HW, C, Fr).permute(0, 3, 1, 2)
    ref_s = torch.einsum("bfpc,oc->bfpo", xs, ws[:, :, 0, 0])
    ref = (ref_t + ref_s).reshape(-1, C)
    torch.testing.assert_close(out, ref, rtol=1e-12, atol=1e-12)


def test_conv1d_k5_and_dgrad_weights():
    B, T, Ci, Co = 3, 8, 24, 16
    x = torch.randn(B, Ci, T, dtype=D, requires_grad=True)
    w = torch.randn(Co, Ci, 5, dtype=D)
    y = F.conv1d(x, w, padding=2)
    prog = convs.conv1d(Ci, B, T, 5, 2)
    xl = x.detach().permute(0, 2, 1).reshape(1, 1, B, T, Ci)
    out = emulate(prog, [xl], convs.conv1d_weight(w), Co)
    torch.testing.assert_close(out, y.detach().permute(0, 2, 1).reshape(-1, Co), rtol=1e-12, atol=1e-12)
    # data gradient = same program over dy with flipped / transposed weights
    dy = torch.randn_like(y)
    (dx,) = torch.autograd.grad(y, x, dy)
    progb = convs.conv1d(Co, B, T, 5, 2)
    dyl = dy.permute(0, 2, 1).reshape(1, 1, B, T, Co)
    outb = emulate(progb, [dyl], convs.conv1d_dgrad_weight(w), Ci)
    torch.testing.assert_close(outb, dx.permute(0, 2, 1).reshape(-1, Ci), rtol=1e-12, atol=1e-12)


def test_input_conv_weight_matches_im2col():
    Co, H, W = 8, 5, 6
    x = torch.randn(1, 6, H, W, dtype=D)
    w = torch.randn(Co, 6, 3, 3, dtype=D)
    cols = F.unfold(x, 3, padding=1)  # [1, 6*9, H*W], index c*9 + tap
    cols = cols.reshape(6, 9, H * W).permute(2, 1, 0).reshape(H * W, 54)  # k = tap*6 + c
    out = F.pad(cols, (0, 10)) @ convs.input_conv_weight(w).t()
    ref = _nhwc(F.conv2d(x, w, padding=1))
    torch.testing.assert_close(out, ref, rtol=1e-12, atol=1e-12)


def test_choose_tile_and_block_n():
    for dims in [(16384, 7, 16, 1), (64, 7, 16, 1), (8, 8, 112, 1), (16, 256, 1, 1), (4, 256, 1, 1), (1, 1, 1, 1)]:
        t = ops.choose_tile(dims)
        assert sum(t) == 7
    assert ops.choose_tile((16384, 7, 16, 1)) == (7, 0, 0, 0)
    # 8x8 frames: pair two batch elements rather than padding 7 frames to 8
    assert ops.choose_tile((64, 7, 16, 1)) == (6, 0, 1, 0)
    assert ops.choose_block_n(128) == 128 and ops.choose_block_n(512) == 256
    assert ops.choose_block_n(384) == 192 and ops.choose_block_n(640) == 160
    assert ops.choose_block_n(3) == 16


def test_bf16_split_error_bound():
    x = torch.randn(4096) * torch.logspace(-3, 3, 4096)
    hl = ops.split_hl_torch(x)
    err = (hl.float() - x).abs() / x.abs().clamp_min(1e-30)
    assert err.max() < 2.0 ** -16


def _im2col_t(x, offsets, stride, Tout):
    """CPU twin of v2a_policy_im2col_t: x [B, Tin, C] -> [C*ntaps, B*Tout]."""
    B, Tin, C = x.shape
    out = torch.zeros(C, len(offsets), B, Tout, dtype=x.dtype)
    for k, off in enumerate(offsets):
        for o in range(Tout):
            t = stride * o + off
            if 0 <= t < Tin:
                out[:, k, :, o] = x[:, t, :].t()
    return out.reshape(C * len(offsets), B * Tout)


def test_downsample1d_forward_dgrad_wgrad_programs():
    B, T, C = 3, 8, 16
    x = torch.randn(B, C, T, dtype=D, requires_grad=True)
    w = torch.randn(C, C, 3, dtype=D, requires_grad=True)
    y = F.conv1d(x, w, stride=2, padding=1)
    dy = torch.randn_like(y)
    dx, dw = torch.autograd.grad(y, (x, w), dy)
    xl = x.detach().permute(0, 2, 1).contiguous()                 # [B, T, C]
    prog = convs.down1d(C, B, T)
    out = emulate(prog, [xl.reshape(1, B, T // 2, 2, C)], convs.conv1d_weight(w.detach()), C)
    torch.testing.assert_close(out, y.detach().permute(0, 2, 1).reshape(-1, C), rtol=1e-12, atol=1e-12)
    dyl = dy.permute(0, 2, 1).contiguous()                        # [B, T/2, C]
    progb = convs.down1d_dgrad(C, B, T)
    outb = emulate(progb, [dyl.reshape(1, 1, B, T // 2, C)], convs.down1d_dgrad_weight(w.detach()), 2 * C)
    torch.testing.assert_close(outb.reshape(B, T, C), dx.permute(0, 2, 1), rtol=1e-12, atol=1e-12)
    # wgrad: dW[co, ci, k] = dy^T [C, B*T/2] @ im2col^T(x)[C*3, B*T/2]^T
    dyT = dyl.reshape(-1, C).t()
    cols = _im2col_t(xl, [-1, 0, 1], 2, T // 2)
    torch.testing.assert_close((dyT @ cols.t()).reshape(C, C, 3), dw, rtol=1e-12, atol=1e-12)


def test_upsample1d_forward_dgrad_wgrad_programs():
    B, T, C = 2, 4, 16
    x = torch.randn(B, C, T, dtype=D, requires_grad=True)
    wt = torch.randn(C, C, 4, dtype=D, requires_grad=True)       # ConvTranspose1d weight [Cin, Cout, k]
    b = torch.randn(C, dtype=D)
    y = F.conv_transpose1d(x, wt, b, stride=2, padding=1)          # [B, C, 2T]
    dy = torch.randn_like(y)
    dx, dw = torch.autograd.grad(y, (x, wt), dy)
    xl = x.detach().permute(0, 2, 1).contiguous()
    prog = convs.up1d(C, B, T)
    out = emulate(prog, [xl.reshape(1, 1, B, T, C)], convs.up1d_weight(wt.detach()), 2 * C) + torch.cat([b, b])
    torch.testing.assert_close(out.reshape(B, 2 * T, C), y.detach().permute(0, 2, 1), rtol=1e-12, atol=1e-12)
    dyl = dy.permute(0, 2, 1).contiguous()                        # [B, 2T, C]
    progb = convs.up1d_dgrad(C, B, T)
    outb = emulate(progb, [dyl.reshape(1, B, T, 2, C)], convs.up1d_dgrad_weight(wt.detach()), C)
    torch.testing.assert_close(outb.reshape(B, T, C), dx.permute(0, 2, 1), rtol=1e-12, atol=1e-12)
    # wgrad: dWt[ci, co, k] = x^T [C, B*T] @ im2col^T(dy, stride 2, offsets k-1)[C*4, B*T]^T
    xT = xl.reshape(-1, C).t()
    cols = _im2col_t(dyl, [-1, 0, 1, 2], 2, T)
    torch.testing.assert_close((xT @ cols.t()).reshape(C, C, 4), dw, rtol=1e-12, atol=1e-12)


def test_conv1d_wgrad_via_transposed_im2col_and_concat_slices():
    B, T, C0, C1, Co = 2, 8, 8, 16, 24
    x = torch.randn(B, C0 + C1, T, dtype=D)
    w = torch.randn(Co, C0 + C1, 5, dtype=D, requires_grad=True)
    y = F.conv1d(x, w, padding=2)
    dy = torch.randn_like(y)
    (dw,) = torch.autograd.grad(y, w, dy)
    dyT = dy.permute(0, 2, 1).reshape(-1, Co).t()
    xl = x.permute(0, 2, 1).contiguous()
    got = torch.zeros(Co, (C0 + C1) * 5, dtype=D)
    for off, cpart in ((0, C0), (C0, C1)):  # concat parts land in column slices of [Co, Cin*k]
        cols = _im2col_t(xl[:, :, off:off + cpart].contiguous(), [-2, -1, 0, 1, 2], 1, T)
        got[:, off * 5:(off + cpart) * 5] = dyT @ cols.t()
    torch.testing.assert_close(got.reshape(Co, C0 + C1, 5), dw, rtol=1e-12, atol=1e-12)
    # forward over a 2-source concat: taps of source 0 then source 1, weights sliced the same way
    prog = convs.conv1d_cat([C0, C1], B, T, 5, 2)
    wpk = convs.conv1d_cat_weight(w.detach(), [C0, C1])
    out = emulate(prog, [xl[:, :, :C0].reshape(1, 1, B, T, C0), xl[:, :, C0:].reshape(1, 1, B, T, C1)], wpk, Co)
    torch.testing.assert_close(out, y.permute(0, 2, 1).reshape(-1, Co), rtol=1e-12, atol=1e-12)


def test_upsample_subpixel_phases_match_interpolate_then_conv2d():
    """`Upsample` (guided_diffusion/unet.py:107-114: nearest x2 on (H, W), then 3x3 conv) as four 2x2-tap convs over
    the coarse grid whose outputs interleave on the fine grid (`out_pix` addressing of the kernel)."""
    N, Ci, Co, H, W = 2, 24, 40, 5, 6
    x = torch.randn(N, Ci, H, W, dtype=D)
    w = torch.randn(Co, Ci, 3, 3, dtype=D)
    ref = _nhwc(F.conv2d(F.interpolate(x, scale_factor=2, mode="nearest"), w, padding=1))     # [N*2H*2W, Co]
    out = torch.full_like(ref, float("nan"))
    for py in range(2):
        for px in range(2):
            prog = convs.upsample3x3_phase(Ci, N, H, W, py, px)
            assert prog.ktot == 4 * 64 and prog.out_dims == (W, H, N, 1)
            y = emulate(prog, [as5d(_nhwc(x), Ci, prog.src_dims[0])], convs.upsample3x3_phase_weight(w, py, px), Co)
            (m0, m1, m2, _), off = convs.upsample3x3_out_pix(H, W, py, px)
            j, i, n = torch.meshgrid(torch.arange(W), torch.arange(H), torch.arange(N), indexing="ij")
            rows = (off + j * m0 + i * m1 + n * m2).permute(2, 1, 0).reshape(-1)               # grid order: n, i, j
            out[rows] = y
    assert not torch.isnan(out).any()                                                         # every fine pixel written once
    torch.testing.assert_close(out, ref, rtol=1e-12, atol=1e-12)


def stencil9_reference(P, bias, N, H, W, cout):
    """What `v2a_stencil9` computes, with torch indexing: y[n,h,w,co] = bias[co] + sum_taps P[n,h+kh-1,w+kw-1,tap*cout+co]."""
    Pp = F.pad(P.reshape(N, H, W, -1), (0, 0, 1, 1, 1, 1))
    y = bias.to(P.dtype).expand(N, H, W, cout).clone()
    for kh in range(3):
        for kw in range(3):
            t = kh * 3 + kw
            y = y + Pp[:, kh:kh + H, kw:kw + W, t * cout:(t + 1) * cout]
    return y.reshape(N * H * W, cout)


def test_out_head_as_1x1_conv_plus_nine_tap_gather_matches_conv2d():
    """The out head's Conv2d(128, 3, 3x3) (guided_diffusion/unet.py:628-632) as a 1x1 conv to 27 columns followed by a
    9-point gather (`convs.taps_as_columns_weight` + `ops.stencil9`)."""
    N, Ci, Co, H, W = 3, 72, 3, 6, 5
    x = torch.randn(N, Ci, H, W, dtype=D)
    w, b = torch.randn(Co, Ci, 3, 3, dtype=D), torch.randn(Co, dtype=D)
    prog = convs.pointwise(Ci, (N * H * W,))
    wp = convs.taps_as_columns_weight(w)
    assert tuple(wp.shape) == (27, 128)
    P = emulate(prog, [as5d(_nhwc(x), Ci, prog.src_dims[0])], wp, 27)
    torch.testing.assert_close(stencil9_reference(P, b, N, H, W, Co), _nhwc(F.conv2d(x, w, b, padding=1)),
                               rtol=1e-12, atol=1e-12)


def test_dual_conv3d_eligibility_and_contiguous_tilings():
    """`ops.dual_conv3d_ok` decides whether a Conv3d's spatial and temporal programs may share ONE dual launch: both
    must cut the same dense row space into the same consecutive 128-row blocks (the temporal tile m then depends on the
    spatial tiles m - tpf, m, m + tpf only) and the layer must be narrow enough for the fused split product."""
    # Libero levels: 128x128, 64x64 (stride-2 output grid), tiny-model 16x16
    assert ops.dual_conv3d_ok((128, 128, 112, 1), (16384, 7, 16, 1), 128, 3, 7)
    assert ops.dual_conv3d_ok((64, 64, 1, 112), (4096, 7, 16, 1), 128, 3, 7)
    assert ops.dual_conv3d_ok((16, 16, 6, 1), (256, 3, 2, 1), 64, 3, 3)
    assert not ops.dual_conv3d_ok((128, 128, 112, 1), (16384, 7, 16, 1), 256, 3, 7)      # wide layer: two launches
    assert not ops.dual_conv3d_ok((128, 128, 112, 1), (16384, 7, 16, 1), 128, 1, 7)      # one-pass class
    assert not ops.dual_conv3d_ok((8, 8, 112, 1), (64, 7, 16, 1), 128, 3, 7)             # tiles straddle frames
    assert not ops.dual_conv3d_ok((24, 16, 6, 1), (384, 3, 2, 1), 64, 3, 3)              # ragged boxes
    # contiguous tiling <=> tile m covers dense rows [128 m, 128 m + 128)
    for dims in [(128, 128, 5, 1), (64, 64, 3, 2), (16, 16, 6, 1), (4096, 7, 2, 1), (32, 8, 4, 3)]:
        tl = ops.choose_tile(dims)
        if not ops._contiguous_tiling(dims, tl):
            continue
        ntile = [d >> l for d, l in zip(dims, tl)]
        idx = torch.arange(dims[0] * dims[1] * dims[2] * dims[3]).reshape(dims[3], dims[2], dims[1], dims[0])
        m = 0
        for j3 in range(ntile[3]):
            for j2 in range(ntile[2]):
                for j1 in range(ntile[1]):
                    for j0 in range(ntile[0]):
                        box = idx[j3 << tl[3]:(j3 + 1) << tl[3], j2 << tl[2]:(j2 + 1) << tl[2],
                                  j1 << tl[1]:(j1 + 1) << tl[1], j0 << tl[0]:(j0 + 1) << tl[0]].reshape(-1)
                        assert box.tolist() == list(range(128 * m, 128 * m + 128)), (dims, tl, m)
                        m += 1


def test_block_n_small_grid_rule():
    """ops.choose_block_n with the launch's row count: unchanged where the widest tile already fills the machine
    (every B = 16 layer), narrower N tiles (>= 64, multiples of 32, no extra padding beyond one 32-column step) on the
    deep levels at small batch."""
    from v2a_b200 import ops
    for cout in (128, 256, 384, 512, 640):
        assert ops.choose_block_n(cout, 114688) == ops.choose_block_n(cout)        # full resolution, any batch
    for cout in (512, 640):                                                              # the channel counts of that level
        assert ops.choose_block_n(cout, 16 * 7 * 64) == ops.choose_block_n(cout)         # B = 16 at 8 x 8
    assert ops.choose_block_n(512, 1792) == 64 and ops.choose_block_n(640, 448) == 64     # B = 1 at 16 x 16 / 8 x 8
    assert ops.choose_block_n(512, 3584) == 128                                          # B = 2: 28 x 4 = 112 tiles
    assert ops.choose_block_n(1536, 1792) == 192                                         # qkv: 14 x 8 tiles, no padding
    for cout in (128, 256, 384, 512, 640, 1536, 1920):
        for rows in (448, 1792, 3584, 7168):
            bn = ops.choose_block_n(cout, rows)
            assert 64 <= bn <= 256 and (bn % 32 == 0 or bn == ops.choose_block_n(cout))   # narrowed tiles keep CTA pairs


def test_wgrad_passes_switch(monkeypatch):
    """ops.wgrad_passes: the engine's class by default; V2A_WGRAD_PASSES=1 / auto (+ V2A_WGRAD_MINK) are probes."""
    from v2a_b200 import ops
    monkeypatch.delenv("V2A_WGRAD_PASSES", raising=False)
    monkeypatch.delenv("V2A_WGRAD_MINK", raising=False)
    assert ops.wgrad_passes(3, 10) == 3 and ops.wgrad_passes(3, 1 << 20) == 3 and ops.wgrad_passes(1, 5) == 1
    monkeypatch.setenv("V2A_WGRAD_PASSES", "1")
    assert ops.wgrad_passes(3, 10) == 1
    monkeypatch.setenv("V2A_WGRAD_PASSES", "auto")
    assert ops.wgrad_passes(3, 4095) == 3 and ops.wgrad_passes(3, 4096) == 1
    monkeypatch.setenv("V2A_WGRAD_MINK", "1024")
    assert ops.wgrad_passes(3, 1024) == 1 and ops.wgrad_passes(3, 1023) == 3
