"""CPU tests of the observation-encoder host logic: data-gradient weight
layouts and weight-gradient unit programs, replayed through tests/emulator.py in float64 against torch's own
convolution gradients; structure of the planned engine's parameter bookkeeping."""
import pytest
import torch
import torch.nn.functional as F

from tests.emulator import _shifted, as5d, emulate
from v2a_b200 import convs, obs_encoder as OE, ops

D = torch.float64


def _nhwc(x):
    return x.permute(0, 2, 3, 1).reshape(-1, x.shape[1])


def test_dgrad3x3_weight_is_the_conv_input_gradient():
    torch.manual_seed(0)
    N, Ci, Co, H, W = 2, 24, 40, 6, 10
    w = torch.randn(Co, Ci, 3, 3, dtype=D)
    dy = torch.randn(N, Co, H, W, dtype=D)
    prog = convs.spatial3x3(Co, N, H, W)
    out = emulate(prog, [as5d(_nhwc(dy), Co, prog.src_dims[0])], OE.dgrad3x3_weight(w), Ci)
    ref = _nhwc(torch.nn.grad.conv2d_input((N, Ci, H, W), w, dy, padding=1))
    torch.testing.assert_close(out, ref, rtol=1e-12, atol=1e-12)


def test_dgrad3x3_stride2_four_phase_gemm():
    """One GEMM over dy with taps (dj, di) in {0,1}^2 and N = 4*Cin yields the input gradient phase-blocked
    [img][H/2][W/2][(py, px)][Cin] (what v2a_enc_unblock_add re-orders)."""
    torch.manual_seed(1)
    N, Ci, Co, Ho, Wo = 2, 16, 24, 4, 6
    w = torch.randn(Co, Ci, 3, 3, dtype=D)
    dy = torch.randn(N, Co, Ho, Wo, dtype=D)
    prog = convs.ConvProgram([Co], [(Wo, Ho, N, 1)],
                             [(0, (di, dj, 0, 0), ops.nchunks(Co)) for dj in range(2) for di in range(2)], (Wo, Ho, N, 1))
    out = emulate(prog, [as5d(_nhwc(dy), Co, (Wo, Ho, N, 1))], OE.dgrad3x3_s2_weight(w), 4 * Ci)
    ref = torch.nn.grad.conv2d_input((N, Ci, 2 * Ho, 2 * Wo), w, dy, stride=2, padding=1)      # [N, Ci, H, W]
    blocked = ref.permute(0, 2, 3, 1).reshape(N, Ho, 2, Wo, 2, Ci).permute(0, 1, 3, 2, 4, 5).reshape(-1, 4 * Ci)
    torch.testing.assert_close(out, blocked, rtol=1e-12, atol=1e-12)


def _emulate_wgrad(src5d, units, dy5d, dy_dims):
    """out[(unit, ci), co] = sum_pixels x[pixel + d(unit), 64*chunk + ci] * dy[pixel, co] (zero outside)."""
    rows = []
    co = dy5d.shape[-1]
    dyf = dy5d.reshape(-1, co)
    for (s, d, chunk) in units:
        xs = _shifted(src5d[s], d, dy_dims).reshape(dyf.shape[0], -1)[:, 64 * chunk:64 * chunk + 64]
        xs = F.pad(xs, (0, 64 - xs.shape[1]))
        rows.append(xs.t() @ dyf)
    return torch.cat(rows, 0)


def _scatter(wt, cout, cin, ntaps):
    nchunk = ops.nchunks(cin)
    dw = torch.zeros(cout, cin, ntaps, dtype=wt.dtype)
    for tap in range(ntaps):
        for ci in range(cin):
            dw[:, ci, tap] = wt[(tap * nchunk + ci // 64) * 64 + ci % 64]
    return dw


@pytest.mark.parametrize("stride", [1, 2])
def test_wgrad_unit_program_matches_conv2d_weight(stride):
    torch.manual_seed(2)
    N, Ci, Co, H, W = 2, 72, 16, 8, 8
    x = torch.randn(N, Ci, H, W, dtype=D)
    Ho, Wo = H // stride, W // stride
    dy = torch.randn(N, Co, Ho, Wo, dtype=D)
    if stride == 1:
        units = [(0, (kw - 1, kh - 1, 0, 0), ch) for kh in range(3) for kw in range(3) for ch in range(ops.nchunks(Ci))]
        src = as5d(_nhwc(x), Ci, (W, H, N, 1))
        dy_dims = (W, H, N, 1)
    else:
        prog = convs.spatial3x3_s2(Ci, N, H, W)
        units = [(0, tuple(t[1]), ch) for t in prog.taps for ch in range(ops.nchunks(Ci))]
        src = x.permute(0, 2, 3, 1).reshape(N, Ho, 2, Wo, 2, Ci).permute(0, 2, 4, 1, 3, 5).reshape(N, 4, Ho, Wo, Ci)
        dy_dims = (Wo, Ho, 1, N)
    wt = _emulate_wgrad([src], units, as5d(_nhwc(dy), Co, dy_dims), dy_dims)
    dw = _scatter(wt, Co, Ci, 9).reshape(Co, Ci, 3, 3)
    ref = torch.nn.grad.conv2d_weight(x, (Co, Ci, 3, 3), dy, stride=stride, padding=1)
    torch.testing.assert_close(dw, ref, rtol=1e-11, atol=1e-11)


def test_choose_tile_for_the_64_pixel_reduction_box():
    for dims in [(32, 32, 256, 1), (4, 4, 256, 1), (8, 8, 1, 256), (4096, 256, 1, 1), (48, 1, 1, 1)]:
        box = ops.choose_tile(dims, 6)
        assert sum(box) == 6
        assert all((1 << b) <= max(1, 2 * d) for b, d in zip(box, dims))


def test_visual_core_cuda_path_fails_loudly_on_cpu(monkeypatch):
    from v2a_b200 import diffusion_policy as DP
    pol = DP.build_libero_policy()
    core = pol.obs_encoder.key_model_map["img_obs_1"]
    assert tuple(pol.obs_encoder.output_shape()) == (128,)          # read off the modules, no forward pass
    with pytest.raises(RuntimeError, match="CUDA"):
        core(torch.zeros(1, 3, 128, 128))
