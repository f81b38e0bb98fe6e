"""GPU parity of the observation-encoder path: the MN-major weight-gradient
GEMM, the stride-2 data-gradient weights, and the whole VisualCore forward + backward against the same
modules run with stock torch ops in fp32 (TF32 off).  Tolerance: 1e-3 relative L2 (north_star)."""
import os

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu
TOL = 1e-3


def rel_l2(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


@pytest.fixture(autouse=True)
def _fp32_truth():
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    yield


def _planes(x):
    from v2a_b200 import ops
    return ops.split_hl(x.reshape(-1, x.shape[-1]).contiguous())


@pytest.mark.parametrize("N,H,W,Ci,Co", [(3, 8, 8, 64, 64), (2, 16, 16, 128, 256), (5, 4, 4, 512, 512),
                                         (2, 32, 32, 64, 128)])
def test_wgrad_3x3_matches_torch(N, H, W, Ci, Co):
    from v2a_b200 import ops
    torch.manual_seed(0)
    x = torch.randn(N, H, W, Ci, device="cuda")
    dy = torch.randn(N, H, W, Co, device="cuda")
    taps9 = [(kw - 1, kh - 1, 0, 0) for kh in range(3) for kw in range(3)]
    units = [(0, d, ch) for d in taps9 for ch in range(ops.nchunks(Ci))]
    scratch = torch.zeros(64 * len(units), Co, device="cuda")
    g = ops.Wgrad(srcs=[(_planes(x), Ci, (W, H, N, 1))], units=units, dy=_planes(dy), dy_channels=Co,
                  dy_dims=(W, H, N, 1), cout=Co, out=scratch)
    g.run()
    dw = torch.zeros(Co, Ci, 3, 3, device="cuda")
    ops.wgrad_scatter(scratch, Co, Ci, 9, dw)
    want = torch.nn.grad.conv2d_weight(x.permute(0, 3, 1, 2), (Co, Ci, 3, 3), dy.permute(0, 3, 1, 2), padding=1)
    assert rel_l2(dw, want) < 1e-5, (g.k_splits, rel_l2(dw, want))
    # one bf16 product (V2A_WGRAD_PASSES probe): stages then hold no lo planes -- the first version still placed the
    # dy operand behind two x planes and ran off the stage; result = the bf16-rounded operands' exact product
    scratch1 = torch.zeros_like(scratch)
    ops.Wgrad(srcs=[(_planes(x), Ci, (W, H, N, 1))], units=units, dy=_planes(dy), dy_channels=Co,
              dy_dims=(W, H, N, 1), cout=Co, out=scratch1, passes=1).run()
    dw1 = torch.zeros(Co, Ci, 3, 3, device="cuda")
    ops.wgrad_scatter(scratch1, Co, Ci, 9, dw1)
    want1 = torch.nn.grad.conv2d_weight(x.bfloat16().float().permute(0, 3, 1, 2), (Co, Ci, 3, 3),
                                        dy.bfloat16().float().permute(0, 3, 1, 2), padding=1)
    assert rel_l2(dw1, want1) < 1e-5 and rel_l2(dw1, want) < 2e-2


def test_wgrad_stride2_phase_split_and_pointwise():
    from v2a_b200 import convs, ops
    torch.manual_seed(1)
    N, H, W, Ci, Co = 3, 16, 16, 64, 128
    x = torch.randn(N, H, W, Ci, device="cuda")
    dy = torch.randn(N, H // 2, W // 2, Co, device="cuda")
    # phase-split operand [N][py*2+px][H/2][W/2][C]
    xps = x.reshape(N, H // 2, 2, W // 2, 2, Ci).permute(0, 2, 4, 1, 3, 5).contiguous()
    prog = convs.spatial3x3_s2(Ci, N, H, W)
    units = [(0, tuple(t[1]), ch) for t in prog.taps for ch in range(ops.nchunks(Ci))]
    scratch = torch.zeros(64 * len(units), Co, device="cuda")
    ops.Wgrad(srcs=[(_planes(xps), Ci, (W // 2, H // 2, 4, N))], units=units, dy=_planes(dy), dy_channels=Co,
              dy_dims=(W // 2, H // 2, 1, N), cout=Co, out=scratch).run()
    dw = torch.zeros(Co, Ci, 3, 3, device="cuda")
    ops.wgrad_scatter(scratch, Co, Ci, 9, dw)
    want = torch.nn.grad.conv2d_weight(x.permute(0, 3, 1, 2), (Co, Ci, 3, 3), dy.permute(0, 3, 1, 2), stride=2, padding=1)
    assert rel_l2(dw, want) < 1e-5
    # pointwise with a narrow (32-channel) gradient: the keypoint conv
    xs = torch.randn(48, 512, device="cuda")
    dl = torch.randn(48, 32, device="cuda")
    sc = torch.zeros(512, 32, device="cuda")
    ops.Wgrad(srcs=[(_planes(xs), 512, (48, 1, 1, 1))], units=[(0, (0, 0, 0, 0), ch) for ch in range(8)], dy=_planes(dl),
              dy_channels=32, dy_dims=(48, 1, 1, 1), cout=32, out=sc).run()
    assert rel_l2(sc, xs.t() @ dl) < 1e-5


def test_dgrad_stride2_weights_match_torch():
    """One GEMM over dy producing the 4 input phases == conv_transpose of the stride-2 3x3 conv."""
    from v2a_b200 import obs_encoder as OE, ops
    torch.manual_seed(2)
    N, Ho, Wo, Ci, Co = 2, 8, 8, 64, 128
    w = torch.randn(Co, Ci, 3, 3, device="cuda") / 10
    dy = torch.randn(N, Ho, Wo, Co, device="cuda")
    wd = ops.split_hl_torch(OE.dgrad3x3_s2_weight(w))
    taps4 = [(0, (di, dj, 0, 0), ops.nchunks(Co)) for dj in range(2) for di in range(2)]
    blocked = torch.zeros(N * Ho * Wo, 4 * Ci, device="cuda")
    ops.Igemm(srcs=[(_planes(dy), Co, (Wo, Ho, N, 1))], taps=taps4, w=wd, out_dims=(Wo, Ho, N, 1), cout=4 * Ci,
              out_f32=blocked).run()
    dx = torch.zeros(N * 2 * Ho * 2 * Wo, Ci, device="cuda")
    from v2a_b200 import _lib
    _lib.check(_lib.load().v2a_enc_unblock_add(blocked.data_ptr(), None, N, 2 * Ho, 2 * Wo, Ci, dx.data_ptr(),
                                               torch.cuda.current_stream().cuda_stream), "unblock")
    want = torch.nn.grad.conv2d_input((N, Ci, 2 * Ho, 2 * Wo), w, dy.permute(0, 3, 1, 2), stride=2, padding=1)
    assert rel_l2(dx.view(N, 2 * Ho, 2 * Wo, Ci).permute(0, 3, 1, 2), want) < 1e-5


def _seeded_core():
    from v2a_b200 import diffusion_policy as DP
    torch.manual_seed(0)
    pol = DP.build_libero_policy()
    core = pol.obs_encoder.key_model_map["img_obs_1"]
    g = torch.Generator().manual_seed(5)
    with torch.no_grad():
        for p in core.parameters():      # GroupNorm gains 1 / biases 0 would hide gamma / beta handling
            if p.numel():
                p.add_(0.1 * torch.randn(p.shape, generator=g))
    return core.cuda()


@pytest.mark.parametrize("B", [3, 8])
def test_visual_core_forward_backward_matches_torch(monkeypatch, B):
    core = _seeded_core()
    core.train()
    torch.manual_seed(3)
    x = torch.rand(B, 3, 128, 128, device="cuda") * 2 - 1
    wout = torch.randn(B, 64, device="cuda")
    # truth: the same modules through stock torch ops in float64 (cuDNN's fp32 Winograd / FFT algorithms are
    # themselves ~1e-3 off on the 64-channel 32x32 layers, measured here: see the fp32 column of the report)
    import copy
    core64 = copy.deepcopy(core).double()
    ref = core64.nets(x.double())          # stock-op twin on the parameter-holder modules (tests/stock_twins.py)
    (ref * wout.double()).sum().backward()
    want = {n: p.grad.float() for n, p in core64.named_parameters() if p.grad is not None}
    ref = ref.float()
    ref32 = core.nets(x)
    (ref32 * wout).sum().backward()
    fp32_err = {n: (p.grad - want[n]).norm().item() / max(want[n].norm().item(), 1e-30)
                for n, p in core.named_parameters() if p.grad is not None}
    core.zero_grad(set_to_none=True)
    from v2a_b200 import ops
    n0 = ops.launch_count()
    got = core(x)
    assert ops.launch_count() > n0, "the CUDA engine did not run"
    assert rel_l2(got, ref) < TOL, rel_l2(got, ref)
    (got * wout).sum().backward()
    scale = max(g.norm().item() for g in want.values())
    errs = {}
    num = den = 0.0
    for n, p in core.named_parameters():
        if n not in want:
            continue
        assert p.grad is not None, n
        # structurally-zero gradients (softmax shift invariance -> pool.nets.bias) are judged on the global scale
        errs[n] = (p.grad - want[n]).norm().item() / max(want[n].norm().item(), 1e-5 * scale)
        num += (p.grad - want[n]).double().pow(2).sum().item()
        den += want[n].double().pow(2).sum().item()
    total = (num / den) ** 0.5
    med = sorted(errs.values())[len(errs) // 2]
    report = "\n".join(f"{n}: ours {e:.3e}  torch-fp32 {fp32_err[n]:.3e}" for n, e in errs.items())
    print(f"VisualCore B={B}: forward rel-L2 {rel_l2(got, ref):.2e}, whole-gradient rel-L2 {total:.2e}, "
          f"median / worst per-parameter {med:.2e} / {max(errs.values()):.2e}")
    # ReLU is discontinuous: where a pre-activation is within the forward's rounding error of zero the mask, and
    # with it that element's gradient, flips against the float64 truth (tools/debug_encoder.py pins the deviating
    # elements: e.g. truth +8.0e-6 vs ours -1.4e-5 on O(1) operands).  ONE flipped element of a late layer
    # (8 x 4x4x512 activations) moves every upstream gradient by ~1/sqrt(#elements) ~ 4e-3, so no two fp32
    # implementations with different summation orders agree to 1e-3 on these gradients -- torch's own fp32 path is
    # 2-5e-4 off the float64 truth here (fp32_err).  Bars: the forward at the north-star 1e-3 (measured 2.4e-5; floor = the
    # tensor core's truncating fp32 accumulate), flip-free parameters sit at ~5e-5 (B = 3: median), and a flip-limited
    # bar on the whole gradient.  (B = 8 on this seed: a flip in layer3 lifts every upstream parameter to ~1e-3.)
    # (which elements flip depends on the accumulation order, i.e. on tile shapes and kernel versions: these are
    # sanity bars against structural errors, which show up as O(1); the mask-for-mask test below is the tight one)
    assert total < 3e-2, f"whole gradient rel-L2 {total:.3e}\n{report}"
    assert max(errs.values()) < 6e-2, report


def test_visual_core_matches_cpu_oracle_and_reference_golden():
    """CUDA engine vs oracle/encoder_oracle.py (CPU, fp32) and vs the golden vectors of the UNMODIFIED reference
    VisualCore (tests/golden/make_encoder_golden.py): features at 1e-3 (measured ~2e-5), gradient fingerprints at
    the flip-limited bar explained above."""
    import json
    from oracle import encoder_oracle as EO
    from oracle import policy_oracle as PO
    from tests.golden.configs import encoder_inputs, grad_fingerprint
    from v2a_b200 import diffusion_policy as DP
    here = os.path.dirname(os.path.abspath(__file__))
    with open(os.path.join(here, "golden", "encoder_golden_meta.json")) as f:
        meta = json.load(f)
    with open(os.path.join(here, "golden", "policy_loss_golden_meta.json")) as f:
        layout = json.load(f)["layout"]
    pol = DP.build_libero_policy()
    sd = pol.state_dict()
    sd.update(PO.seeded_full_policy_state_dict(layout, meta["seed"]))
    pol.load_state_dict(sd, strict=True)
    sd_cpu = {k: v.detach().clone() for k, v in pol.state_dict().items()}
    core = pol.obs_encoder.key_model_map["img_obs_1"].cuda()
    core.train()
    x, w = encoder_inputs(meta["B"], meta["seed"])
    got = core(x.cuda())
    with torch.no_grad():
        want = EO.visual_core_forward(sd_cpu, meta["key"], x)
    gold = torch.load(os.path.join(here, "golden", "encoder_golden.pt"))
    assert rel_l2(got, want) < TOL and rel_l2(got, gold["feat"]) < TOL
    (got * w.cuda()).sum().backward()
    floor = 1e-5 * max(v[0] for v in meta["grad_fingerprints"].values())
    worst = 0.0
    for n, p in core.named_parameters():
        if n not in meta["grad_fingerprints"]:
            continue
        norm, proj = meta["grad_fingerprints"][n]
        n2, p2 = grad_fingerprint(meta["key"] + n, p.grad)
        worst = max(worst, abs(n2 - norm) / max(norm, floor))
        assert abs(n2 - norm) <= 5 * TOL * max(norm, floor), (n, n2, norm)
        assert abs(p2 - proj) <= 5 * TOL * max(norm, floor) * p.numel() ** 0.5, (n, p2, proj)
    print(f"VisualCore vs reference golden: features rel-L2 {rel_l2(got, gold['feat']):.2e}, worst gradient-norm deviation {worst:.2e}")


def test_visual_core_full_batch_is_batch_independent():
    """configs[2] size (B = 256): GroupNorm never crosses images, so the features of a sub-batch equal the features
    of the same images inside the full batch, and a gradient restricted to that sub-batch equals the sub-batch's own."""
    core = _seeded_core()
    core.eval()
    torch.manual_seed(11)
    x = torch.rand(256, 3, 128, 128, device="cuda") * 2 - 1
    w = torch.zeros(256, 64, device="cuda")
    w[:8] = torch.randn(8, 64, device="cuda")
    full = core(x)
    (full * w).sum().backward()
    g_full = {n: p.grad.clone() for n, p in core.named_parameters() if p.grad is not None}
    core.zero_grad(set_to_none=True)
    sub = core(x[:8])
    (sub * w[:8]).sum().backward()
    assert rel_l2(full[:8], sub) < 1e-4      # different tile shapes -> different fp32 accumulation orders
    num = sum((g_full[n] - p.grad).double().pow(2).sum().item() for n, p in core.named_parameters() if n in g_full)
    den = sum(p.grad.double().pow(2).sum().item() for n, p in core.named_parameters() if n in g_full)
    assert (num / den) ** 0.5 < 3e-2, (num / den) ** 0.5     # flip-limited (see above)


@pytest.mark.parametrize("B", [3, 8])
def test_visual_core_gradients_match_float64_mask_for_mask(B):
    """The tight backward check: the float64 truth is evaluated with OUR forward's ReLU sign patterns (every
    ReLU of the oracle replaced by multiplication with the mask the CUDA engine produced), so the gradients are
    comparable element for element and the ReLU discontinuity drops out.  Every parameter gradient then has to
    meet the 1e-3 tolerance (measured ~5e-5; the stem's three parameters also sit behind the MaxPool arg-max, whose
    near-ties are left to chance: ~1e-5 of the windows, each worth ~1/sqrt(3M elements))."""
    from oracle import encoder_oracle as EO
    from v2a_b200 import obs_encoder as OE
    core = _seeded_core()
    core.train()
    torch.manual_seed(3)
    x = torch.rand(B, 3, 128, 128, device="cuda") * 2 - 1
    wout = torch.randn(B, 64, device="cuda")
    got = core(x)
    (got * wout).sum().backward()
    eng = OE.last_engine(core)

    def nchw(t2d, a):
        return t2d.view(a.N, a.H, a.W, a.C).permute(0, 3, 1, 2)
    masks = {}
    for k, (a1, out) in enumerate(zip(eng.inner_acts, eng.acts[1:])):
        masks[2 * k + 1] = nchw(a1.float() > 0, out).double()
        masks[2 * k + 2] = nchw(out.f32 > 0, out).double()
    # stem: sign pattern of GroupNorm(raw0) recomputed from the engine's conv output and statistics
    sp = eng.stem_probe
    gn0 = core.backbone.nets[1]
    raw0 = sp["raw0"].view(B, sp["H"], sp["W"], -1).permute(0, 3, 1, 2).double()
    mr0 = sp["mr0"].view(B, gn0.num_groups, 2).double()
    cpg = raw0.shape[1] // gn0.num_groups
    mean = mr0[:, :, 0].repeat_interleave(cpg, 1)[:, :, None, None]
    rstd = mr0[:, :, 1].repeat_interleave(cpg, 1)[:, :, None, None]
    pre0 = (raw0 - mean) * rstd * gn0.weight.detach().double()[None, :, None, None] + gn0.bias.detach().double()[None, :, None, None]
    masks[0] = (pre0 > 0).double()
    key = "k."
    sd = {key + n: p.detach().double().clone().requires_grad_(True) for n, p in core.named_parameters()}
    sd.update({key + n: b.detach().double() for n, b in core.named_buffers()})
    relu = lambda t, i: t * masks[i] if i in masks else torch.relu(t)
    ref = EO.visual_core_forward(sd, key, x.double(), relu)
    assert rel_l2(got, ref) < TOL
    (ref * wout.double()).sum().backward()
    worst, report = 0.0, []
    scale = max(v.grad.norm().item() for k_, v in sd.items() if v.requires_grad and v.grad is not None)
    for n, p in core.named_parameters():
        g = sd[key + n].grad
        if g is None or p.grad is None:
            continue
        err = (p.grad.double() - g).norm().item() / max(g.norm().item(), 1e-4 * scale)
        report.append(f"{n}: {err:.2e}")
        if n == "backbone.nets.0.weight":
            # the stem conv sits behind the MaxPool: a near-tie between two candidates of ONE window (pinned with
            # tools/debug_stem.py: image 2, pixels (54,34) / (55,35) on this seed) routes that window's gradient to the
            # other pixel -> 3-5e-3 on this one parameter; the kernels themselves measure 4e-6 on the same operands
            assert err < 2e-2, report[-1]
            continue
        worst = max(worst, err)
    print(f"VisualCore B={B} mask-for-mask: worst parameter-gradient rel-L2 {worst:.2e}")
    assert worst < TOL, "\n".join(report)
