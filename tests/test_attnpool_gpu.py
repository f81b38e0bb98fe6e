"""Row V13 on the GPU: the PerceiverResampler kernels (csrc/perceiver.cu) one by one against torch, and the whole
`task_attnpool(y).mean(1)` on them against the CPU oracle at the Libero sizes."""
import pytest
import torch

from tests.stock_twins import _task_pool
import torch.nn.functional as F

from oracle import video_oracle as VO
from tests.test_attnpool_host import _TorchOps

pytestmark = pytest.mark.gpu


def _rel(a, b):
    return ((a.double() - b.double()).norm() / b.double().norm()).item()


def test_kernels_against_torch():
    from v2a_b200 import ops
    g = torch.Generator().manual_seed(0)
    B, n, D, heads = 3, 7, 96, 4
    big = torch.randn(B, n + 5, D, generator=g).cuda()
    x = big[:, 2:2 + n]                                            # strided [B, n, D] view, like kvin[:, :n]
    gamma, beta = torch.randn(D, generator=g).cuda(), torch.randn(D, generator=g).cuda()
    pos = torch.randn(n + 3, D, generator=g).cuda()
    for act in (0, 3):
        for b_, p_ in ((beta, pos), (None, None)):
            out_big = torch.zeros(B, n + 4, D, device="cuda")
            ops.pr_layernorm(x, gamma, b_, out_big[:, 4:], pos=p_, act=act)
            want = torch.empty(B, n, D, device="cuda")
            _TorchOps.pr_layernorm(x, gamma, b_, want, pos=p_, act=act)
            assert _rel(out_big[:, 4:], want) < 2e-6
            assert out_big[:, :4].abs().max().item() == 0.0        # nothing written outside the slice
    wide = torch.randn(B * n, 2 * D, generator=g).cuda()
    scale = torch.randn(D // heads, generator=g).cuda()
    got, want = torch.empty(B * n, D, device="cuda"), torch.empty(B * n, D, device="cuda")
    ops.pr_l2norm_scale(wide[:, D:], heads, scale, got)
    _TorchOps.pr_l2norm_scale(wide[:, D:], heads, scale, want)
    assert _rel(got, want) < 2e-6
    nk = 11
    q = torch.randn(B * n, D, generator=g).cuda()
    kv = torch.randn(B * nk, 2 * D, generator=g).cuda()
    got, want = torch.empty(B * n, D, device="cuda"), torch.empty(B * n, D, device="cuda")
    ops.pr_attention(q, kv[:, :D], kv[:, D:], B, heads, 8.0, got)
    _TorchOps.pr_attention(q, kv[:, :D].contiguous(), kv[:, D:].contiguous(), B, heads, 8.0, want)
    assert _rel(got, want) < 5e-6
    got = torch.empty(B, D, device="cuda")
    ops.pr_token_mean(x, got)
    assert _rel(got, x.mean(dim=1)) < 2e-6
    out_big = torch.zeros(B, n + 2, D, device="cuda")
    src = torch.randn(n, D, generator=g).cuda()
    ops.pr_broadcast_rows(src, out_big[:, 2:])
    assert torch.equal(out_big[:, 2:], src.expand(B, n, D)) and out_big[:, :2].abs().max().item() == 0.0
    a, b = torch.randn(B * n, D, generator=g).cuda(), torch.randn(B * n, D, generator=g).cuda()
    want = a + b
    ops.add_rows_(a, b)
    assert torch.equal(a, want)


@pytest.mark.parametrize("B,n", [(2, 12), (1, 8)])
def test_task_pool_on_kernels_matches_oracle(B, n):
    from v2a_b200 import unet as U
    torch.manual_seed(3)
    seq = torch.nn.Sequential(U.PerceiverResampler(dim=512, depth=2), torch.nn.Linear(512, 512))
    with torch.no_grad():
        for p in seq.parameters():
            p.add_(0.05 * torch.randn_like(p))
    y = torch.randn(B, n, 512)
    sd = {"pool.0." + k: v for k, v in seq[0].state_dict().items()}
    with torch.no_grad():
        want = F.linear(VO.perceiver_resampler(sd, "pool.0.", y), seq[1].weight, seq[1].bias).mean(dim=1)
    seq = seq.cuda()
    got = torch.empty(B, 512, device="cuda")
    n0 = U.ops.launch_count()
    with torch.no_grad():
        U._task_pool_cuda(seq, y.cuda(), got)
        stock = _task_pool(seq, y.cuda())
    launches = U.ops.launch_count() - n0
    assert _rel(got.cpu(), want) < 1e-5, _rel(got.cpu(), want)
    assert _rel(got, stock) < 1e-5
    assert launches == 36                                        # 15 per layer x 2 + 6: ran on the v2a kernels
