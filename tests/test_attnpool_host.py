"""Row V13 host logic (no GPU): the launch sequence of `_task_pool_cuda` (which buffer slice every kernel reads and
writes, operand order, residuals) evaluated with torch stand-ins for the C-ABI wrappers, against the oracle's
`perceiver_resampler` restatement.  The kernels themselves are checked on the GPU (tests/test_attnpool_gpu.py)."""
import pytest
import torch

from tests.stock_twins import _task_pool
import torch.nn.functional as F

from oracle import video_oracle as VO


class _TorchOps:
    """What each `ops.*` wrapper used by `_task_pool_cuda` is specified to compute (include/v2a_b200.h)."""
    ACT_NONE, ACT_GELU = 0, 3

    @staticmethod
    def pr_broadcast_rows(src, out):
        out.copy_(src.unsqueeze(0).expand_as(out))

    @staticmethod
    def pr_token_mean(x, out):
        out.copy_(x.mean(dim=1))

    @staticmethod
    def pr_layernorm(x, gamma, beta, out, *, pos=None, act=0, eps=1e-5):
        v = F.gelu(x) if act == 3 else x
        if pos is not None:
            v = v + pos[:x.shape[1]]
        out.copy_(F.layer_norm(v, (x.shape[-1],), gamma, beta, eps))

    @staticmethod
    def linear(x, W, bias, y, *, add=None, act_in=0, act_out=0):
        r = F.linear(x, W, bias)
        if add is not None:
            r = r + add
        y[:, :W.shape[0]].copy_(r)

    @staticmethod
    def pr_l2norm_scale(x, heads, scale, out):
        rows, width = x.shape
        out.copy_((F.normalize(x.reshape(rows, heads, -1), dim=-1) * scale).reshape(rows, width))

    @staticmethod
    def pr_attention(q, k, v, B, heads, scale, out):
        sp = lambda t: t.reshape(B, t.shape[0] // B, heads, -1).permute(0, 2, 1, 3)
        att = (torch.einsum("bhid,bhjd->bhij", sp(q), sp(k)) * scale).softmax(dim=-1)
        o = torch.einsum("bhij,bhjd->bhid", att, sp(v)).permute(0, 2, 1, 3)
        out.copy_(o.reshape(out.shape))

    @staticmethod
    def add_rows_(dst, src):
        dst.add_(src)


@pytest.mark.parametrize("B,n", [(2, 6), (1, 12), (3, 1)])
def test_task_pool_cuda_launch_sequence_matches_oracle(monkeypatch, B, n):
    from v2a_b200 import unet as U
    torch.manual_seed(0)
    seq = torch.nn.Sequential(U.PerceiverResampler(dim=64, depth=2, dim_head=16, heads=4, num_latents=8,
                                                   num_latents_mean_pooled=2, max_seq_len=32),
                              torch.nn.Linear(64, 40))     # time_embed_dim != token width (tiny UNet configs)
    with torch.no_grad():
        for p in seq.parameters():                       # gains 1 / biases 0 / scales 1 would hide operand mix-ups
            p.add_(0.3 * torch.randn_like(p))
    y = torch.randn(B, n, 64)
    sd = {"pool.0." + k: v for k, v in seq[0].state_dict().items()}
    want = F.linear(VO.perceiver_resampler(sd, "pool.0.", y, heads=4), seq[1].weight, seq[1].bias).mean(dim=1)
    for name in ("pr_broadcast_rows", "pr_token_mean", "pr_layernorm", "linear", "pr_l2norm_scale", "pr_attention",
                 "add_rows_"):
        monkeypatch.setattr(U.ops, name, getattr(_TorchOps, name))
    got = torch.empty(B, 40)
    with torch.no_grad():
        U._task_pool_cuda(seq, y, got)
        ref_torch = _task_pool(seq, y)
    rel = lambda a, b: ((a - b).norm() / b.norm()).item()
    assert rel(got, want) < 5e-6, rel(got, want)      # fp32 re-association only (mean before the last Linear)
    assert rel(ref_torch, want) < 5e-6


def test_task_pool_cuda_chunks_large_batches(monkeypatch):
    """B x (n + latents) token rows above `_POOL_MAX_ROWS` are processed in sample chunks (samples are independent)."""
    from v2a_b200 import unet as U
    torch.manual_seed(2)
    seq = torch.nn.Sequential(U.PerceiverResampler(dim=32, depth=1, dim_head=8, heads=2, num_latents=6,
                                                   num_latents_mean_pooled=2, max_seq_len=16),
                              torch.nn.Linear(32, 24))
    y = torch.randn(7, 4, 32)
    rows = []
    for name in ("pr_broadcast_rows", "pr_token_mean", "pr_layernorm", "linear", "pr_l2norm_scale", "pr_attention",
                 "add_rows_"):
        monkeypatch.setattr(U.ops, name, getattr(_TorchOps, name))
    lin = U.ops.linear
    monkeypatch.setattr(U.ops, "linear", lambda x, *a, **k: (rows.append(x.shape[0]), lin(x, *a, **k))[1])
    monkeypatch.setattr(U, "_POOL_MAX_ROWS", 40)            # 12 token rows per sample -> 3 samples per call
    got = torch.empty(7, 24)
    with torch.no_grad():
        U._task_pool_cuda(seq, y, got)
        want = _task_pool(seq, y)
    assert ((got - want).norm() / want.norm()).item() < 5e-6
    assert max(rows) <= 40 and len(rows) == 3 * 7            # 3 chunks (3 + 3 + 1 samples) x 7 linears each


def test_task_pool_cuda_without_mean_pooled_latents(monkeypatch):
    from v2a_b200 import unet as U
    torch.manual_seed(1)
    seq = torch.nn.Sequential(U.PerceiverResampler(dim=32, depth=1, dim_head=8, heads=2, num_latents=5,
                                                   num_latents_mean_pooled=0, max_seq_len=16),
                              torch.nn.Linear(32, 32))
    y = torch.randn(2, 4, 32)
    for name in ("pr_broadcast_rows", "pr_token_mean", "pr_layernorm", "linear", "pr_l2norm_scale", "pr_attention",
                 "add_rows_"):
        monkeypatch.setattr(U.ops, name, getattr(_TorchOps, name))
    got = torch.empty(2, 32)
    with torch.no_grad():
        U._task_pool_cuda(seq, y, got)
        want = _task_pool(seq, y)
    assert ((got - want).norm() / want.norm()).item() < 5e-6
