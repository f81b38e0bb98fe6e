"""Host-side multi-GPU logic on CPU: shard arithmetic, bucket windows, and a world_size-2 gloo
run of the gradient all-reduce (mean) and the prompt-sharded sampler wrapper."""
import os
import socket
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from v2a_b200 import distributed as D
from v2a_b200.train_step import ema_decay


def test_shard_range_covers_everything_once():
    for n in (0, 1, 7, 16, 128, 129):
        for world in (1, 2, 3, 8):
            spans = [D.shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            for (a, b), (c, d) in zip(spans, spans[1:]):
                assert b == c and b - a >= d - c >= 0
            assert max(b - a for a, b in spans) - min(b - a for a, b in spans) <= 1
    with pytest.raises(ValueError):
        D.shard_range(4, 2, 2)


def test_bucket_ranges():
    assert D.bucket_ranges(10, 4) == [(0, 4), (4, 8), (8, 10)]
    assert D.bucket_ranges(0, 4) == []
    assert D.bucket_ranges(4, 4) == [(0, 4)]
    with pytest.raises(ValueError):
        D.bucket_ranges(4, 0)


def test_shard_batch_and_single_rank_noops():
    a, b = torch.arange(10)[:, None], torch.arange(10) * 2
    sa, sb = D.shard_batch([a, b], 1, 3)
    assert sa[:, 0].tolist() == [4, 5, 6] and sb.tolist() == [8, 10, 12]
    with pytest.raises(ValueError):
        D.shard_batch([a, b[:5]], 0, 2)
    g = torch.ones(8)
    assert D.allreduce_mean_([g]) == 0 and torch.equal(g, torch.ones(8))   # no process group: untouched
    assert D.gather_videos(g) is g


def test_ema_decay_schedule_matches_ema_pytorch_0_2_3():
    # first two update() calls copy the online weights (decay 0), then 1 - (1 + c)^-0.75 capped at beta
    assert ema_decay(0) == 0.0 and ema_decay(1) == 0.0
    assert abs(ema_decay(2) - (1 - 3 ** -0.75)) < 1e-12
    assert abs(ema_decay(99) - (1 - 100 ** -0.75)) < 1e-12
    assert ema_decay(10 ** 9) == 0.9999


class _FakeDiffusion:
    """Stands in for GoalGaussianDiffusion on CPU: 'samples' = cond + per-rank seeded noise."""

    image_size, channels = (1, 4), 1

    def sample(self, x_cond, task_embed, batch_size):
        assert x_cond.shape[0] == batch_size == task_embed.shape[0]
        return x_cond + torch.randn(x_cond.shape)


def _worker(rank, world, port, tmp):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    r, _, w = D.init_from_env("gloo")
    assert (r, w) == (rank, world)
    # gradient slabs: rank r holds (r + 1) everywhere -> mean 1.5; odd sizes exercise the last bucket
    g1, g2 = torch.full((1000,), float(rank + 1)), torch.full((37,), float(10 * (rank + 1)))
    n = D.allreduce_mean_([g1, g2], bucket_bytes=4 * 256)
    assert n == 4 + 1
    assert torch.allclose(g1, torch.full((1000,), 1.5)) and torch.allclose(g2, torch.full((37,), 15.0))
    # split exchange (V2A_OVERLAP_ALLREDUCE): one slab's collectives start early, the rest later, one wait for all
    e1, e2 = torch.full((600,), float(rank + 1)), torch.full((5,), float(3 * (rank + 1)))
    early = D.allreduce_mean_start([e1], bucket_bytes=4 * 256)
    late = D.allreduce_mean_start([e2], bucket_bytes=4 * 256)
    assert D.allreduce_wait(early + late) == 3 + 1
    assert torch.allclose(e1, torch.full((600,), 1.5)) and torch.allclose(e2, torch.full((5,), 4.5))
    # prompt sharding: 5 prompts over 2 ranks -> 3 + 2, per-rank seeds, no collective in the loop
    cond = torch.arange(5.0)[:, None].expand(5, 4).contiguous()
    te = torch.zeros(5, 2, 8)
    out = D.sample_sharded(_FakeDiffusion(), cond, te, seed=100)
    lo, hi = D.shard_range(5, rank, world)
    torch.manual_seed(100 + rank)
    assert torch.equal(out, cond[lo:hi] + torch.randn(hi - lo, 4))
    # equal shards can be gathered
    out2 = D.sample_sharded(_FakeDiffusion(), cond[:4], te[:4], seed=7, gather=True)
    assert out2.shape == (4, 4)
    torch.manual_seed(7 + 0)
    assert torch.equal(out2[:2], cond[:2] + torch.randn(2, 4))
    # uneven shards cannot be all-gathered (would error or hang under NCCL) -> refused up front
    with pytest.raises(ValueError, match="multiple of the world size"):
        D.sample_sharded(_FakeDiffusion(), cond[:3], te[:3], gather=True)
    # fewer prompts than ranks: the empty shard returns an empty batch instead of calling sample(batch_size=0)
    one = D.sample_sharded(_FakeDiffusion(), cond[:1, None], te[:1], seed=3)
    assert one.shape[0] == (1 if rank == 0 else 0)
    torch.save(out2, os.path.join(tmp, f"g{rank}.pt"))
    dist.barrier()
    dist.destroy_process_group()


def test_gloo_world2_allreduce_and_sharded_sampling(tmp_path):
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    a, b = torch.load(tmp_path / "g0.pt"), torch.load(tmp_path / "g1.pt")
    assert torch.equal(a, b)   # every rank sees the same gathered batch
