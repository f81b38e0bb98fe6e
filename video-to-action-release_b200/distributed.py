"""Multi-GPU plumbing: one process per GPU, ``torch.distributed`` (NCCL over NVLink 5 / NVSwitch).

Only two things cross ranks on the hot paths:

  * video sampling shards the batch of prompts — NO collective inside the denoise loop
    (GroupNorm / attention never cross the batch dimension, noise is per sample); an optional
    final all-gather returns every rank's videos;
  * the policy step has ONE exchange: the gradient all-reduce (mean) of the flat gradient
    slab between backward and the fused clip + AdamW + EMA kernel.

The reference itself is single process (scripts/train_libero_dp.sh:11-12; the Accelerator at
diffuser/libero/lb_online_trainer_v7.py:72-76 would wrap the policy in DDP under
``accelerate launch``) — this module is what replaces that DDP wrapper.

Host logic here is backend agnostic and is exercised on CPU with ``gloo`` at world_size 2
(tests/test_distributed_host.py).
"""
from __future__ import annotations

import os
from typing import List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist


def env_world() -> Tuple[int, int, int]:
    """(rank, local_rank, world_size) from the torchrun environment; (0, 0, 1) when absent."""
    return (int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0")),
            int(os.environ.get("WORLD_SIZE", "1")))


def init_from_env(backend: Optional[str] = None):
    """Initialise the default process group from RANK / WORLD_SIZE / MASTER_* (no-op for 1 rank).

    ``backend`` defaults to nccl when CUDA is present (binding this rank to ``cuda:LOCAL_RANK``),
    gloo otherwise.  Returns (rank, local_rank, world_size).
    """
    rank, local, world = env_world()
    if world == 1 or dist.is_initialized():
        if torch.cuda.is_available():
            torch.cuda.set_device(local)
        return rank, local, world
    if backend is None:
        backend = "nccl" if torch.cuda.is_available() else "gloo"
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    if backend == "nccl":
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    else:
        dist.init_process_group(backend)
    return rank, local, world


def world_size(group=None) -> int:
    return dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1


def shard_range(n: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous [lo, hi) of ``n`` items for ``rank``; the first ``n % world`` ranks get one extra."""
    if not (0 <= rank < world):
        raise ValueError(f"rank {rank} outside world {world}")
    base, extra = divmod(n, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_batch(tensors: Sequence[torch.Tensor], rank: int, world: int) -> List[torch.Tensor]:
    """Slice dim 0 of every tensor to this rank's shard (all must share dim 0)."""
    n = tensors[0].shape[0]
    for t in tensors:
        if t.shape[0] != n:
            raise ValueError("shard_batch: tensors disagree on the batch dimension")
    lo, hi = shard_range(n, rank, world)
    return [t[lo:hi] for t in tensors]


def bucket_ranges(numel: int, bucket_elems: int) -> List[Tuple[int, int]]:
    """[lo, hi) element windows of at most ``bucket_elems`` covering a flat slab."""
    if bucket_elems <= 0:
        raise ValueError("bucket_elems must be positive")
    return [(lo, min(lo + bucket_elems, numel)) for lo in range(0, numel, bucket_elems)]


def allreduce_mean_start(slabs: Sequence[torch.Tensor], group=None, bucket_bytes: int = 64 << 20) -> list:
    """Launch the in-place mean over ranks of flat fp32 gradient slabs and return the pending works (empty without
    a process group or with one rank).  The collectives run on the NCCL stream, ordered after what the current
    stream holds NOW — kernels enqueued afterwards overlap with them until ``allreduce_wait``."""
    if not dist.is_available() or not dist.is_initialized():
        return []
    world = dist.get_world_size(group)
    if world == 1:
        return []
    # NCCL averages inside the collective (ReduceOp.AVG): no extra read + write pass over the 349 MB of gradients.
    # gloo (the CPU tests) has no AVG: pre-divide, then SUM.
    avg = dist.get_backend(group) == "nccl"
    works = []
    for slab in slabs:
        flat = slab.view(-1)
        for lo, hi in bucket_ranges(flat.numel(), max(1, bucket_bytes // flat.element_size())):
            chunk = flat[lo:hi]
            if avg:
                works.append(dist.all_reduce(chunk, op=dist.ReduceOp.AVG, group=group, async_op=True))
            else:
                chunk.mul_(1.0 / world)
                works.append(dist.all_reduce(chunk, op=dist.ReduceOp.SUM, group=group, async_op=True))
    return works


def allreduce_wait(works: list) -> int:
    """Order the pending collectives before the caller's next kernel; returns how many there were."""
    for w in works:
        w.wait()
    return len(works)


def allreduce_mean_(slabs: Sequence[torch.Tensor], group=None, bucket_bytes: int = 64 << 20) -> int:
    """In-place mean over ranks of flat fp32 gradient slabs; returns the number of collectives issued.

    NVSwitch gives every peer full bandwidth and reduces in the switch (NVLS), so the bucket size is
    chosen for launch latency only: 64 MiB buckets -> 5 collectives for the 259 MB UNet1D slab.  The
    collectives are asynchronous on the NCCL stream and ordered before the caller's next kernel by
    ``wait()``; with one rank this is a no-op.
    """
    return allreduce_wait(allreduce_mean_start(slabs, group, bucket_bytes))


def gather_videos(local: torch.Tensor, group=None) -> torch.Tensor:
    """All-gather per-rank sample() outputs [b_r, ...] along dim 0 (equal b_r on every rank)."""
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return local
    world = dist.get_world_size(group)
    out = torch.empty((world * local.shape[0],) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, local.contiguous(), group=group)
    return out


def sample_sharded(diffusion, x_cond: torch.Tensor, task_embed: torch.Tensor, *, gather: bool = False,
                   seed: Optional[int] = None, group=None) -> torch.Tensor:
    """Data-parallel ``GoalGaussianDiffusion.sample`` over the prompts of a global batch.

    Every rank passes the SAME global ``x_cond`` / ``task_embed``; rank r samples rows
    ``shard_range(B, r, world)`` with its own generator seed (``seed + r`` when given) so the result
    equals a single-GPU run on that sub-batch with that seed.  No collective runs inside the loop.
    """
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    n = x_cond.shape[0]
    if gather and n % world != 0:
        # all_gather_into_tensor needs equal shards; uneven or empty ones would error or hang under NCCL
        raise ValueError(f"sample_sharded(gather=True): batch {n} is not a multiple of the world size {world}")
    xc, te = shard_batch([x_cond, task_embed], rank, world)
    if xc.shape[0] == 0:            # fewer prompts than ranks: this rank has nothing to sample
        H, W = diffusion.image_size
        return x_cond.new_zeros((0, diffusion.channels, H, W))
    if seed is not None:
        torch.manual_seed(seed + rank)
    out = diffusion.sample(xc, te, batch_size=xc.shape[0])
    return gather_videos(out, group) if gather else out
