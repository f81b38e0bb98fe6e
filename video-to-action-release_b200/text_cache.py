"""Task-text embedding cache (last item).

``Video_PredModel.sample`` (diffuser/models/video_model.py:49-70) runs the CLIP tokenizer + text encoder on the task
strings of every call, although the online trainer only ever asks for its 8 fixed Libero task descriptions
(lb_online_trainer_v7.py:880-891, one task per call).  The embeddings are constants of the task strings, so they are
computed once and kept on the GPU.

The cache key is the whole TUPLE of strings of a call, not the single string: the tokenizer pads a batch to its
longest member and the UNet's PerceiverResampler attends to every position (no mask, guided_diffusion/unet.py:671),
so a task's embedding rows depend on what it was batched with.  Returning the identical tensor object for a repeated
batch also lets the UNet engine skip its step-invariant ``task_attnpool`` evaluation (it keys on the tensor's
``data_ptr`` / version, unet.py ``set_task_embed``).

The CLIP encoder itself stays the reference's (out of scope); this only wraps the call.
"""
from __future__ import annotations

from collections import OrderedDict
from typing import Callable, Sequence

import torch


class TaskEmbedCache:
    """Memoises ``encode(batch_text) -> Tensor [B, L, D]`` per tuple of strings (LRU, ``max_entries`` batches)."""

    def __init__(self, encode: Callable[[Sequence[str]], torch.Tensor], max_entries: int = 256):
        self.encode = encode
        self.max_entries = max_entries
        self._store: "OrderedDict[tuple, torch.Tensor]" = OrderedDict()
        self.hits = 0
        self.misses = 0

    def __call__(self, batch_text: Sequence[str]) -> torch.Tensor:
        key = tuple(batch_text)
        hit = self._store.get(key)
        if hit is not None:
            self._store.move_to_end(key)
            self.hits += 1
            return hit
        self.misses += 1
        with torch.no_grad():
            out = self.encode(list(batch_text)).detach()
        self._store[key] = out
        while len(self._store) > self.max_entries:
            self._store.popitem(last=False)
        return out

    def clear(self) -> None:
        self._store.clear()


def install_text_cache(video_model, max_entries: int = 256) -> TaskEmbedCache:
    """Wrap ``video_model.encode_batch_text`` (a ``Video_PredModel``) with a ``TaskEmbedCache``; returns the cache.
    Call again after swapping the text encoder (or ``cache.clear()``)."""
    cache = TaskEmbedCache(video_model.encode_batch_text, max_entries)
    video_model.encode_batch_text = cache
    return cache
