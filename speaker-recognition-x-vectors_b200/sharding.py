"""Utterance-sharded extraction over the GPUs of one box: one process per GPU (torch.distributed), a length-balanced
LPT partition, NO collective on the data path; the only exchange is the final gather of (N, 512) embeddings to rank 0.

The reference is single-GPU (main.py:220); utterances are independent in eval mode, so this is pure data parallelism.
`extract_fn(list_of_host_utterances) -> (n, D) float array` is the per-rank extractor (HostExtractor.extract_all on a
GPU rank); it is a parameter so that the host logic can be exercised with gloo on CPU.
"""
from __future__ import annotations

from typing import Callable, Sequence

import numpy as np
import torch
import torch.distributed as dist

from .layout import lpt_partition


def my_shard(lengths, rank: int, world: int) -> np.ndarray:
    """Original indices of the utterances this rank extracts (deterministic, identical on every rank)."""
    return lpt_partition(lengths, world)[rank]


def extract_sharded(utts: Sequence[torch.Tensor], extract_fn: Callable, dim: int | None = None, dst: int = 0):
    """Every rank holds (or can load) the same utterance list; each extracts its LPT shard; rank `dst` returns the
    float64 (N, D) matrix in the original order, the others return None."""
    world = dist.get_world_size() if dist.is_initialized() else 1
    rank = dist.get_rank() if dist.is_initialized() else 0
    lengths = np.asarray([int(u.shape[0]) for u in utts], dtype=np.int64)
    idx = my_shard(lengths, rank, world)
    local = np.asarray(extract_fn([utts[i] for i in idx]), dtype=np.float64) if len(idx) else np.zeros((0, dim or 0))
    if world == 1:
        out = np.empty((len(utts), local.shape[1]), dtype=np.float64)
        out[idx] = local
        return out
    gathered = [None] * world if rank == dst else None
    dist.gather_object((idx, local), gathered, dst=dst)  # host gather: ~10 MB for 4874 x 512 float64
    if rank != dst:
        return None
    d = next(g[1].shape[1] for g in gathered if g[1].shape[0])
    out = np.empty((len(utts), d), dtype=np.float64)
    seen = np.zeros(len(utts), dtype=bool)
    for gi, gx in gathered:
        out[gi] = gx
        seen[gi] = True
    if not seen.all():
        raise RuntimeError("sharded extraction lost utterances")
    return out
