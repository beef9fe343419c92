"""Utterance-sharded extraction over the GPUs of one box: one process per GPU (torch.distributed), a length-balanced
partition (shard_batches: whole batches, bit-identical results for any number of GPUs; lpt_partition: single utterances),
NO collective on the data path; the only exchange is the final gather of (N, 512) embeddings.

The reference is single-GPU (main.py:220, extraction loop main.py:135-146); utterances are independent in eval mode, so this
is pure data parallelism.  The gather moves fixed-shape float32 tensors, never pickled objects: every rank knows every shard's
size (the partition is a deterministic function of the lengths), pads its (n_r, D) block to the largest shard and takes part in
ONE all_gather_into_tensor — over NCCL that is device memory to device memory through NVLink / NVSwitch (c5: 8 x 610 x 512 x 4 B
= 10 MB in total), over gloo (the CPU tests) the same call on host tensors.  Every rank ends up with the whole matrix in the
original utterance order, so trial scoring can run on any of them without another hop.
"""
from __future__ import annotations

from typing import Callable, Sequence

import numpy as np
import torch
import torch.distributed as dist

from .layout import balanced_batches, lpt_partition


def _world():
    return (dist.get_world_size(), dist.get_rank()) if dist.is_available() and dist.is_initialized() else (1, 0)


def my_shard(lengths, rank: int, world: int) -> np.ndarray:
    """Original indices of the utterances this rank extracts (deterministic, identical on every rank)."""
    return lpt_partition(lengths, world)[rank]


def shard_batches(lengths, world: int, target_frames: int = 49152):
    """Batch-granular sharding: (parts, batch_sizes) with parts[r] = original indices of the utterances worker r extracts, in
    extraction order, and batch_sizes[r] = utterance counts of its batches.  The batches come from layout.balanced_batches —
    contiguous runs of the utterance list of nearly equal frames, their number a multiple of 8 — and worker r takes batches
    r, r + world, ...: which utterances share a batch never depends on `world`, so the x-vectors are bit-identical for 1, 2, 4
    or 8 GPUs (a batch's composition fixes the fp32 summation order of the statistics pooling)."""
    batches = balanced_batches(lengths, target_frames=target_frames, multiple_of=8)
    parts, sizes = [], []
    for r in range(world):
        mine = batches[r::world]
        parts.append(np.concatenate(mine) if mine else np.zeros(0, dtype=np.int64))
        sizes.append([len(b) for b in mine])
    return parts, sizes


class RowGather:
    """All-gather of per-rank row blocks into the original order, with everything that does not change between calls built once:
    the padded send buffer (a rank's extractor can write its x-vectors straight into `local_view`), the receive buffer and the
    index tensors of the final reorder.  __call__ is ONE fixed-shape all_gather_into_tensor + one indexed copy — no pickling, no
    host round trip, no per-call allocation.  Under NCCL the buffers are device memory (the collective runs GPU to GPU over
    NVLink / NVSwitch); under gloo (CPU tests) they are host tensors.

    parts[r]: original row indices of rank r's block, known to every rank (a deterministic function of the lengths)."""

    def __init__(self, parts: Sequence[np.ndarray], n_total: int, dim: int, device, dtype=torch.float32):
        self.world, self.rank = _world()
        if len(parts) != self.world:
            raise ValueError("one part per rank expected")
        dst = np.concatenate([np.asarray(p, dtype=np.int64) for p in parts])
        if dst.size != n_total or np.unique(dst).size != n_total or (dst.size and (dst.min() < 0 or dst.max() >= n_total)):
            raise RuntimeError("sharded extraction would lose or duplicate utterances")
        device = torch.device(device)
        if self.world > 1 and dist.get_backend() == "nccl" and device.type != "cuda":
            raise ValueError("under the NCCL backend the embeddings must be CUDA tensors (HostExtractor.extract_flat(to_host=False))")
        self.n_local, self.n_total = len(parts[self.rank]), n_total
        cap = max(max(len(p) for p in parts), 1)
        self.send = torch.zeros((cap, dim), dtype=dtype, device=device)
        self.local_view = self.send[: self.n_local]
        self.recv = self.send if self.world == 1 else torch.empty((self.world * cap, dim), dtype=dtype, device=device)
        # rows of rank r sit at [r*cap, r*cap + len(parts[r])) of recv
        src = np.concatenate([r * cap + np.arange(len(p), dtype=np.int64) for r, p in enumerate(parts)])
        self.src = torch.from_numpy(src).to(device)
        self.dst = torch.from_numpy(dst).to(device)
        self.out = torch.empty((n_total, dim), dtype=dtype, device=device)

    def __call__(self, local: torch.Tensor | None = None) -> torch.Tensor:
        """local: this rank's (n_local, dim) block, or None when it was written into `local_view` already.  Returns the
        (n_total, dim) matrix in the original order (a buffer owned by this object, overwritten by the next call), on EVERY rank."""
        if local is not None:
            if local.shape != self.local_view.shape:
                raise ValueError("local block does not match this rank's shard")
            self.local_view.copy_(local)
        if self.world > 1:
            dist.all_gather_into_tensor(self.recv, self.send)
        self.out.index_copy_(0, self.dst, self.recv.index_select(0, self.src))
        return self.out


def gather_rows(local: torch.Tensor, parts: Sequence[np.ndarray], n_total: int) -> torch.Tensor:
    """One-shot form of RowGather: local (len(parts[rank]), D) -> (n_total, D) on the same device, on every rank."""
    if local.dim() != 2:
        raise ValueError("local must be (n_local, D)")
    return RowGather(parts, n_total, local.shape[1], local.device, local.dtype)(local)


def extract_sharded(utts: Sequence[torch.Tensor], extract_fn: Callable, dim: int | None = None, dst: int = 0):
    """Every rank holds (or can load) the same utterance list; each extracts its LPT shard with
    `extract_fn(list_of_host_utterances) -> (n, D)` (numpy or torch, host or CUDA; HostExtractor.extract_all on a GPU rank —
    a parameter so that the host logic can be exercised with gloo on CPU).  Rank `dst` returns the float64 numpy (N, D) matrix in
    the original order (the dtype test_epoch_end stores, main.py:145), the others return None."""
    world, rank = _world()
    lengths = np.asarray([int(u.shape[0]) for u in utts], dtype=np.int64)
    parts = lpt_partition(lengths, world)
    idx = parts[rank]
    if len(idx):
        local = extract_fn([utts[i] for i in idx])
        local = local if isinstance(local, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(local))
    else:
        if dim is None:
            raise ValueError("a rank with an empty shard needs `dim`")
        local = torch.zeros((0, dim))
    local = local.to(torch.float32)
    if world > 1 and dist.get_backend() == "nccl" and not local.is_cuda:
        local = local.cuda()
    full = gather_rows(local.contiguous(), parts, len(utts))
    return full.double().cpu().numpy() if rank == dst else None


def extract_sharded_flat(hx, flat_host: torch.Tensor, lengths_shard, parts: Sequence[np.ndarray], n_total: int,
                         max_frames: int = 1 << 17, batch_sizes=None, gather: RowGather | None = None) -> torch.Tensor:
    """The GPU-native form: this rank's shard is already ONE flat pinned host tensor (sum(lengths_shard), C) in the order of
    parts[rank]; `hx` is this rank's HostExtractor.  The x-vectors never leave the devices: batches write into a (n_r, D) CUDA
    matrix, one NCCL all-gather, one indexed copy.  Returns the float32 CUDA (n_total, D) matrix in the original order, on every
    rank.  With (parts, batch_sizes) from shard_batches the result is bit-identical for any world size.  Pass a RowGather built
    once for (parts, n_total) to reuse its buffers across calls: the batches then write straight into its send buffer."""
    if gather is None:
        dim = (hx.model.segment_layer7 if hx.model.x_vec_extract_layer == 7 else hx.model.segment_layer6).out_features
        gather = RowGather(parts, n_total, dim, hx.device)
    if len(lengths_shard):  # a rank may own nothing (fewer batches than ranks): it still takes part in the collective
        hx.extract_flat(flat_host, lengths_shard, max_frames=max_frames, to_host=False, batch_sizes=batch_sizes, out_dev=gather.local_view)
    return gather()
