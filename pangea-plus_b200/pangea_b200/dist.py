"""Multi-GPU plumbing (one process per GPU, torch.distributed): the path shards by read, so the
only exchanges are a one-off broadcast of the model's integer training counts and a gather of
the 64-byte result records in rank order.  Backend-agnostic: NCCL on GPUs,
gloo in the CPU tests."""
from __future__ import annotations

import numpy as np


def shard_range(n: int, rank: int, world: int) -> tuple[int, int]:
    """contiguous read range of `rank`: [rank*n/world, (rank+1)*n/world) -- concatenating the
    ranks' results in rank order restores input order, which Stage C requires (row C2)."""
    return (rank * n) // world, ((rank + 1) * n) // world


def broadcast_buffers(tensors, src: int = 0) -> None:
    """one-off replication of the model: rank `src` trained, everybody else receives the raw
    count buffers and derives the fp32 table locally (bit-identical, no table broadcast)."""
    import torch.distributed as dist

    for t in tensors:
        dist.broadcast(t, src=src)


def allreduce_counts(tensors) -> None:
    """sharded training (BASELINE configs[3]): every rank counted its own slice of the training set
    into a model of its own; an integer SUM all-reduce of the count buffers (pg_model_buffers: m, n, M as int32, N as
    int64) leaves every rank with the counts of the whole set -- integer adds commute, so the result is bit for bit
    the one-rank result whatever the sharding.  `tensors`: uint8 views of the buffers, in pg_model_buffers() order."""
    import torch
    import torch.distributed as dist

    for i, t in enumerate(tensors):
        v = t.view(torch.int64 if i == 3 else torch.int32)
        dist.all_reduce(v, op=dist.ReduceOp.SUM)


def gather_records(local, n_total: int, rank: int, world: int, dst: int = 0):
    """gather per-rank result records (uint8 tensors of 64*count bytes) on `dst` in rank order.
    Shards may differ by one read, so every rank pads to the largest shard."""
    import torch
    import torch.distributed as dist

    sizes = [shard_range(n_total, r, world) for r in range(world)]
    longest = max(b - a for a, b in sizes) * 64
    padded = torch.zeros(longest, dtype=torch.uint8, device=local.device)
    padded[: local.numel()] = local
    bufs = [torch.empty_like(padded) for _ in range(world)] if rank == dst else None
    dist.gather(padded, bufs, dst=dst)
    if rank != dst:
        return None
    return torch.cat([b[: (hi - lo) * 64] for b, (lo, hi) in zip(bufs, sizes)])


class DeviceBuffer:
    """zero-copy torch view of a raw device pointer (pg_model_buffers) via __cuda_array_interface__"""

    def __init__(self, ptr: int, nbytes: int):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 2}


def records_checksum(records: np.ndarray) -> int:
    """order-sensitive checksum of result records (genus, votes) for cross-rank comparisons"""
    g = records["genus"].astype(np.int64)
    v = records["votes"].astype(np.int64).sum(axis=1)
    idx = np.arange(1, len(records) + 1, dtype=np.int64)
    return int(((g * 1000003 + v) * idx % 2147483647).sum() % 2147483647)
