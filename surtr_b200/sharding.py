"""Multi-GPU partitioning of the fracture path: independent events / objects are dealt round-robin to the ranks
(event e -> rank e mod N), every rank cuts its own events with NO collective on the hot
path, and one gather of the variable-length fragment arrays to rank 0 closes the job (NCCL on GPUs; the same code
runs over gloo on CPU tensors in the tests)."""
from __future__ import annotations

import numpy as np
import torch
import torch.distributed as dist


def event_shard(n_events: int, world: int, rank: int) -> np.ndarray:
    """Events owned by `rank`: e mod world == rank, ascending."""
    return np.arange(rank, n_events, world, dtype=np.int64)


def merge_order(n_events: int, world: int):
    """For rank-0 reassembly: (rank, local index) of every event in global event order."""
    return [(e % world, e // world) for e in range(n_events)]


def gather_variable(t: torch.Tensor, dst: int = 0, group=None):
    """Gather 1-D tensors of different lengths to `dst`: all_gather of the lengths, then a padded gather.
    Returns the list of per-rank tensors on `dst`, None elsewhere."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    dtype = t.dtype
    t = t.contiguous().view(torch.uint8)      # ship raw bytes: NCCL has no 16-bit integer type
    n = torch.tensor([t.numel()], dtype=torch.int64, device=t.device)
    counts = [torch.zeros_like(n) for _ in range(world)]
    dist.all_gather(counts, n, group=group)
    counts = [int(c.item()) for c in counts]
    m = max(counts) if counts else 0
    pad = torch.zeros(m, dtype=t.dtype, device=t.device)
    pad[:t.numel()] = t
    bufs = [torch.empty_like(pad) for _ in range(world)] if rank == dst else None
    dist.gather(pad, bufs, dst=dst, group=group)
    if rank != dst:
        return None
    return [b[:c].view(dtype) for b, c in zip(bufs, counts)]


def gather_fragments(rec_bytes: torch.Tensor, verts: torch.Tensor, ring_off: torch.Tensor, ring: torch.Tensor, dst: int = 0,
                     group=None):
    """Final fragment gather: records (as bytes), vertices, per-vertex ring offsets and ring entries of every rank."""
    parts = [gather_variable(x.reshape(-1), dst, group) for x in (rec_bytes, verts, ring_off, ring)]
    if parts[0] is None:
        return None
    return list(zip(*parts))
