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


def gather_blobs(blob: torch.Tensor, dst: int = 0, group=None):
    """Final fragment gather, one blob per rank: `blob` = the rank's fragments as ONE contiguous uint8 tensor (the
    output blob of surtr_download_blob_async packed into device memory, or several of them back to back).  One
    all_gather of the sizes (one int64 per rank, the only host synchronisation), then ONE grouped send / recv
    (ncclGroupStart .. ncclGroupEnd through batch_isend_irecv): every rank sends its bytes straight into its slice of a
    single receive buffer on `dst` -- no padding, no per-array collectives.  Returns (buffer, offsets[world + 1]) on
    `dst`, None elsewhere.  Runs over gloo on CPU tensors in the tests."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    blob = blob.contiguous().view(torch.uint8).reshape(-1)
    sizes = torch.zeros(world, dtype=torch.int64, device=blob.device)
    mine = torch.tensor([blob.numel()], dtype=torch.int64, device=blob.device)
    dist.all_gather_into_tensor(sizes, mine, group=group) if blob.is_cuda else dist.all_gather(list(sizes.split(1)), mine, group=group)
    sizes = sizes.tolist()
    off = [0]
    for n in sizes:
        off.append(off[-1] + int(n))
    ops = []
    buf = None
    if rank == dst:
        buf = torch.empty(off[-1], dtype=torch.uint8, device=blob.device)
        buf[off[rank]:off[rank + 1]].copy_(blob)
        for r in range(world):
            if r != dst and sizes[r]:
                ops.append(dist.P2POp(dist.irecv, buf[off[r]:off[r + 1]], r, group))
    elif blob.numel():
        ops.append(dist.P2POp(dist.isend, blob, dst, group))
    if ops:
        for w in dist.batch_isend_irecv(ops):
            w.wait()
    return (buf, off) if rank == dst else None
