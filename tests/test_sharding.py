"""Multi-GPU logic on CPU: world_size-2 gloo run of the event sharding + final fragment gather.
Each rank cuts its own events (here with the oracle port standing in for its GPU) and rank 0 reassembles the job;
the result must equal the single-process result event by event."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import common
from oracle import portapi as P
from surtr_b200 import sharding

N_EVENTS = 5


def _event(e):
    pieces = common.voronoi(100 + e, 40)
    cells = common.voronoi(200 + e, 8 + e)
    return P.apply_fracture(pieces, cells.planes, cells.plane_off)


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mine = sharding.event_shard(N_EVENTS, world, rank)
    frs = [_event(int(e)) for e in mine]
    counts = torch.tensor([f.n for f in frs], dtype=torch.int64)
    verts = torch.from_numpy(np.concatenate([f.verts for f in frs]).reshape(-1)) if frs else torch.zeros(0)
    vols = torch.from_numpy(np.concatenate([f.volume for f in frs])) if frs else torch.zeros(0, dtype=torch.float64)
    cells = torch.from_numpy(np.concatenate([f.cell for f in frs]).astype(np.int64)) if frs else torch.zeros(0, dtype=torch.int64)
    got = [sharding.gather_variable(t) for t in (counts, verts, vols, cells)]
    # the one-blob gather: everything the rank holds as one byte string
    blob = torch.cat([t.contiguous().view(torch.uint8).reshape(-1) for t in (counts, verts, vols, cells)])
    one = sharding.gather_blobs(blob)
    if rank == 0:
        buf, off = one
        q.put(([[x.numpy() for x in part] for part in got], buf.numpy(), off,
               [int(t.numel() * t.element_size()) for t in (counts, verts, vols, cells)]))
    dist.barrier()
    dist.destroy_process_group()


def test_shard_and_gather_world2():
    assert sharding.event_shard(7, 2, 1).tolist() == [1, 3, 5]
    assert sharding.merge_order(4, 2) == [(0, 0), (1, 0), (0, 1), (1, 1)]
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    (counts, verts, vols, cells), buf, off, sizes0 = q.get(timeout=180)
    # one-blob gather: rank r's slice is its four arrays back to back
    assert len(off) == world + 1 and off[-1] == len(buf)
    for r in range(world):
        want = np.concatenate([np.ascontiguousarray(a[r]).view(np.uint8).reshape(-1) for a in (counts, verts, vols, cells)])
        assert np.array_equal(buf[off[r]:off[r + 1]], want)
    assert off[1] == sum(sizes0)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    # reassemble in global event order and compare with the single-process results
    cursor_f = [0] * world
    cursor_v = [0] * world
    for e, (r, k) in enumerate(sharding.merge_order(N_EVENTS, world)):
        want = _event(e)
        n = int(counts[r][k])
        assert n == want.n
        f0 = cursor_f[r]
        assert np.array_equal(vols[r][f0:f0 + n], want.volume)
        assert np.array_equal(cells[r][f0:f0 + n], want.cell.astype(np.int64))
        nv = want.verts.size
        assert np.array_equal(verts[r][cursor_v[r]:cursor_v[r] + nv].view(np.uint32), want.verts.reshape(-1).view(np.uint32))
        cursor_f[r] += n
        cursor_v[r] += nv
