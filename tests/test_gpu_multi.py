"""Multi-GPU parity (-m gpu, needs two GPUs): two ranks over NCCL, events dealt e mod 2, every rank cuts its share on its
own B200 through the C ABI, the fragments are gathered to rank 0 as one blob per rank (sharding.gather_blobs) and rank 0
checks EVERY event of the job against the oracle port, bit for bit.  Skipped on a one-GPU box."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

N_EVENTS = 6


def _worker(rank, world, port, q):
    import sys
    here = os.path.dirname(os.path.abspath(__file__))
    sys.path.insert(0, here)
    sys.path.insert(0, os.path.dirname(here))
    import torch.distributed as dist
    import common
    from oracle import portapi as P
    from surtr_b200 import FractureContext, sharding
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        mine = sharding.event_shard(N_EVENTS, world, rank)
        psets = [common.voronoi(500 + int(e), 120) for e in mine]
        csets = [common.voronoi(900 + int(e), 12 + int(e)) for e in mine]
        pieces, ev_p = common.concat(psets)
        cells, ev_c = common.concat(csets)
        ctx = FractureContext(rank)
        ctx.upload_pieces(pieces.verts, pieces.vert_off, pieces.ring_off, pieces.ring, ev_p)
        ctx.upload_cells(cells.planes, cells.plane_off, cells.verts, cells.vert_off, ev_c)
        ctx.fracture_event()
        c = ctx.counts()
        cap = 64 * c.n_fragments + 13 * c.n_verts + 2 * c.n_ring + 4 * 256
        blob = torch.zeros(cap + 256, dtype=torch.uint8, device=dev)
        hdr = 256                                        # first bytes: the layout, so that rank 0 can unpack the slice
        L = ctx.download_blob_into_async(blob.data_ptr() + hdr, cap)
        ctx.sync()
        lay = np.array([L.fragments, L.verts3, L.ring_len, L.ring, L.total, L.n_fragments, L.n_verts, L.n_ring, L.ring_entry_bytes], np.uint64)
        blob[:lay.nbytes].copy_(torch.from_numpy(lay.view(np.uint8).copy()))
        blob = blob[:hdr + int(L.total)]
        got = sharding.gather_blobs(blob, 0)
        if rank == 0:
            buf, off = got
            host = buf.cpu().numpy()
            from surtr_b200.engine import OutLayout
            order = sharding.merge_order(N_EVENTS, world)
            per_rank = []
            for r in range(world):
                sl = host[off[r]:off[r + 1]]
                lay = sl[:hdr].view(np.uint64)[:9]
                LL = OutLayout(*[int(x) for x in lay])
                per_rank.append(FractureContext.unpack_output_blob(sl[hdr:], LL))
            # event e is the k-th event of rank r: its fragments are the k-th cell range of that rank's batch
            failures = []
            for e, (r, k) in enumerate(order):
                want = P.apply_fracture(common.voronoi(500 + e, 120), common.voronoi(900 + e, 12 + e).planes, common.voronoi(900 + e, 12 + e).plane_off)
                fr = per_rank[r]
                shard = sharding.event_shard(N_EVENTS, world, r)
                c0 = sum(12 + int(x) for x in shard[:k])
                sel = np.nonzero((fr.rec["cell"] >= c0) & (fr.rec["cell"] < c0 + 12 + e))[0]
                ok = (len(sel) == want.n and np.array_equal(fr.rec["cell"][sel] - c0, want.cell) and
                      np.array_equal(fr.rec["piece"][sel] - 120 * k, want.piece) and np.array_equal(fr.rec["n_verts"][sel], want.nverts) and
                      np.array_equal(fr.rec["n_faces"][sel], want.nfaces) and fr.rec["volume"][sel].tobytes() == want.volume.tobytes())
                if ok and len(sel):
                    v0 = int(fr.rec["vert_off"][sel[0]])
                    ok = fr.verts[v0:v0 + len(want.verts)].tobytes() == want.verts.tobytes()
                if not ok:
                    failures.append(e)
            q.put(failures)
        dist.barrier()
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs (gpurun --gpus 2)")
def test_two_ranks_nccl_gather_matches_oracle():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29600 + (os.getpid() % 1000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    # a rank that dies leaves the queue empty: poll it and the processes, do not sit out a long timeout
    import queue, time
    failures, t_end = None, time.monotonic() + 240
    while failures is None and time.monotonic() < t_end:
        try:
            failures = q.get(timeout=2)
        except queue.Empty:
            if any(p.exitcode not in (None, 0) for p in procs):
                break
    for p in procs:
        p.join(timeout=60)
        if p.is_alive():
            p.terminate()
    assert failures is not None, f"a rank failed before the gather (exit codes {[p.exitcode for p in procs]})"
    assert all(p.exitcode == 0 for p in procs)
    assert failures == [], f"events that differ from the oracle after the gather: {failures}"
