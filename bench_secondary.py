"""bench_secondary.py -- the parts of bench.py that are not the headline timed region: the end-to-end loop of the
headline workload, BASELINE's other configs (2, 3, 5) as secondary results, the host<->device DMA ceiling of the box,
and the ncu traffic figure of the dominant kernel.  GPU arm only; nothing here touches oracle/."""
from __future__ import annotations

import hashlib
import json
import os
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
PIECES_PER_EVENT = 1000
CELLS_PER_EVENT = 64
FLUSH_MIB = 160
L2_MIB = 126


def _pinned(torch, a: np.ndarray):
    return torch.from_numpy(np.ascontiguousarray(a)).pin_memory()


# ------------------------------------------------------------------------------------------------ headline e2e
def config4_e2e(args, torch, dev, local, rank, world, streams, batch_views, res, n_my, barrier, allreduce, dist):
    """The whole job through the C ABI with HOST buffers: every pass uploads every event of the rank from pinned host
    memory and downloads every fragment into pinned host memory, ONE blob per direction and batch (surtr_upload_blob /
    surtr_download_blob_async: float3 wire format; copy rates on these boxes depend strongly on the copy size).
    Batches of --e2e-batch events rotate over --e2e-contexts contexts (own stream + copy stream each).  The kernels of
    consecutive batches are ordered by a CUDA event between the contexts' streams (batch i+1's kernels wait for batch
    i's), its upload is not: so the upload of batch i+1 and the download of batch i-1 run under the kernels of batch i
    instead of all contexts uploading, computing and downloading in a convoy.  ONE host thread drives it; it blocks
    only in surtr_download_blob_async, on the event whose fragments it is about to copy."""
    from surtr_b200 import FractureContext, FRAGMENT_DTYPE
    nctx = max(1, args.e2e_contexts)
    est = [torch.cuda.Stream(device=dev) for _ in range(nctx)]
    ctxs = [FractureContext(local, s.cuda_stream) for s in est]
    for cx in ctxs:
        cx.set_kdop_directions(args.kdop)
    ev_frags = np.concatenate([b.ev_frags for b in res])
    ev_verts = np.concatenate([b.ev_verts for b in res])
    ev_ring = np.concatenate([b.ev_ring for b in res])
    al = lambda x: (int(x) + 255) // 256 * 256

    class B:
        pass

    batches, h2d, d2h = [], 0, 0
    for e0 in range(0, n_my, args.e2e_batch):
        b = B()
        b.e0, b.e1 = e0, min(n_my, e0 + args.e2e_batch)
        p, c, evp, evc = batch_views(b.e0, b.e1)
        b.sizes, total = FractureContext.fill_input_blob(None, p, c, evp, evc)
        b.h_in = torch.empty(total, dtype=torch.uint8, pin_memory=True)
        FractureContext.fill_input_blob(b.h_in.numpy(), p, c, evp, evc)
        b.nf, b.nv, b.nr = int(ev_frags[b.e0:b.e1].sum()), int(ev_verts[b.e0:b.e1].sum()), int(ev_ring[b.e0:b.e1].sum())
        b.cap = al(FRAGMENT_DTYPE.itemsize * b.nf) + al(12 * b.nv) + al(b.nv) + al(2 * b.nr)
        b.h_out = torch.empty(b.cap, dtype=torch.uint8, pin_memory=True)
        h2d += total
        batches.append(b)

    def download(cx, b):
        b.L = cx.download_blob_into_async(b.h_out.data_ptr(), b.cap)

    def passes(k):
        pending = [None] * nctx
        prev_done = None
        for i in range(k * len(batches)):
            b = batches[i % len(batches)]
            s = i % nctx
            if pending[s] is not None:
                download(ctxs[s], pending[s])
            ctxs[s].upload_blob_ptr(b.h_in.data_ptr(), b.sizes)
            if prev_done is not None:
                est[s].wait_event(prev_done)
            ctxs[s].fracture_event()
            prev_done = torch.cuda.Event()
            prev_done.record(est[s])
            pending[s] = b
        for s in range(nctx):
            if pending[s] is not None:
                download(ctxs[s], pending[s])
        for cx in ctxs:
            cx.sync()

    passes(1)                    # warm-up: every context grows its buffers once
    d2h = sum(int(b.L.total) for b in batches)   # the bytes surtr_download_blob_async actually copies per pass
    # single synchronous batch: what one blocking caller sees
    t0 = time.perf_counter()
    ctxs[0].upload_blob_ptr(batches[0].h_in.data_ptr(), batches[0].sizes); ctxs[0].fracture_event(); download(ctxs[0], batches[0]); ctxs[0].sync()
    sync_batch_s = time.perf_counter() - t0
    barrier()
    t0 = time.perf_counter()
    passes(args.steps)
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0

    # ---- the end-to-end fragments are the resident-input fragments, bit for bit ----
    bi = 0
    for rb in res:
        fr = rb.cx.download()
        v3 = np.ascontiguousarray(fr.verts[:, :3])
        rl = np.diff(fr.ring_off).astype(np.uint8)
        fo = vo = ro = 0
        while bi < len(batches) and batches[bi].e1 <= rb.e1:
            b = batches[bi]
            L, out = b.L, b.h_out.numpy()
            assert (int(L.n_fragments), int(L.n_verts), int(L.n_ring)) == (b.nf, b.nv, b.nr), "e2e fragment counts differ"
            got = np.frombuffer(out[L.fragments:L.fragments + 64 * b.nf].tobytes(), dtype=FRAGMENT_DTYPE)
            want = fr.rec[fo:fo + b.nf]
            for f in ("n_verts", "n_faces", "volume", "centroid", "inertia", "n_ring"):
                assert got[f].tobytes() == want[f].tobytes(), f"e2e fragment records differ from the resident-input result ({f})"
            assert np.array_equal(got["piece"], want["piece"] - np.uint32((b.e0 - rb.e0) * PIECES_PER_EVENT)), "e2e piece ids differ"
            assert np.array_equal(got["cell"], want["cell"] - np.uint32((b.e0 - rb.e0) * CELLS_PER_EVENT)), "e2e cell ids differ"
            assert out[L.verts3:L.verts3 + 12 * b.nv].tobytes() == v3[vo:vo + b.nv].tobytes(), "e2e vertex positions differ"
            assert np.array_equal(out[L.ring_len:L.ring_len + b.nv], rl[vo:vo + b.nv]), "e2e ring lengths differ"
            rbytes = int(L.ring_entry_bytes)
            got_ring = out[L.ring:L.ring + rbytes * b.nr].view(np.uint8 if rbytes == 1 else np.uint16)
            assert np.array_equal(got_ring, fr.ring[ro:ro + b.nr]), "e2e ring entries differ"
            fo, vo, ro = fo + b.nf, vo + b.nv, ro + b.nr
            bi += 1
        assert fo == fr.n
        del fr, v3, rl
    assert bi == len(batches)
    for cx in ctxs:
        cx.close()

    e2e_max = allreduce(e2e_s, dist.ReduceOp.MAX)
    frags_job = allreduce(int(ev_frags.sum()), dist.ReduceOp.SUM)
    h2d_job = allreduce(h2d, dist.ReduceOp.SUM)
    d2h_job = allreduce(d2h, dist.ReduceOp.SUM)
    ms_step = 1e3 * e2e_max / args.steps
    return {"value": frags_job * args.steps / e2e_max, "unit": "fragments/s",
            "h2d_bytes_per_step": int(h2d_job), "d2h_bytes_per_step": int(d2h_job), "ms_per_step": ms_step,
            "achieved_gbs": {"h2d": h2d_job / (ms_step * 1e-3) / 1e9, "d2h": d2h_job / (ms_step * 1e-3) / 1e9,
                             "both": (h2d_job + d2h_job) / (ms_step * 1e-3) / 1e9},
            "batch_events": args.e2e_batch, "contexts_in_flight": nctx, "batches_per_rank": len(batches),
            "single_sync_batch_ms": 1e3 * sync_batch_s,
            "checked": "every fragment of the last pass (records, float3 positions, ring lengths, ring entries) equals the resident-input result bit for bit",
            "wire_format": "one blob per direction and batch, compact: surtr_upload_blob (float3 positions, one ring-length byte per vertex, one-byte ring "
                           "entries, 32-bit offsets per piece / cell only; expanded to the resident float4 / 32-bit / 16-bit arrays by one kernel) and "
                           "surtr_download_blob_async (64-byte records, float3 positions, one ring-length byte per vertex, one-byte ring entries, "
                           "assembled on the device)",
            "timing": f"wall clock (barrier + synchronize on both sides, max over ranks) around K passes of upload + event + download of every "
                      f"event of the rank, pinned host buffers, batches of {args.e2e_batch} events over {nctx} contexts from one host thread, "
                      "kernels of consecutive batches ordered by a CUDA event"}


# ------------------------------------------------------------------------------------------------ final gather
def final_gather(torch, dev, rank, world, res, barrier, allreduce, dist):
    """The only collective of the job (north star: NCCL over NVLink only for the final fragment gather): every rank packs
    the fragments of its resident batches into ONE device blob (surtr_download_blob_async with a device destination:
    records | float3 positions | ring lengths | ring entries per batch, back to back) and rank 0 receives all of them with
    one all_gather of the sizes + one grouped send / recv (sharding.gather_blobs).  Timed with CUDA events on rank 0,
    after a warm-up gather; checked by a device-side checksum of every rank's bytes."""
    from surtr_b200 import sharding
    al = lambda x: (int(x) + 255) // 256 * 256
    caps = []
    for b in res:
        c = b.cx.counts()
        caps.append(al(64 * int(c.n_fragments)) + al(12 * int(c.n_verts)) + al(int(c.n_verts)) + al(2 * int(c.n_ring)))
    blob = torch.zeros(sum(caps) + 8, dtype=torch.uint8, device=dev)
    at, frags = 0, 0
    for b, cap in zip(res, caps):
        L = b.cx.download_blob_into_async(blob.data_ptr() + at, cap)
        frags += int(L.n_fragments)
        at += int(L.total)
    for b in res:
        b.cx.sync()
    blob = blob[:at]
    check = lambda t: int(t[:t.numel() // 8 * 8].view(torch.int64).sum().item())
    mine = torch.tensor([check(blob)], dtype=torch.int64, device=dev)
    sums = torch.zeros(world, dtype=torch.int64, device=dev)
    dist.all_gather_into_tensor(sums, mine)
    sharding.gather_blobs(blob, 0)          # NCCL warm-up (connections, buffers)
    barrier()
    g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    g0.record()
    got = sharding.gather_blobs(blob, 0)
    g1.record()
    torch.cuda.synchronize()
    total_frags = allreduce(frags, dist.ReduceOp.SUM)
    out = None
    if rank == 0:
        buf, off = got
        for r in range(world):
            assert check(buf[off[r]:off[r + 1]]) == int(sums[r].item()), f"gathered bytes of rank {r} differ from what it sent"
        ms = g0.elapsed_time(g1)
        out = {"ms": ms, "bytes_total": int(off[-1]), "bytes_received": int(off[-1] - off[1]), "fragments_gathered": int(total_frags),
               "gbs_into_rank0": (off[-1] - off[1]) / (ms * 1e-3) / 1e9,
               "backend": "nccl: all_gather(sizes) + one grouped send/recv (batch_isend_irecv), one contiguous blob per rank, no padding",
               "checked": "int64 checksum of every rank's blob, computed by the sender and on the received slice"}
    barrier()
    return out


# ------------------------------------------------------------------------------------------------ DMA ceiling
def dma_ceiling(torch, dev, world, barrier, allreduce, dist, mib: int = 256, reps: int = 4):
    """Host<->device copy bandwidth of the box with every rank copying at once: pinned 256 MiB buffers, one direction at
    a time and both together on two streams.  The end-to-end numbers cannot exceed it."""
    n = mib * 2 ** 20
    h_a = torch.empty(n, dtype=torch.uint8, pin_memory=True)
    h_b = torch.empty(n, dtype=torch.uint8, pin_memory=True)
    d_a = torch.empty(n, dtype=torch.uint8, device=dev)
    d_b = torch.empty(n, dtype=torch.uint8, device=dev)
    s1, s2 = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)

    def timed(up, down):
        barrier()
        t0 = time.perf_counter()
        for _ in range(reps):
            if up:
                with torch.cuda.stream(s1):
                    d_a.copy_(h_a, non_blocking=True)
            if down:
                with torch.cuda.stream(s2):
                    h_b.copy_(d_b, non_blocking=True)
        torch.cuda.synchronize()
        dt = allreduce(time.perf_counter() - t0, dist.ReduceOp.MAX)
        return world * reps * n * (int(up) + int(down)) / dt / 1e9

    timed(True, True)
    out = {"h2d_gbs_all_gpus": timed(True, False), "d2h_gbs_all_gpus": timed(False, True), "duplex_gbs_all_gpus": timed(True, True),
           "n_gpus": world, "how": f"{reps} x {mib} MiB pinned copies per direction per GPU, all ranks at once, wall clock, max over ranks"}
    del h_a, h_b, d_a, d_b
    return out


# ------------------------------------------------------------------------------------------------ ncu traffic
def kernel_source_sha() -> str:
    h = hashlib.sha1()
    for f in ("clip_fast.cuh", "clip_sub.cuh", "clip_warp.cuh", "kernels.cuh", "surtr_math.cuh"):
        h.update(open(os.path.join(ROOT, "surtr_b200", "csrc", f), "rb").read())
    return h.hexdigest()[:16]


def k3_traffic_from_profile(alg_bytes_rank: float):
    """dram__bytes_read.sum + dram__bytes_write.sum of the K3 small-tier launch from the committed ncu capture of the
    SAME workload (profiles/r2_k3_traffic.json, written by tools/ncu_traffic.py from an `ncu --set full` run), used only
    if the capture was taken from the kernel source that is compiled now; otherwise null."""
    try:
        prof = json.load(open(os.path.join(ROOT, "profiles", "r2_k3_traffic.json")))
    except Exception:
        return None
    if prof.get("kernel_source_sha") != kernel_source_sha():
        return None
    return {"bytes_per_launch": prof.get("dram_bytes_per_launch"),
            "note": f"ncu --set full, {prof.get('workload')}, {prof.get('launches')} launch(es); algorithmic bytes of that launch "
                    f"{prof.get('algorithmic_bytes_per_launch')}"}


# ------------------------------------------------------------------------------------------------ configs 2, 3, 5
def run_all(args, torch, dev, local, rank, world, streams, barrier, allreduce, dist):
    from surtr_b200 import FractureContext, synth
    out = {}
    flush = torch.empty(FLUSH_MIB * 2 ** 20, dtype=torch.uint8, device=dev)
    ctx = FractureContext(local, streams[0].cuda_stream)
    ctx.set_kdop_directions(args.kdop)
    if rank == 0:
        out["fp32_peak_tflops"] = ctx.measure_fp32_peak()
        out["config3"] = config3(args, torch, ctx, synth, flush, streams[0])
        if world == 1:
            out["config2"] = config2(args, torch, dev, local, ctx, synth, flush, streams[0])
    barrier()
    out["config5"] = config5(args, torch, local, rank, world, ctx, synth, streams[0], barrier, allreduce, dist)
    ctx.close()
    del flush
    return out


def _latency(torch, ctx, flush, stream, reps):
    ms = []
    for _ in range(reps):
        flush.zero_()
        ctx.fracture_event()
        ctx.counts()
        ms.append(ctx.last_event_ms()[0])
    return ms


def config3(args, torch, ctx, synth, flush, stream, reps: int = 120):
    """BASELINE configs[2]: 10 000 convex pieces x 256 cells, ONE event: p50 latency from 'pieces + cells resident' to
    'fragment arrays + moments resident' (CUDA events on the context stream, L2 flushed before every event)."""
    pieces = synth.voronoi_cells_batch(ctx, synth.seeds_uniform(1234, 10000), np.array([0, 10000], np.uint32), planes=False)
    cells = synth.voronoi_cells_batch(ctx, synth.seeds_uniform(46354, 256), np.array([0, 256], np.uint32))
    ctx.upload_pieces(pieces.verts, pieces.vert_off, pieces.ring_off, pieces.ring)
    ctx.upload_cells(cells.planes, cells.plane_off, cells.verts, cells.vert_off)
    for _ in range(5):
        ctx.fracture_event()
    c = ctx.counts()
    ms = _latency(torch, ctx, flush, stream, reps)
    ctx.set_profiling(True)
    flush.zero_()
    ctx.fracture_event()
    ph = ctx.last_event_phases()
    ctx.set_profiling(False)
    # one synchronous event end to end (host buffers in and out)
    fr = ctx.download()
    t = []
    for _ in range(10):
        flush.zero_()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        ctx.upload_pieces3(pieces.verts, pieces.vert_off, pieces.ring_off, pieces.ring)
        ctx.upload_cells3(cells.planes, cells.plane_off, cells.verts, cells.vert_off)
        ctx.fracture_event()
        ctx.download_packed()
        t.append(time.perf_counter() - t0)
    # the same through the one-blob wire format with PINNED host buffers: upload blob -> event -> download blob -> sync
    from surtr_b200 import FractureContext
    sizes, total = FractureContext.fill_input_blob(None, pieces, cells)
    h_in = torch.empty(total, dtype=torch.uint8, pin_memory=True)
    FractureContext.fill_input_blob(h_in.numpy(), pieces, cells)
    al = lambda x: (int(x) + 255) // 256 * 256
    cap = al(64 * int(c.n_fragments)) + al(12 * int(c.n_verts)) + al(int(c.n_verts)) + al(2 * int(c.n_ring)) + 4096
    h_out = torch.empty(cap, dtype=torch.uint8, pin_memory=True)
    tb, L = [], None
    for i in range(23):
        flush.zero_()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        ctx.upload_blob_ptr(h_in.data_ptr(), sizes)
        ctx.fracture_event()
        L = ctx.download_blob_into_async(h_out.data_ptr(), cap)
        ctx.sync()
        if i >= 3:
            tb.append(time.perf_counter() - t0)
    got = FractureContext.unpack_output_blob(h_out.numpy(), L)
    assert got.rec.tobytes() == fr.rec.tobytes() and got.verts.tobytes() == fr.verts.tobytes() and np.array_equal(got.ring, fr.ring), \
        "config3: the blob round trip differs from the resident-input result"
    ctx.upload_pieces(pieces.verts, pieces.vert_off, pieces.ring_off, pieces.ring)   # (leave the context as the callers expect it)
    ctx.upload_cells(cells.planes, cells.plane_off, cells.verts, cells.vert_off)
    return {"workload": "config3: 10000 Voronoi pieces (mt19937(1234)) x 256 Voronoi cells (mt19937(46354)), one event",
            "pairs": int(c.n_pairs), "candidates": int(c.n_candidates), "fragments": int(c.n_fragments),
            "tier2_pairs": int(c.n_tier2), "tier3_pairs": int(c.n_tier3),
            "p50_event_ms": float(np.median(ms)), "min_event_ms": float(np.min(ms)), "p90_event_ms": float(np.percentile(ms, 90)),
            "reps": reps, "fragments_per_s": int(c.n_fragments) / (float(np.median(ms)) * 1e-3),
            "kernel_ms": {k: round(v, 4) for k, v in ph.items()},
            "e2e_sync_event_ms": 1e3 * float(np.median(t)),
            "e2e_blob_event_ms": {"p50": 1e3 * float(np.median(tb)), "min": 1e3 * float(np.min(tb)), "reps": len(tb),
                                  "h2d_bytes": int(total), "d2h_bytes": int(L.total)},
            "sum_volume": float(np.sum(fr.rec["volume"])),
            "timing": "CUDA events on the context stream around the whole event, inputs resident, 160 MiB L2 flush before every event (outside the events); "
                      "e2e_sync_event_ms = pageable host arrays in, event, pageable host arrays out, wall clock; e2e_blob_event_ms = one pinned blob up "
                      "(surtr_upload_blob), event, one pinned blob down (surtr_download_blob_async), sync: wall clock of the blocking caller, L2 flushed "
                      "before each, result checked against the resident-input fragments bit for bit"}


def config2(args, torch, dev, local, ctx, synth, flush, stream, n_streams: int = 12, depth: int = 6):
    """BASELINE configs[1]: the unit cube (1 piece) x 4096 Voronoi cells.  One event is a single wave of 4096 K3 warps, so
    the resident figure issues independent events over several streams and the end-to-end figure streams them through
    the C ABI with pinned host buffers (round-1 headline, kept for continuity)."""
    from surtr_b200 import FractureContext, FRAGMENT_DTYPE
    n_seeds = 4096
    cells = synth.voronoi_cells_batch(ctx, synth.seeds_uniform(46354, n_seeds), np.array([0, n_seeds], np.uint32))
    cube_v, cube_vo, cube_ro, cube_r = synth.unit_cube()
    ctx.set_kdop_directions(3)            # one piece that contains every cell: nothing to cull, the AABB is the cheapest set
    ctx.upload_pieces(cube_v, cube_vo, cube_ro, cube_r)
    ctx.upload_cells(cells.planes, cells.plane_off, cells.verts, cells.vert_off)
    for _ in range(5):
        ctx.fracture_event()
    c = ctx.counts()
    n_frag = int(c.n_fragments)
    fr0 = ctx.download()
    ms = _latency(torch, ctx, flush, stream, 60)
    ctx.set_kdop_directions(args.kdop)

    set_bytes = sum(a.nbytes for a in (cube_v, cube_vo, cube_ro, cube_r, cells.planes, cells.plane_off, cells.verts, cells.vert_off))
    n_sets = int(np.ceil(1.3 * L2_MIB * 2 ** 20 / set_bytes / n_streams)) * n_streams
    sts = [stream] + [torch.cuda.Stream(device=dev) for _ in range(1, n_streams)]

    class S:
        pass

    def host_inputs(cs):
        return {k: _pinned(torch, np.ascontiguousarray(v)) for k, v in dict(
            pv=cube_v[:, :3], pvo=cube_vo, pro=cube_ro, pr=cube_r, planes=cs.planes, plane_off=cs.plane_off,
            cverts=cs.verts[:, :3], cvo=cs.vert_off).items()}

    def out_buffers():
        return dict(rec=torch.empty(n_frag * FRAGMENT_DTYPE.itemsize, dtype=torch.uint8, pin_memory=True),
                    verts=torch.empty(int(c.n_verts) * 3, dtype=torch.float32, pin_memory=True),
                    ring_len=torch.empty(int(c.n_verts), dtype=torch.uint8, pin_memory=True),
                    ring=torch.empty(int(c.n_ring), dtype=torch.int16, pin_memory=True))

    sets = []
    for j in range(n_sets):
        s = S()
        s.cx = FractureContext(local, sts[j % n_streams].cuda_stream)
        s.cx.set_kdop_directions(3)
        s.cx.set_clip_build(1)            # several one-wave events in flight on different streams: the throughput build of K3
        rolled = synth.roll_cells(cells, j * (n_seeds // n_sets))
        s.h_in, s.h_out = host_inputs(rolled), out_buffers()
        s.cx.upload_pieces(cube_v, cube_vo, cube_ro, cube_r)
        s.cx.upload_cells(rolled.planes, rolled.plane_off, rolled.verts, rolled.vert_off)
        for _ in range(2):
            s.cx.fracture_event()
        assert int(s.cx.counts().n_fragments) == n_frag
        s.rec_sha = hashlib.sha1(s.cx.download(geometry=False).rec.tobytes()).hexdigest()
        sets.append(s)
    torch.cuda.synchronize()

    # resident: enough steps for a >= 100 ms region
    steps = max(n_sets * 2, 2000)
    t0e, t1e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0e.record(stream)
    for st in sts[1:]:
        st.wait_event(t0e)
    for i in range(steps):
        sets[i % n_sets].cx.fracture_event()
    for st in sts[1:]:
        done = torch.cuda.Event()
        done.record(st)
        stream.wait_event(done)
    t1e.record(stream)
    torch.cuda.synchronize()
    res_ms = t0e.elapsed_time(t1e)

    def up(s):
        hi = s.h_in
        s.cx.upload_pieces3_ptr(hi["pv"].data_ptr(), hi["pvo"].data_ptr(), hi["pro"].data_ptr(), hi["pr"].data_ptr(), 1)
        s.cx.upload_cells3_ptr(hi["planes"].data_ptr(), hi["plane_off"].data_ptr(), hi["cverts"].data_ptr(), hi["cvo"].data_ptr(), n_seeds)

    def down(s):
        ho = s.h_out
        s.cx.download_packed_into_async(ho["rec"].data_ptr(), ho["verts"].data_ptr(), ho["ring_len"].data_ptr(), ho["ring"].data_ptr())

    def pipelined(n):
        for i in range(n + depth):
            if i >= depth:
                down(sets[(i - depth) % n_sets])
            if i < n:
                s = sets[i % n_sets]
                up(s)
                s.cx.fracture_event()
        for s in sets:
            s.cx.sync()

    pipelined(n_sets)
    torch.cuda.synchronize()
    e_steps = max(n_sets * 2, 1000)
    t0 = time.perf_counter()
    pipelined(e_steps)
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    for s in sets:
        assert hashlib.sha1(s.h_out["rec"].numpy().tobytes()).hexdigest() == s.rec_sha, "config2 e2e result differs from the resident-input result"
    assert sets[0].h_out["verts"].numpy().tobytes() == np.ascontiguousarray(fr0.verts[:, :3]).tobytes()
    h2d = sum(t.numel() * t.element_size() for t in sets[0].h_in.values())
    d2h = sum(t.numel() * t.element_size() for t in sets[0].h_out.values())
    for s in sets:
        s.cx.close()
    return {"workload": "config2: unit-cube VMACH (1 piece) x 4096 Voronoi cells, one fracture event per step",
            "fragments_per_event": n_frag, "candidates_per_event": int(c.n_candidates), "kdop_directions": 3,
            "p50_event_ms": float(np.median(ms)),
            "resident": {"value": n_frag * steps / (res_ms * 1e-3), "unit": "fragments/s", "ms_per_event": res_ms / steps, "events": steps,
                         "streams": n_streams, "input_sets": n_sets, "timed_region_ms": res_ms},
            "e2e": {"value": n_frag * e_steps / e2e_s, "unit": "fragments/s", "ms_per_event": 1e3 * e2e_s / e_steps, "events": e_steps,
                    "h2d_bytes_per_event": int(h2d), "d2h_bytes_per_event": int(d2h), "download_lag_events": depth}}


def config5(args, torch, local, rank, world, ctx, synth, stream, barrier, allreduce, dist, n_objects: int = 4096, depth: int = 3, seeds_per_level: int = 64):
    """BASELINE configs[4]: recursive re-fracture to depth 3: every object (unit cube) is cut by 64 cells, every fragment
    again by the next level's 64 cells (mt19937(1000 + level)), and once more; object o belongs to rank o mod N and its
    fragments never leave the device between levels (surtr_fragments_to_pieces).  The level's pattern is uploaded once
    and replicated per object on the device (surtr_upload_pattern / surtr_place_pattern)."""
    from surtr_b200 import sharding
    mine = sharding.event_shard(n_objects, world, rank)
    n_obj = len(mine)
    cube_v, cube_vo, cube_ro, cube_r = synth.unit_cube()
    verts = np.tile(cube_v, (n_obj, 1))
    vo = (np.arange(n_obj + 1, dtype=np.uint32) * 8)
    ro = (np.arange(8 * n_obj + 1, dtype=np.uint32) * 3)
    ring = np.tile(cube_r, n_obj)
    ev0 = np.arange(n_obj + 1, dtype=np.uint32)
    patterns = []
    for lvl in range(depth):
        cs = synth.voronoi_cells_batch(ctx, synth.seeds_uniform(1000 + lvl, seeds_per_level), np.array([0, seeds_per_level], np.uint32))
        patterns.append(synth.pattern_arrays(cs.verts, cs.vert_off, cs.ring_off, cs.ring))
    ones, zeros = np.ones((n_obj, 3), np.float32), np.zeros((n_obj, 3), np.float32)

    def recurse():
        ctx.upload_pieces(verts, vo, ro, ring, ev0)
        dev_ms, counts = 0.0, []
        for fv, fvo, cfo in patterns:
            ctx.upload_pattern(fv, fvo, cfo)
            ctx.place_pattern(ones, zeros)
            ctx.fracture_event()
            cc = ctx.counts()
            dev_ms += ctx.last_event_ms()[0]
            counts.append(int(cc.n_fragments))
            ctx.fragments_to_pieces_per_event()   # object o keeps its fragments: boundaries found on the device
        return dev_ms, counts

    recurse()
    barrier()
    runs = []
    for _ in range(3):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        dev_ms, counts = recurse()
        torch.cuda.synchronize()
        runs.append((time.perf_counter() - t0, dev_ms, counts))
    runs.sort(key=lambda r: r[0])
    wall_s, dev_ms, counts = runs[1]
    wall_max = allreduce(wall_s, dist.ReduceOp.MAX)
    dev_max = allreduce(dev_ms, dist.ReduceOp.MAX)
    final = allreduce(counts[-1], dist.ReduceOp.SUM)
    allf = allreduce(sum(counts), dist.ReduceOp.SUM)
    return {"workload": f"config5: {n_objects} unit cubes re-fractured to depth {depth} ({seeds_per_level} cells per level), object o -> rank o mod N",
            "objects": n_objects, "objects_per_rank": n_obj, "fragments_per_object_per_level": [cnt // max(1, n_obj) for cnt in counts],
            "final_fragments": int(final), "fragments_all_levels": int(allf),
            "device_ms_max_over_ranks": dev_max, "wall_ms_max_over_ranks": 1e3 * wall_max,
            "final_fragments_per_s_device": final / (dev_max * 1e-3), "fragments_all_levels_per_s_device": allf / (dev_max * 1e-3),
            "final_fragments_per_s_wall": final / wall_max,
            "timing": "device = sum of the three events' CUDA-event times; wall = whole recursion incl. the pattern uploads and the regrouping of fragments by object "
                      "(surtr_fragments_to_pieces_per_event: the boundaries are found on the device, 4 bytes per object come back) "
                      "between levels; median of 3, max over ranks"}
