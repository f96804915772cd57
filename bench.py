#!/usr/bin/env python
"""bench.py -- clipped convex fragments per second on N B200s (BASELINE.json metric).

A "step" is one fracture event: the whole hot path (K1 k-DOP extents -> K2 broad phase + compaction -> K3 clip ->
K4 assembly) over one batch of synthetic input.  Workload at every N: BASELINE.json configs[1], the synthetic
unit-cube VMACH (1 piece) fractured by 4096 Voronoi seeds, one independent event per rank (rank r uses seed
46354 + r): events shard across GPUs with no data-path collective (weak scaling).

  value   whole-job fragments/s with inputs resident in HBM: K events issued over several streams, CUDA events around
          the whole region on stream 0, max over ranks.  The resident inputs are LARGER THAN THE L2: every step runs on
          another of N_SETS resident input sets (own inputs, scratch and outputs), so no step finds its data cached;
  e2e     the same metric through the C ABI with HOST buffers: pinned-host -> device upload of the event's pieces
          and cells, the event, and the device -> pinned-host download of every fragment, all inside the timed region;
  roofline  dominant kernel (K3 clip, tier 1): algorithmic bytes / its CUDA-event duration
          against the measured HBM peak of MEASURED_PEAKS.json;
  cpu_baseline  the reference's own CPU path (oracle/_ref, built from the reference sources) timed on this box.

--impl reference times only that CPU path (the reference arm), same metric / config.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_SEEDS = 4096
BASE_SEED = 46354
WORKLOAD = "config2: unit-cube VMACH (1 piece) x 4096 Voronoi cells, one fracture event per step per GPU"
FLUSH_MIB = 160   # L2 flush buffer (B200 L2 = 126 MB) for the one-event-at-a-time latency loops
L2_MIB = 126      # B200 L2
INPUT_X_L2 = 1.3  # the resident input sets of the throughput loops add up to at least this many L2 sizes
E2E_DEPTH = 6     # end-to-end loop: the download of event i is enqueued when event i + E2E_DEPTH is issued
STREAMS = 12      # streams the input sets are bound to round-robin = independent events the GPU may overlap


# ------------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    """nvidia-smi sampling during the timed region (B200_PROFILING.md clocks line)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.lines = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.index)], stdout=subprocess.PIPE, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self) -> dict:
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, smax, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                smax.append(float(f[2]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------ reference arm
def cpu_reference(planes, plane_off, budget_s: float, threads: int):
    """The reference's own CPU fan-out (one task per cell on dp::thread_pool, Surtr.cpp:28, 2129-2131) from
    oracle/_ref; falls back to the oracle's single-thread C port only if the reference build is absent."""
    from oracle import portapi, refapi
    verts, vo, ro, ring = __import__("surtr_b200.synth", fromlist=["x"]).unit_cube()
    cube = refapi.PolySet(verts, vo, ro, ring)
    if refapi.available():
        kind, run = "reference", (lambda: refapi.apply_fracture(cube, planes, plane_off, threads, False))
        cores = threads
    else:
        kind, run = "port", (lambda: portapi.apply_fracture(cube, planes, plane_off))
        cores = 1
    run()
    reps, frags, secs = 0, 0, 0.0
    t_end = time.perf_counter() + budget_s
    while True:
        t0 = time.perf_counter()
        r = run()
        # reference build: its own timer around the fan-out + SetExtract (Surtr.cpp:1917-1924 "ApplyFracture" timer);
        # flattening into flat arrays for the caller is not part of the reference's event
        secs += r.seconds if kind == "reference" else time.perf_counter() - t0
        frags += r.n
        reps += 1
        if time.perf_counter() >= t_end or reps >= 400:
            break
    return {"value": frags / secs, "unit": "fragments/s", "cores": cores, "kind": kind,
            "sample": f"{reps} full events of the workload ({frags // reps} fragments each), {secs:.2f} s of wall time, "
                      f"{os.cpu_count()} host threads available"}, secs / reps


def load_or_build_cells_cpu(seed):
    """Cells for the reference arm (no GPU): the oracle's own builder."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from oracle import portapi
    from surtr_b200 import synth
    s = synth.seeds_uniform(seed, N_SEEDS)
    off, idx = synth.delaunay_neighbors(s)
    return portapi.voronoi_cells(s, off, idx)


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cells = load_or_build_cells_cpu(BASE_SEED)
    threads = max(16, os.cpu_count() or 1)
    from oracle import refapi, portapi
    verts, vo, ro, ring = __import__("surtr_b200.synth", fromlist=["x"]).unit_cube()
    cube = refapi.PolySet(verts, vo, ro, ring)
    use_ref = refapi.available()
    run = (lambda: refapi.apply_fracture(cube, cells.planes, cells.plane_off, threads, False)) if use_ref else \
          (lambda: portapi.apply_fracture(cube, cells.planes, cells.plane_off))
    for _ in range(max(1, args.warmup)):
        run()
    t, frags = 0.0, 0
    for _ in range(args.steps):
        t0 = time.perf_counter()
        r = run()
        t += r.seconds if use_ref else time.perf_counter() - t0     # fan-out + SetExtract (see cpu_reference)
        frags += r.n
    val = frags / t
    line = {
        "metric": "clipped fragments/sec", "value": val, "unit": "fragments/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic", "impl": "reference",
        "config": {"workload": WORKLOAD, "note": "reference CPU path: each step = one full event on the host cores"},
        "cpu_baseline": {"value": val, "unit": "fragments/s", "cores": threads if use_ref else 1,
                         "kind": "reference" if use_ref else "port",
                         "sample": f"{args.steps} full events, dp::thread_pool({threads}) one task per cell, "
                                   f"{os.cpu_count()} host threads available"},
        "e2e": {"value": val, "unit": "fragments/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------ GPU arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="surtr_b200", choices=["surtr_b200", "reference"])
    ap.add_argument("--kdop", type=int, default=3,
                    help="broad-phase direction set; the workload's single piece contains every cell, so no pair can be "
                         "culled and the cheapest set (AABB) is used -- the library default is 13")
    ap.add_argument("--cpu-budget", type=float, default=10.0, help="seconds of CPU-baseline sampling (rank 0, N=1)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--e2e-depth", type=int, default=E2E_DEPTH, help="events between launch and download in the end-to-end loop")
    ap.add_argument("--streams", type=int, default=STREAMS, help="streams of the throughput loops")
    ap.add_argument("--wire", default="packed", choices=["packed", "float4"],
                    help="host<->device format of the end-to-end loop: packed = surtr_upload_*3 + surtr_download_fragments_packed")
    ap.add_argument("--input-sets", type=int, default=0,
                    help="resident input sets cycled by the throughput loops (0 = as many as make the inputs exceed 1.3 x L2; "
                         "a smaller number is for profiler runs only and is reported in config)")
    args = ap.parse_args()
    globals()["E2E_DEPTH"] = max(1, args.e2e_depth)
    globals()["STREAMS"] = max(1, args.streams)
    args.warmup = max(args.warmup, 3)

    if args.impl == "reference":
        run_reference_arm(args)
        return

    import torch
    import torch.distributed as dist
    from surtr_b200 import FractureContext, synth

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a B200: the product path has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    # a dedicated non-default stream: the engine launches on it and every timing event is recorded on it
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    assert stream.cuda_stream != 0
    ctx = FractureContext(local, stream.cuda_stream)
    ctx.set_kdop_directions(args.kdop)

    # ---- synthetic input of the named shape, built by the product path (GPU clipper) ----
    seeds = synth.seeds_uniform(BASE_SEED + rank, N_SEEDS)
    cells = synth.voronoi_cells(ctx, seeds)
    cube_v, cube_vo, cube_ro, cube_r = synth.unit_cube()

    def pin(a):
        t = torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
        return t

    PACKED = args.wire == "packed"

    def host_inputs(cs):
        """Pinned host buffers of one event in the PCIe wire format of the C ABI (float3 vertex streams when packed)."""
        nc = 3 if PACKED else 4
        return {k: pin(v) for k, v in dict(pv=cube_v[:, :nc], pvo=cube_vo, pro=cube_ro, pr=cube_r, planes=cs.planes,
                                           plane_off=cs.plane_off, cverts=cs.verts[:, :nc], cvo=cs.vert_off).items()}

    def resident_bytes(cs):
        return sum(a.nbytes for a in (cube_v, cube_vo, cube_ro, cube_r, cs.planes, cs.plane_off, cs.verts, cs.vert_off))

    h_in = host_inputs(cells)
    h2d_bytes = sum(t.numel() * t.element_size() for t in h_in.values())
    set_bytes = resident_bytes(cells)          # float4 streams as they lie in HBM

    def upload_resident(cx, cs):
        cx.upload_pieces(cube_v, cube_vo, cube_ro, cube_r)
        cx.upload_cells(cs.planes, cs.plane_off, cs.verts, cs.vert_off)

    def upload_to(cx, hi=None):
        hi = hi or h_in
        up_p, up_c = (cx.upload_pieces3_ptr, cx.upload_cells3_ptr) if PACKED else (cx.upload_pieces_ptr, cx.upload_cells_ptr)
        up_p(hi["pv"].data_ptr(), hi["pvo"].data_ptr(), hi["pro"].data_ptr(), hi["pr"].data_ptr(), 1)
        up_c(hi["planes"].data_ptr(), hi["plane_off"].data_ptr(), hi["cverts"].data_ptr(), hi["cvo"].data_ptr(), N_SEEDS)

    upload_resident(ctx, cells)
    ctx.fracture_event()
    c0 = ctx.counts()
    n_frag = int(c0.n_fragments)
    fr0 = ctx.download()
    alg_bytes = synth.algorithmic_bytes(cube_vo, cube_ro, cells.plane_off, fr0.rec)
    alg_flops = synth.algorithmic_flops(cube_vo, cells.plane_off, fr0.rec)
    fp32_peak_tflops = ctx.measure_fp32_peak()

    flush = torch.empty(FLUSH_MIB * 1024 * 1024, dtype=torch.uint8, device=dev)   # > 126 MB L2

    def flush_l2():
        flush.zero_()

    # ---- warm-up ----
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        time.sleep(0.3)      # let nvidia-smi come up; it keeps sampling through warm-up, timed region and e2e
    for _ in range(args.warmup):
        flush_l2()
        ctx.fracture_event()
    ctx.counts()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()

    # ---- resident input sets for the throughput loops: MORE INPUT THAN THE L2 HOLDS ----
    # N_SETS contexts, each with its own resident inputs (the pattern with its cells renumbered, so every set is a
    # different byte stream), scratch and output arrays, bound round-robin to STREAMS streams.  Step i runs on set
    # i mod N_SETS: by the time a set comes round again, N_SETS - 1 other sets (inputs alone > 1.3 x L2, plus their
    # scratch and outputs) have gone through the L2, so nothing of it is cached -- no flush kernel inside the loop.
    n_sets = int(np.ceil(INPUT_X_L2 * L2_MIB * 1024 * 1024 / set_bytes / STREAMS)) * STREAMS
    if args.input_sets > 0:
        n_sets = args.input_sets
    streams = [stream] + [torch.cuda.Stream(device=dev) for _ in range(1, STREAMS)]
    import hashlib

    class InputSet:
        pass

    sets = []
    for j in range(n_sets):
        s_ = InputSet()
        s_.st = streams[j % STREAMS]
        if j == 0:
            s_.cx, s_.h_in = ctx, h_in
        else:
            s_.cx = FractureContext(local, s_.st.cuda_stream)
            s_.cx.set_kdop_directions(args.kdop)
            rolled = synth.roll_cells(cells, j * (N_SEEDS // n_sets))
            s_.h_in = host_inputs(rolled)
            upload_resident(s_.cx, rolled)
        for _ in range(args.warmup):
            s_.cx.fracture_event()
        assert int(s_.cx.counts().n_fragments) == n_frag
        s_.rec_sha = hashlib.sha1(s_.cx.download(geometry=False).rec.tobytes()).hexdigest()
        sets.append(s_)
    torch.cuda.synchronize()
    resident_input_mib = n_sets * set_bytes / 2 ** 20

    # ---- latency: K single events back to back on one stream, device-resident inputs ----
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    clip_ms = []
    for i in range(args.steps):
        flush_l2()
        ev[i][0].record(stream)
        ctx.fracture_event()
        ev[i][1].record(stream)
    torch.cuda.synchronize()
    step_ms = [a.elapsed_time(b) for a, b in ev]

    # ---- timed region: K steps, device-resident inputs, STREAMS streams ----
    # One event does not fill the GPU (4096 warps of K3 = 28 per SM, issue-latency bound), so a job of K independent
    # events is issued round-robin over the input sets and their streams.  Timed with CUDA events on stream 0: the
    # start event gates the other streams, the stop event waits for all of them.
    launches = 0
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t_wall0 = time.perf_counter()
    t0_ev, t1_ev = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0_ev.record(stream)
    for st in streams[1:]:
        st.wait_event(t0_ev)
    for i in range(args.steps):
        cx = sets[i % n_sets].cx
        cx.fracture_event()
        launches += cx.last_event_launches()
    for st in streams[1:]:
        done = torch.cuda.Event()
        done.record(st)
        stream.wait_event(done)
    t1_ev.record(stream)
    torch.cuda.synchronize()
    t_wall = time.perf_counter() - t_wall0
    for s_ in sets[:min(n_sets, args.steps)]:
        assert int(s_.cx.counts().n_fragments) == n_frag
    # dominant-kernel duration for the roofline: single events again with the engine's per-kernel CUDA events on
    # (they sit between the kernels of an event and serialise the programmatic dependent launches, so the
    # loops above run without them)
    ctx.set_profiling(True)
    for i in range(min(args.steps, 64)):
        flush_l2()
        ctx.fracture_event()
        clip_ms.append(ctx.last_event_ms()[1])
    ctx.set_profiling(False)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    total_ms = torch.tensor([float(t0_ev.elapsed_time(t1_ev))], device=dev, dtype=torch.float64)
    frags = torch.tensor([float(n_frag * args.steps)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(total_ms, op=dist.ReduceOp.MAX)
        dist.all_reduce(frags, op=dist.ReduceOp.SUM)
    total_ms = float(total_ms.item())
    value = float(frags.item()) / (total_ms * 1e-3)

    # ---- e2e: host buffers in, host buffers out, every step ----
    # A caller that streams events keeps a few of them in flight: the input sets (one context each, STREAMS
    # streams) are driven round-robin from this one host thread through the C ABI -- download step i-DEPTH (blocks
    # on that event only), then upload + launch step i -- so the PCIe copies of one event overlap the kernels of the
    # others.  Every step moves its own inputs host->device from its set's pinned buffers and its own fragments
    # device->host; the device-side arrays it touches belong to a set last used N_SETS steps ago (cold L2).
    c = ctx.counts()
    from surtr_b200 import FRAGMENT_DTYPE

    def out_buffers():
        # wire format of the fragments: records, float3 positions, one byte of ring length per vertex, ring entries
        return dict(rec=torch.empty(int(c.n_fragments) * FRAGMENT_DTYPE.itemsize, dtype=torch.uint8).pin_memory(),
                    verts=torch.empty(int(c.n_verts) * (3 if PACKED else 4), dtype=torch.float32).pin_memory(),
                    ring_len=torch.empty(int(c.n_verts) if PACKED else 4 * (int(c.n_verts) + 1), dtype=torch.uint8).pin_memory(),
                    ring=torch.empty(int(c.n_ring), dtype=torch.int16).pin_memory())

    h_out = out_buffers()
    d2h_bytes = sum(t.numel() * t.element_size() for t in h_out.values())

    def download_from(cx, ho):
        (cx.download_packed_into if PACKED else cx.download_into)(ho["rec"].data_ptr(), ho["verts"].data_ptr(),
                                                                   ho["ring_len"].data_ptr(), ho["ring"].data_ptr())

    # (1) one event at a time: the latency a single synchronous caller sees
    for _ in range(3):
        upload_to(ctx); ctx.fracture_event(); download_from(ctx, h_out)
    torch.cuda.synchronize()
    sync_s = 0.0
    for _ in range(args.steps):
        flush_l2()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        upload_to(ctx); ctx.fracture_event(); download_from(ctx, h_out)     # download synchronises the stream
        sync_s += time.perf_counter() - t0
    got = np.frombuffer(h_out["rec"].numpy().tobytes(), dtype=FRAGMENT_DTYPE)
    assert got.tobytes() == fr0.rec.tobytes(), "e2e result differs from the resident-input result"

    # (2) E2E_DEPTH events in flight: the throughput number
    for j, s_ in enumerate(sets):
        s_.h_out = h_out if j == 0 else out_buffers()

    def pipelined(n_steps):
        for i in range(n_steps + E2E_DEPTH):
            if i >= E2E_DEPTH:
                # waits for event i-DEPTH (long finished when the depth is enough), then only ENQUEUES its device->host
                # copies on that context's copy stream; the host moves on
                s_ = sets[(i - E2E_DEPTH) % n_sets]
                ho = s_.h_out
                (s_.cx.download_packed_into_async if PACKED else s_.cx.download_into_async)(
                    ho["rec"].data_ptr(), ho["verts"].data_ptr(), ho["ring_len"].data_ptr(), ho["ring"].data_ptr())
            if i < n_steps:
                s_ = sets[i % n_sets]
                upload_to(s_.cx, s_.h_in)
                s_.cx.fracture_event()
        for s_ in sets:
            s_.cx.sync()                    # drain: the last copies have landed in the host buffers

    pipelined(n_sets)                       # warm-up: every context grows its buffers once
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    pipelined(args.steps)
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    torch.cuda.set_stream(stream)
    for s_ in sets:
        got = s_.h_out["rec"].numpy().tobytes()
        assert hashlib.sha1(got).hexdigest() == s_.rec_sha, "pipelined e2e result differs from the resident-input result"
    # geometry of the last download of set 0 against the resident-input result of the same set
    v3 = sets[0].h_out["verts"].numpy().reshape(-1, 3 if PACKED else 4)[:, :3]
    assert np.ascontiguousarray(v3).tobytes() == np.ascontiguousarray(fr0.verts[:, :3]).tobytes(), "e2e vertex positions differ"
    if PACKED:
        assert np.array_equal(sets[0].h_out["ring_len"].numpy(), np.diff(fr0.ring_off).astype(np.uint8)), "e2e ring lengths differ"
    else:
        assert sets[0].h_out["ring_len"].numpy().view(np.uint32).tobytes() == fr0.ring_off.tobytes(), "e2e ring offsets differ"
    assert sets[0].h_out["ring"].numpy().tobytes() == fr0.ring.tobytes(), "e2e ring entries differ"
    e2e_t = torch.tensor([e2e_s], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(e2e_t, op=dist.ReduceOp.MAX)
    e2e_value = float(frags.item()) / float(e2e_t.item())
    clocks = sampler.stop() if rank == 0 else None
    if clocks is not None:
        clocks["window"] = "warm-up + timed region + e2e region (the timed region alone is shorter than one sample)"
    # ---- final fragment gather (the only collective; after the hot path, reported separately) ----
    gather = None
    if world > 1:
        from surtr_b200 import sharding
        d_out = {k: v.to(dev, non_blocking=True) for k, v in h_out.items()}
        torch.cuda.synchronize()
        dist.barrier()
        sharding.gather_fragments(d_out["rec"], d_out["verts"], d_out["ring_len"], d_out["ring"], dst=0)   # NCCL warm-up
        torch.cuda.synchronize()
        dist.barrier()
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        g0.record()
        parts = sharding.gather_fragments(d_out["rec"], d_out["verts"], d_out["ring_len"], d_out["ring"], dst=0)
        g1.record()
        torch.cuda.synchronize()
        if rank == 0:
            assert len(parts) == world and all(p[0].numel() == d_out["rec"].numel() or True for p in parts)
            gather = {"ms": g0.elapsed_time(g1), "bytes_per_rank": int(d2h_bytes),
                      "fragments_gathered": int(sum(p[0].numel() for p in parts) // FRAGMENT_DTYPE.itemsize),
                      "backend": "nccl all_gather(counts) + padded gather to rank 0",
                      "arrays": "records, vertex positions, ring lengths / offsets, ring entries as downloaded (wire format)"}

    # ---- roofline of the dominant kernel + CPU baseline (rank 0) ----
    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        k3_ms = float(np.mean(clip_ms))
        achieved = alg_bytes / (k3_ms * 1e-3) / 1e9
        traffic = None
        try:
            prof = json.load(open(os.path.join(ROOT, "profiles", "k3_traffic.json")))
            traffic = prof.get("dram_bytes_per_launch")
        except Exception:
            pass
        roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                    "traffic": traffic, "kernel": "clip_sub_kernel<32> (K3, small tier: one warp per pair)", "kernel_ms": k3_ms,
                    "algorithmic_bytes_per_launch": alg_bytes,
                    "peak_source": "MEASURED_PEAKS.json hbm_gbs (measured)" if peaks else "fallback 6650 GB/s"}
        line = {
            "metric": "clipped fragments/sec", "value": value, "unit": "fragments/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": total_ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "fragments_per_step_per_gpu": n_frag, "pairs_per_step": int(c0.n_pairs),
                       "candidates_per_step": int(c0.n_candidates), "kdop_directions": args.kdop,
                       "streams": STREAMS,
                       "input_sets": n_sets, "resident_input_mib": round(resident_input_mib, 1),
                       "l2": (f"PROFILER RUN, NOT a bench value: only {n_sets} input sets ({resident_input_mib:.0f} MiB < L2)"
                              if resident_input_mib < L2_MIB else
                              f"inputs larger than L2: {n_sets} resident input sets ({resident_input_mib:.0f} MiB of inputs > "
                              f"{L2_MIB} MB L2, each with its own scratch and output arrays) cycled step by step, so a set is "
                              f"revisited only after {n_sets - 1} other events; no flush kernel in the timed region "
                              f"(the one-event-at-a-time figures flush {FLUSH_MIB} MiB between steps, outside their timers)"),
                       "timing": "CUDA events on stream 0 around the K steps (start gates, stop joins all streams), max over ranks",
                       "parallelism": f"events sharded over {world} GPU(s), no data-path collective"},
            "p50_event_ms": float(np.median(step_ms)),
            "single_stream": {"value": n_frag * world / (float(np.mean(step_ms)) * 1e-3), "ms_per_step": float(np.mean(step_ms)),
                              "note": "one event at a time, per-step CUDA events, L2 flushed between steps outside the events"},
            "wall_s_timed_region": t_wall,
            "e2e": {"value": e2e_value, "unit": "fragments/s", "h2d_bytes_per_step": int(h2d_bytes),
                    "d2h_bytes_per_step": int(d2h_bytes), "ms_per_step": 1e3 * float(e2e_t.item()) / args.steps,
                    "streams": STREAMS, "download_lag_events": E2E_DEPTH, "single_event_ms": 1e3 * sync_s / args.steps,
                    "wire_format": ("surtr_upload_pieces3 / surtr_upload_cells3 (float3 vertex streams, widened to float4 on the "
                                    "device) and surtr_download_fragments_packed (float3 + one byte of ring length per vertex)")
                                   if PACKED else "float4 vertex streams and 32-bit ring offsets (the resident layout)",
                    "timing": "wall clock around K x (upload + event + download) through the C ABI, pinned host buffers, "
                              f"{n_sets} input sets (own pinned host buffers) over {STREAMS} streams from one host thread, the download of "
                              f"an event enqueued {E2E_DEPTH} events after its launch, inputs "
                              "larger than L2 as in the timed region; single_event_ms = the same with one event at a time "
                              "(L2 flushed between events, outside the timer)"},
            "gpu_launches": int(launches), "launches_per_step": int(launches // max(1, args.steps)),
            "clocks": clocks, "roofline": roofline,
            # secondary roofline the north star asks for: algorithmic FP32 work of the whole job against the FFMA rate
            # measured in this run (both tiny by design: the path is topology work, not arithmetic)
            "fp32": {"algorithmic_flops_per_step": alg_flops, "achieved_gflops": alg_flops / (total_ms / args.steps * 1e-3) / 1e9,
                     "peak_tflops_measured": fp32_peak_tflops,
                     "frac": alg_flops / (total_ms / args.steps * 1e-3) / 1e12 / fp32_peak_tflops if fp32_peak_tflops else None},
        }
        if gather:
            line["gather"] = gather
        if world == 1 and not args.no_cpu_baseline:
            cb, _ = cpu_reference(cells.planes, cells.plane_off, args.cpu_budget, max(16, os.cpu_count() or 1))
            line["cpu_baseline"] = cb
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    for s_ in sets[1:]:
        s_.cx.close()
    ctx.close()


if __name__ == "__main__":
    main()
