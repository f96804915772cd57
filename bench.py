#!/usr/bin/env python
"""bench.py -- clipped convex fragments per second on N B200s (BASELINE.json metric), on BASELINE's own configs.

Workload of the headline numbers, at every N: BASELINE.json configs[3] ("config 4"): 4096 independent fracture events,
event e = the 1000 Voronoi cells of mt19937(1234 + e) (the object's convex pieces) cut by the 64 Voronoi cells of
mt19937(46354 + e) (the fracture pattern).  Event e belongs to rank e mod N (surtr_b200/sharding.py): total work is
fixed, so this is a STRONG-scaling curve with no collective on the hot path.  A "step" is one pass of the whole hot path
(K1 k-DOP extents -> K2 broad phase + ordered compaction -> K3 clip -> K4 assembly with moments) over the rank's share
of the 4096 events, issued as batches of independent events through ev_piece_off / ev_cell_off.

  value     whole-job fragments/s with every event's inputs resident in HBM (a few GB per rank: far more than the L2),
            CUDA events around K steps on the engine's stream, max over ranks;
  e2e       the same job through the C ABI with HOST buffers: every step uploads every event (pinned host memory, float3
            wire format), cuts it and downloads every fragment, batches pipelined over a few contexts;
  roofline  K3 small tier (clip_fast_kernel), the dominant kernel: algorithmic bytes of the launch /
            its duration from CUDA events between the kernels, against the measured HBM peak (MEASURED_PEAKS.json);
  kernels   the same for every kernel of the event (share of the step, GB/s of its own algorithmic bytes);
  config3 / config2 / config5   the other BASELINE configs as secondary results: p50 latency of the 10 000 x 256 event,
            the unit cube x 4096 cells event (resident + end-to-end event stream), depth-3 re-fracture of 4096 objects;
  cpu_baseline  the reference's own CPU path (oracle/_ref, built from the reference sources, dp::thread_pool(16)) timed
            on this box on a bounded sample of the same events.

--impl reference times only that CPU path (the reference arm): same metric, same config dict.
"""
from __future__ import annotations

import argparse
import hashlib
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_EVENTS = 4096
PIECES_PER_EVENT = 1000
CELLS_PER_EVENT = 64
REF_THREADS = 16          # Src/Surtr.cpp:28: dp::thread_pool<> g_threadPool(16)
L2_MIB = 126              # B200 L2
FLUSH_MIB = 160           # L2 flush buffer for the one-event-at-a-time latency loops (configs 2, 3)
RESIDENT_BATCH = 512      # events per resident batch (one context each)
E2E_BATCH = 128           # events per end-to-end batch (sweep in profiles/r2_e2e_sweep.md: 128 x 4 contexts > 64 x 3 > 32 x 4)
E2E_CONTEXTS = 4          # batches in flight in the end-to-end loop


def workload_config(n_events: int) -> dict:
    """The config dict BOTH arms print (identical by construction)."""
    return {"workload": f"config4: {n_events} independent fracture events (1000 Voronoi pieces x 64 Voronoi cells each), "
                        "event e -> rank e mod N, no data-path collective",
            "events": n_events, "pieces_per_event": PIECES_PER_EVENT, "cells_per_event": CELLS_PER_EVENT,
            "seeds": "pieces mt19937(1234 + e), cells mt19937(46354 + e), uniform in the unit box (SURVEY.md section 8d)"}


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

    def mark(self) -> int:
        return len(self.lines)

    def stop(self, lo: int = 0, hi: int | None = None) -> dict:
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, smax, reasons = [], [], set()
        for ln in self.lines[lo:hi]:
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
def reference_events(n_sample: int):
    """Events 0..n_sample-1 of the workload built WITHOUT the product path (oracle port + qhull neighbours)."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import common
    return [(common.voronoi(1234 + e, PIECES_PER_EVENT), common.voronoi(46354 + e, CELLS_PER_EVENT)) for e in range(n_sample)]


def time_reference(events, threads: int, budget_s: float, max_passes: int = 1000):
    """The reference's own CPU fan-out (one task per cell on dp::thread_pool, Surtr.cpp:28, 2129-2131, + SetExtract)
    from oracle/_ref over the sample events, repeated until the budget is spent.  Falls back to the oracle's
    single-thread C port only if the reference build is absent.  Returns (fragments/s, kind, cores, passes, seconds)."""
    from oracle import portapi, refapi
    use_ref = refapi.available()

    def run(p, c):
        if use_ref:
            r = refapi.apply_fracture(p, c.planes, c.poly_face_off, threads, False)
            return r.n, r.seconds          # the reference build's own timer around the fan-out + SetExtract
        t0 = time.perf_counter()
        r = portapi.apply_fracture(p, c.planes, c.poly_face_off)
        return r.n, time.perf_counter() - t0

    for p, c in events[:2]:
        run(p, c)
    frags, secs, passes = 0, 0.0, 0
    t_end = time.perf_counter() + budget_s
    while passes < max_passes:
        for p, c in events:
            n, s = run(p, c)
            frags += n
            secs += s
        passes += 1
        if time.perf_counter() >= t_end:
            break
    return frags / secs, ("reference" if use_ref else "port"), (threads if use_ref else 1), passes, secs


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n_sample = max(1, min(args.ref_sample, args.events))
    events = reference_events(n_sample)
    from oracle import refapi
    use_ref = refapi.available()

    def run(p, c, threads):
        if use_ref:
            r = refapi.apply_fracture(p, c.planes, c.poly_face_off, threads, False)
            return r.n, r.seconds
        from oracle import portapi
        t0 = time.perf_counter()
        r = portapi.apply_fracture(p, c.planes, c.poly_face_off)
        return r.n, time.perf_counter() - t0

    def steps(threads, n_steps, warm):
        for _ in range(warm):
            for p, c in events:
                run(p, c, threads)
        t, frags = 0.0, 0
        for _ in range(n_steps):
            for p, c in events:
                n, s = run(p, c, threads)
                frags += n
                t += s
        return frags, t

    frags, t = steps(REF_THREADS, args.steps, max(1, args.warmup))
    val = frags / t
    nproc = os.cpu_count() or 1
    wide = None
    if use_ref and nproc != REF_THREADS:
        f2, t2 = steps(nproc, max(2, args.steps // 4), 1)
        wide = {"threads": nproc, "value": f2 / t2, "note": "same sample with one pool thread per host thread (not the reference's stock pool size)"}
    sample = (f"each step = events 0..{n_sample - 1} of the workload ({frags // args.steps} fragments), fan-out + SetExtract timed by the "
              f"reference build's own timer; dp::thread_pool({REF_THREADS}) as in Src/Surtr.cpp:28; {nproc} host threads available")
    line = {
        "metric": "clipped fragments/sec", "value": val, "unit": "fragments/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t / args.steps, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic", "impl": "reference",
        "config": workload_config(args.events),
        "cpu_baseline": {"value": val, "unit": "fragments/s", "cores": REF_THREADS if use_ref else 1,
                         "kind": "reference" if use_ref else "port", "sample": sample, "all_host_threads": wide},
        "e2e": {"value": val, "unit": "fragments/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------ helpers (GPU arm)
def numa_bind(local: int):
    """Bind this rank's threads to the CPUs next to its GPU before any pinned allocation (first-touch places the
    pinned pages on that NUMA node, so host<->device copies do not cross the socket interconnect)."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(local)
        words = (os.cpu_count() + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, words)
        cpus = [64 * w + b for w, m in enumerate(mask) for b in range(64) if (int(m) >> b) & 1]
        if cpus:
            os.sched_setaffinity(0, cpus)
            return {"cpus": f"{cpus[0]}-{cpus[-1]}", "n": len(cpus)}
    except Exception as e:       # not fatal: the numbers are then whatever the default placement gives
        return {"error": str(e)[:80]}
    return None


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="surtr_b200", choices=["surtr_b200", "reference"])
    ap.add_argument("--events", type=int, default=N_EVENTS, help="events of the job (BASELINE config 4: 4096); smaller = a test run, reported in config")
    ap.add_argument("--kdop", type=int, default=13, help="broad-phase direction set (library default 13 = 26-DOP)")
    ap.add_argument("--cpu-budget", type=float, default=12.0, help="seconds of CPU-baseline sampling (rank 0, N=1)")
    ap.add_argument("--ref-sample", type=int, default=8, help="events per step of the reference arm / CPU baseline")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-secondary", action="store_true", help="skip configs 2, 3 and 5 (profiler runs)")
    ap.add_argument("--resident-batch", type=int, default=RESIDENT_BATCH)
    ap.add_argument("--e2e-batch", type=int, default=E2E_BATCH)
    ap.add_argument("--e2e-contexts", type=int, default=E2E_CONTEXTS)
    ap.add_argument("--no-numa", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    if args.resident_batch % args.e2e_batch:
        raise SystemExit("--resident-batch must be a multiple of --e2e-batch (the end-to-end batches are checked against the resident ones)")

    if args.impl == "reference":
        run_reference_arm(args)
        return

    import torch
    import torch.distributed as dist
    from surtr_b200 import FractureContext, synth, sharding, FRAGMENT_DTYPE
    import bench_secondary

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a B200: the product path has no CPU fallback (use --impl reference for the CPU arm)")
    numa = None if args.no_numa else numa_bind(local)
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    def allreduce(x, op):
        t = torch.tensor([float(x)], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=op)
        return float(t.item())

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    # two non-default streams; the engine launches on them and every timing event is recorded on stream 0
    streams = [torch.cuda.Stream(device=dev) for _ in range(2)]
    torch.cuda.set_stream(streams[0])
    t_setup0 = time.perf_counter()

    # ---- synthetic input of the named shape, built by the product path (host DT3D + GPU clipper) ----
    my_events = sharding.event_shard(args.events, world, rank)
    gen = FractureContext(local, streams[0].cuda_stream)
    pieces, cells, ev_p, ev_c = synth.config4_events(gen, my_events, PIECES_PER_EVENT, CELLS_PER_EVENT)
    gen.close()
    n_my = len(my_events)
    t_gen = time.perf_counter() - t_setup0

    def batch_views(e0, e1):
        """Events [e0, e1) of this rank as (pieces CellSet, cells CellSet, ev_piece_off, ev_cell_off)."""
        p = synth.slice_sets(pieces, int(ev_p[e0]), int(ev_p[e1]))
        c = synth.slice_sets(cells, int(ev_c[e0]), int(ev_c[e1]))
        return p, c, (ev_p[e0:e1 + 1] - ev_p[e0]).astype(np.uint32), (ev_c[e0:e1 + 1] - ev_c[e0]).astype(np.uint32)

    # ---- resident batches: one context each, inputs stay in HBM ----
    class Batch:
        pass

    res = []
    for bi, e0 in enumerate(range(0, n_my, args.resident_batch)):
        b = Batch()
        b.e0, b.e1 = e0, min(n_my, e0 + args.resident_batch)
        b.cx = FractureContext(local, streams[bi % 2].cuda_stream)
        b.cx.set_kdop_directions(args.kdop)
        p, c, evp, evc = batch_views(b.e0, b.e1)
        b.cx.upload_pieces(p.verts, p.vert_off, p.ring_off, p.ring, evp)
        b.cx.upload_cells(c.planes, c.plane_off, c.verts, c.vert_off, evc)
        b.cx.fracture_event()
        b.counts = b.cx.counts()
        rec = b.cx.download(geometry=False).rec
        b.n_frag = len(rec)
        b.alg_bytes = synth.algorithmic_bytes(p.vert_off, p.ring_off, c.plane_off, rec)
        b.alg_flops = synth.algorithmic_flops(p.vert_off, c.plane_off, rec)
        # per-event output sizes (the end-to-end loop sizes its pinned output buffers from them)
        ev_of = (rec["cell"] // CELLS_PER_EVENT).astype(np.int64)
        ne = b.e1 - b.e0
        b.ev_frags = np.bincount(ev_of, minlength=ne)
        b.ev_verts = np.bincount(ev_of, weights=rec["n_verts"], minlength=ne).astype(np.int64)
        b.ev_ring = np.bincount(ev_of, weights=rec["n_ring"], minlength=ne).astype(np.int64)
        b.in_bytes = sum(a.nbytes for a in (p.verts, p.vert_off, p.ring_off, p.ring, c.planes, c.plane_off, c.verts, c.vert_off))
        res.append(b)
    torch.cuda.synchronize()
    frags_rank = sum(b.n_frag for b in res)
    cand_rank = sum(int(b.counts.n_candidates) for b in res)
    pairs_rank = sum(int(b.counts.n_pairs) for b in res)
    resident_in_bytes = sum(b.in_bytes for b in res)

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        time.sleep(0.3)

    def one_pass():
        n = 0
        for b in res:
            b.cx.fracture_event()
            n += b.cx.last_event_launches()
        return n

    def join_streams():
        done = torch.cuda.Event()
        done.record(streams[1])
        streams[0].wait_event(done)

    # ---- warm-up, then the timed region: K passes over the rank's resident events ----
    for _ in range(args.warmup):
        one_pass()
    barrier()
    mark0 = sampler.mark()
    t_wall0 = time.perf_counter()
    t0_ev, t1_ev = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0_ev.record(streams[0])
    streams[1].wait_event(t0_ev)
    launches = 0
    for _ in range(args.steps):
        launches += one_pass()
    join_streams()
    t1_ev.record(streams[0])
    torch.cuda.synchronize()
    t_wall = time.perf_counter() - t_wall0
    mark1 = sampler.mark()
    if world > 1:
        dist.barrier()
    for b in res:
        assert int(b.cx.counts().n_fragments) == b.n_frag
    total_ms = allreduce(t0_ev.elapsed_time(t1_ev), dist.ReduceOp.MAX)
    frags_job = allreduce(frags_rank, dist.ReduceOp.SUM)
    cand_job = allreduce(cand_rank, dist.ReduceOp.SUM)
    pairs_job = allreduce(pairs_rank, dist.ReduceOp.SUM)
    value = frags_job * args.steps / (total_ms * 1e-3)

    # ---- per-kernel durations (CUDA events between the kernels: they serialise the PDL chain, so a separate pass) ----
    phases = {}
    for b in res:
        b.cx.set_profiling(True)
        b.cx.fracture_event()
        for k, v in b.cx.last_event_phases().items():
            phases[k] = phases.get(k, 0.0) + v
        b.cx.set_profiling(False)
    torch.cuda.synchronize()
    k3_ms = phases["k3_clip_small"]
    alg_bytes = sum(b.alg_bytes for b in res)
    alg_flops = sum(b.alg_flops for b in res)

    # ---- e2e: host buffers in, host buffers out, every step ----
    e2e = bench_secondary.config4_e2e(args, torch, dev, local, rank, world, streams, batch_views, res, n_my, barrier, allreduce, dist)

    # ---- the final fragment gather (the only collective; after the hot path, reported separately) ----
    gather = bench_secondary.final_gather(torch, dev, rank, world, res, barrier, allreduce, dist) if world > 1 else None

    # ---- secondary configs (rank 0 drives configs 2 and 3: a single event is never split; config 5 is sharded) ----
    secondary = {}
    if not args.no_secondary:
        for b in res:          # free the resident batches' memory first
            b.cx.close()
        res_closed = True
        secondary = bench_secondary.run_all(args, torch, dev, local, rank, world, streams, barrier, allreduce, dist)
    else:
        res_closed = False

    dma = bench_secondary.dma_ceiling(torch, dev, world, barrier, allreduce, dist)
    clocks = sampler.stop(mark0, max(mark1, mark0 + 1)) if rank == 0 else None
    if clocks is not None:
        clocks["window"] = "the timed region (K passes over the resident events)"

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        achieved = alg_bytes / (k3_ms * 1e-3) / 1e9
        traffic = bench_secondary.k3_traffic_from_profile(alg_bytes)
        step_ms_rank = sum(phases.values())
        kernels = {k: {"ms": round(v, 4), "share_of_step": round(v / step_ms_rank, 4)} for k, v in phases.items()}
        roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                    "traffic": traffic["bytes_per_launch"] if traffic else None,
                    "kernel": "clip_fast_kernel<2> (K3, small tier: one warp per candidate pair, resident warps)",
                    "kernel_ms": k3_ms, "launches": len(res),
                    "algorithmic_bytes_per_launch": alg_bytes / len(res),
                    "units_per_launch": f"{frags_rank // len(res)} surviving pairs x (16 V_in + 4 E2_in + 16 P + 16 V_out + 4 E2_out + 64) bytes",
                    "regime": "kernel_ms = sum over the rank's resident batches of the K3 small-tier duration, CUDA events between the "
                              "kernels of each batch in a separate pass (same launches, same resident inputs as the timed region)",
                    "traffic_note": traffic["note"] if traffic else "no ncu capture of the current kernel source under profiles/",
                    "peak_source": "MEASURED_PEAKS.json hbm_gbs (measured)" if peaks else "fallback 6650 GB/s"}
        fp32_peak = secondary.get("fp32_peak_tflops")
        ms_per_step = total_ms / args.steps
        line = {
            "metric": "clipped fragments/sec", "value": value, "unit": "fragments/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args.events),
            "run": {"fragments_per_step": int(frags_job), "pairs_per_step": int(pairs_job), "candidates_per_step": int(cand_job),
                    "kdop_directions": args.kdop, "events_per_rank": n_my, "resident_batch_events": args.resident_batch,
                    "resident_batches_per_rank": len(res), "resident_input_mib_per_rank": round(resident_in_bytes / 2 ** 20, 1),
                    "l2": f"inputs larger than L2: {resident_in_bytes / 2 ** 20:.0f} MiB of resident inputs per rank (plus scratch and outputs) "
                          f"stream through a {L2_MIB} MB L2 every step; no flush kernel in the timed region",
                    "timing": "CUDA events on stream 0 around K passes (start gates the second stream, stop joins it), max over ranks",
                    "input_build_s": round(t_gen, 2), "numa": numa},
            "wall_s_timed_region": t_wall,
            "e2e": e2e, "gpu_launches": int(launches), "launches_per_step": int(launches // max(1, args.steps)),
            "clocks": clocks, "roofline": roofline, "kernels": kernels,
            "fp32": {"algorithmic_flops_per_step_rank0": alg_flops, "achieved_gflops": alg_flops / (ms_per_step * 1e-3) / 1e9,
                     "peak_tflops_measured": fp32_peak,
                     "frac": alg_flops / (ms_per_step * 1e-3) / 1e12 / fp32_peak if fp32_peak else None},
            "dma_ceiling": dma,
        }
        if gather:
            line["gather"] = gather
        for k in ("config3", "config2", "config5"):
            if k in secondary:
                line[k] = secondary[k]
        if world == 1 and not args.no_cpu_baseline:
            from oracle.refapi import PolySet
            sample = []
            for e in range(min(args.ref_sample, n_my)):
                p, c, _, _ = batch_views(e, e + 1)
                ps = PolySet(p.verts, p.vert_off, p.ring_off, p.ring)
                cs = PolySet(c.verts, c.vert_off, c.ring_off, c.ring)
                cs.planes, cs.poly_face_off = c.planes, c.plane_off
                sample.append((ps, cs))
            v, kind, cores, passes, secs = time_reference(sample, REF_THREADS, args.cpu_budget)
            line["cpu_baseline"] = {"value": v, "unit": "fragments/s", "cores": cores, "kind": kind,
                                    "sample": f"{passes} passes over events 0..{len(sample) - 1} of the workload, {secs:.2f} s inside the reference "
                                              f"build's own timer (fan-out + SetExtract), dp::thread_pool({REF_THREADS}); {os.cpu_count()} host threads available"}
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if not res_closed:
        for b in res:
            b.cx.close()


if __name__ == "__main__":
    main()
