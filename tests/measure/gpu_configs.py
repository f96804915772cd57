"""Measurement: BASELINE configs 1, 3, 4, 5 at (or near) full size on one GPU -- parity fingerprints, GPU timings and the
reference build's CPU timings (oracle/_ref, dp::thread_pool(16) fan-out + SetExtract, the reference's own timer) on
the same box.  Writes gpurun_out/configs_r2.json (copied to profiles/)."""
import json, os, sys, time, numpy as np
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')   # run from the repo root
import torch
from surtr_b200 import FractureContext, synth
import common
from oracle import refapi as R

HAVE_REF = R.available()
CPU_THREADS = 16   # dp::thread_pool(16), Src/Surtr.cpp:28


def cpu_seconds(pieces, cells, reps=1):
    """Reference CPU time of one event (fan-out + SetExtract), best of reps."""
    if not HAVE_REF:
        return None
    return min(R.apply_fracture(pieces, cells.planes, cells.plane_off, CPU_THREADS, moments=False).seconds for _ in range(reps))

out = {"host": {"nproc": os.cpu_count(), "cpu_threads_used": CPU_THREADS, "reference_build": HAVE_REF}}
try:
    out["host"]["cpu_model"] = [l.split(":")[1].strip() for l in open("/proc/cpuinfo") if l.startswith("model name")][0]
except Exception:
    pass
ctx = FractureContext(0)

# config 1: bunny, 32 seeds, full PrepareFracture through the host classes (two GPU events + refit) vs the restatement
# over the reference build
import hostapi
d1 = np.load("tests/golden/config1_full_bunny32.npz")
hostapi.config1_full(d1["verts"], d1["indices"], d1["seeds"])            # warm-up (context creation, tier enabling)
ts = []
for _ in range(5):
    t0 = time.perf_counter(); c1, m1, _ = hostapi.config1_full(d1["verts"], d1["indices"], d1["seeds"]); ts.append(time.perf_counter() - t0)
off1, idx1 = hostapi.dt3d_neighbors(d1["seeds"])
cpu1 = None
if HAVE_REF:
    t0 = time.perf_counter(); R.config1_full(d1["verts"], d1["indices"], d1["seeds"], off1, idx1); cpu1 = time.perf_counter() - t0
out["config1"] = {"pieces": c1.n, "host_classes_wall_ms": 1e3 * float(np.median(ts)), "reference_cpu_wall_ms": None if cpu1 is None else 1e3 * cpu1,
                  "note": "wall clock of the whole PrepareFracture (ICH, k-DOP, ACH, mesh rings, DT3D cells, convex + mesh clip, islands, refit, extract); reference: single thread, inline"}
print(out["config1"], flush=True)

# config 1, second stage: the reference's real event path, one DoFracture (32-cell radial pattern at an impact point) on
# the 27-piece compound PrepareFracture produced
from test_oracle_port import load_polyset
dd = np.load("tests/golden/do_fracture_bunny.npz")
cvx, msh = load_polyset(d1, "convex_"), load_polyset(d1, "mesh_")
for mode in ("general", "partial"):
    a = (cvx, msh, dd[mode + "_seeds"], dd["cloud"], dd["impact"], float(dd[mode + "_radius"]), float(dd["max_axis_scale"]), mode == "partial")
    hostapi.do_fracture(*a)
    ts, inner = [], []
    for _ in range(5):
        t0 = time.perf_counter(); r = hostapi.do_fracture(*a); ts.append(time.perf_counter() - t0); inner.append(hostapi.last_do_fracture_ms())
    cpu = None
    if HAVE_REF:
        o, i = hostapi.dt3d_neighbors(dd[mode + "_seeds"])
        t0 = time.perf_counter()
        R.do_fracture(cvx, msh, dd[mode + "_seeds"], o, i, dd["cloud"], dd["impact"], float(dd[mode + "_radius"]), float(dd["max_axis_scale"]), mode == "partial")
        cpu = time.perf_counter() - t0
    out["config1_do_fracture_" + mode] = {"pieces": r[0].n, "compounds": r[2], "host_classes_wall_ms": 1e3 * float(np.median(ts)), "do_fracture_call_ms": float(np.median(inner)),
                                          "reference_cpu_wall_ms": None if cpu is None else 1e3 * cpu}
    print(mode, out["config1_do_fracture_" + mode], flush=True)
def timed_events(n=30):
    ts = []
    for _ in range(n):
        ctx.fracture_event(); ctx.counts(); ts.append(ctx.last_event_ms()[0])
    return float(np.median(ts)), float(np.min(ts))

# config 3: 10000 pieces x 256 cells, single event latency
pieces, cells = common.voronoi(1234, 10000), common.voronoi(46354, 256)
fr = common.run_gpu(ctx, pieces, cells)
p50, best = timed_events(100)
c = ctx.counts()
out["config3"] = {"pieces": 10000, "cells": 256, "pairs": int(c.n_pairs), "candidates": int(c.n_candidates), "fragments": fr.n,
                  "p50_event_ms": p50, "min_event_ms": best, "fragments_per_s": fr.n / (p50 * 1e-3),
                  "fingerprint_matches_reference": common.summary_of_fragments(fr) == json.load(open("tests/golden/summaries.json"))["config3_10000x256"],
                  "reference_cpu_event_ms": None if not HAVE_REF else 1e3 * cpu_seconds(pieces, cells, 2)}
print(out["config3"], flush=True)

# config 4: N independent events (1000 pieces x 64 cells each) in one batch on one GPU
n_ev = int(sys.argv[1]) if len(sys.argv) > 1 else 256
t0 = time.time()
base_p = [common.voronoi(1234 + e, 1000) for e in range(8)]
base_c = [common.voronoi(46354 + e, 64) for e in range(8)]
psets = [base_p[e % 8] for e in range(n_ev)]       # 8 distinct events tiled to n_ev (host build time bound)
csets = [base_c[e % 8] for e in range(n_ev)]
pieces, ev_p = common.concat(psets)
cells, ev_c = common.concat(csets)
print("built", n_ev, "events in", time.time() - t0, "s", flush=True)
fr = common.run_gpu(ctx, pieces, cells, ev_p, ev_c)
p50, best = timed_events(10)
c = ctx.counts()
per_event = [common.run_gpu(FractureContext(0), base_p[e], base_c[e]).n for e in range(2)]
out["config4"] = {"events": n_ev, "pairs": int(c.n_pairs), "candidates": int(c.n_candidates), "fragments": fr.n,
                  "p50_batch_ms": p50, "fragments_per_s": fr.n / (p50 * 1e-3), "events_per_s": n_ev / (p50 * 1e-3),
                  "event0_fragments": per_event[0],
                  "reference_cpu_event0_ms": None if not HAVE_REF else 1e3 * cpu_seconds(base_p[0], base_c[0], 3)}
print(out["config4"], flush=True)

# config 5: depth-3 recursion for many objects at once (objects = events; fragments stay on the device)
n_obj = int(sys.argv[2]) if len(sys.argv) > 2 else 256
cube = common.unit_cube()
levels = common.recursion_levels()
pieces, ev_p = common.concat([cube] * n_obj)
ctx2 = FractureContext(0)
ctx2.upload_pieces(pieces.verts, pieces.vert_off, pieces.ring_off, pieces.ring, ev_p)
tot_ms, counts = 0.0, []
for lvl, cells in enumerate(levels):
    cl, ev_c = common.concat([cells] * n_obj)
    ctx2.upload_cells(cl.planes, cl.plane_off, cl.verts, cl.vert_off, ev_c)
    ctx2.fracture_event()
    cc = ctx2.counts()
    tot_ms += ctx2.last_event_ms()[0]
    counts.append(int(cc.n_fragments))
    rec = ctx2.download(geometry=False).rec
    # regroup the fragments by object for the next level: fragments are event-major already
    ev_of_frag = np.searchsorted(ev_c, rec["cell"], side="right") - 1
    new_ev = np.concatenate([[0], np.cumsum(np.bincount(ev_of_frag, minlength=n_obj))]).astype(np.uint32)
    ctx2.fragments_to_pieces(new_ev)
cpu5 = None
if HAVE_REF:
    cpu5, pcs = 0.0, cube
    for cells in levels:
        r = R.apply_fracture(pcs, cells.planes, cells.plane_off, CPU_THREADS, moments=False)
        cpu5 += r.seconds
        pcs = r
out["config5"] = {"objects": n_obj, "reference_cpu_per_object_ms": None if cpu5 is None else 1e3 * cpu5, "fragments_per_level": counts, "per_object": [x // n_obj for x in counts],
                  "sum_event_ms": tot_ms, "final_fragments_per_s": counts[-1] / (tot_ms * 1e-3)}
print(out["config5"], flush=True)
json.dump(out, open("gpurun_out/configs_r2.json", "w"), indent=1)
