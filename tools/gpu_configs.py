"""Dev/measurement: BASELINE configs 3, 4, 5 at (or near) full size on one GPU: parity fingerprints + timings."""
import json, sys, time, numpy as np
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import torch
from surtr_b200 import FractureContext, synth
import common

out = {}
ctx = FractureContext(0)
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
                  "fingerprint_matches_reference": common.summary_of_fragments(fr) == json.load(open("tests/golden/summaries.json"))["config3_10000x256"]}
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
                  "event0_fragments": per_event[0]}
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
out["config5"] = {"objects": n_obj, "fragments_per_level": counts, "per_object": [x // n_obj for x in counts],
                  "sum_event_ms": tot_ms, "final_fragments_per_s": counts[-1] / (tot_ms * 1e-3)}
print(out["config5"], flush=True)
json.dump(out, open("gpurun_out/configs_r1.json", "w"), indent=1)
