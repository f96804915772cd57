"""Measurement: broad-phase direction set k in {3, 7, 13} (AABB, 14-DOP, 26-DOP) on BASELINE configs 3 and 4 -- candidates
that reach the clipper, event time, K3 time.  More directions cost more in K1 / K2 and cull more dead pairs before K3.
Writes gpurun_out/kdop_sweep.json."""
import json, sys, numpy as np
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')   # run from the repo root
from surtr_b200 import FractureContext
import common

out = {}
ctx = FractureContext(0)


def sweep(name, pieces, cells, ev_p=None, ev_c=None, reps=60):
    rows = []
    ref = None
    for k in (3, 7, 13):
        ctx.set_kdop_directions(k)
        fr = common.run_gpu(ctx, pieces, cells, ev_p, ev_c)
        if ref is None:
            ref = fr.rec.tobytes()
        assert fr.rec.tobytes() == ref, "the direction set changed the fragments"
        for _ in range(3):
            ctx.fracture_event(); ctx.counts()
        tot, clip = [], []
        ctx.set_profiling(True)
        for _ in range(10):
            ctx.fracture_event(); ctx.counts(); clip.append(ctx.last_event_ms()[1])
        ctx.set_profiling(False)
        for _ in range(reps):
            ctx.fracture_event(); ctx.counts(); tot.append(ctx.last_event_ms()[0])
        c = ctx.counts()
        rows.append({"k": k, "pairs": int(c.n_pairs), "candidates": int(c.n_candidates), "fragments": int(c.n_fragments),
                     "event_ms_p50": float(np.median(tot)), "event_ms_min": float(np.min(tot)), "k3_ms": float(np.median(clip))})
        print(name, rows[-1], flush=True)
    out[name] = rows


sweep("config3_10000x256", common.voronoi(1234, 10000), common.voronoi(46354, 256))
n_ev = 64
psets = [common.voronoi(1234 + e % 8, 1000) for e in range(n_ev)]
csets = [common.voronoi(46354 + e % 8, 64) for e in range(n_ev)]
pieces, ev_p = common.concat(psets)
cells, ev_c = common.concat(csets)
sweep("config4_64_events_1000x64", pieces, cells, ev_p, ev_c, reps=30)
json.dump(out, open("gpurun_out/kdop_sweep.json", "w"), indent=1)
