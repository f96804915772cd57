"""Measurement: phase times of the full PrepareFracture on the bunny (SURTR_TRACE=1)."""
import os, sys, time, numpy as np
os.environ["SURTR_TRACE"] = "1"
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')   # run from the repo root
import hostapi
d = np.load("tests/golden/config1_full_bunny32.npz")
for rep in range(3):
    print("---- rep", rep, file=sys.stderr, flush=True)
    t0 = time.perf_counter(); hostapi.config1_full(d["verts"], d["indices"], d["seeds"]); print("total ms", 1e3 * (time.perf_counter() - t0), file=sys.stderr, flush=True)
