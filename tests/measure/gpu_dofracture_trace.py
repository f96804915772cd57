"""Dev: phase times of SurtrHost::DoFracture on the bunny compound (SURTR_TRACE=1)."""
import os, sys, time, numpy as np
os.environ["SURTR_TRACE"] = "1"
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')   # run from the repo root
import hostapi
from test_oracle_port import load_polyset
d1 = np.load("tests/golden/config1_full_bunny32.npz"); dd = np.load("tests/golden/do_fracture_bunny.npz")
cvx, msh = load_polyset(d1, "convex_"), load_polyset(d1, "mesh_")
for mode in ("general", "partial"):
    a = (cvx, msh, dd[mode + "_seeds"], dd["cloud"], dd["impact"], float(dd[mode + "_radius"]), float(dd["max_axis_scale"]), mode == "partial")
    for rep in range(3):
        print("----", mode, rep, file=sys.stderr, flush=True)
        t0 = time.perf_counter(); hostapi.do_fracture(*a); print("total ms", 1e3 * (time.perf_counter() - t0), "DoFracture itself ms", hostapi.last_do_fracture_ms(), file=sys.stderr, flush=True)
