import sys, time, numpy as np
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')   # run from the repo root
from oracle import refapi as R, portapi as P
from surtr_b200 import FractureContext
import common
ctx = FractureContext(0)
cube = R.unit_cube()
for name, pieces, cells in [
    ("cube x 64", cube, common.voronoi(46354, 64)),
    ("cube x 4096", cube, common.voronoi(46354, 4096)),
    ("1000 x 64", common.voronoi(1234, 1000), common.voronoi(46354, 64)),
    ("10000 x 256", common.voronoi(1234, 10000), common.voronoi(46354, 256)),
]:
    want = P.apply_fracture(pieces, cells.planes, cells.plane_off)
    got = common.run_gpu(ctx, pieces, cells)
    c = ctx.counts()
    print(name, "frags gpu", got.n, "oracle", want.n, "cands", c.n_candidates, "pairs", c.n_pairs, "seq", c.n_seq_cuts, "tier2", c.n_tier2, "ms", ctx.last_event_ms(), "launches", ctx.last_event_launches())
    try:
        common.assert_fragments_equal(got, want)
        print("   BIT-EXACT (assignments, V/F, positions, rings, volume, centroid)")
    except AssertionError as e:
        print("   MISMATCH:", e)
        n = min(got.n, want.n)
        bad = np.nonzero((got.rec["n_verts"][:n] != want.nverts[:n]) | (got.rec["n_faces"][:n] != want.nfaces[:n]))[0]
        print("   VF mismatches:", len(bad), bad[:10])
    for rep in range(3):
        ctx.fracture_event(); ctx.counts(); print("   rerun ms", ctx.last_event_ms())
