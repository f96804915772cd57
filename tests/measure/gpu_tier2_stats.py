"""Measurement: how many pairs of configs 3 / 4 leave the small K3 tier, and what the large tier costs per event."""
import sys, numpy as np
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')   # run from the repo root
import common
from surtr_b200 import FractureContext
ctx = FractureContext(0)
for name, pieces, cells in (("config3", common.voronoi(1234, 10000), common.voronoi(46354, 256)),
                            ("config4 e0", common.voronoi(1234, 1000), common.voronoi(46354, 64)),
                            ("config2", common.unit_cube(), common.voronoi(46354, 4096))):
    fr = common.run_gpu(ctx, pieces, cells)
    c = ctx.counts()
    ts = []
    for _ in range(20):
        ctx.fracture_event(); ctx.counts(); ts.append(ctx.last_event_ms()[0])
    nvp = np.diff(pieces.vert_off)
    print(name, "candidates", c.n_candidates, "tier2", c.n_tier2, "tier3", c.n_tier3, "seq cuts", c.n_seq_cuts, "max piece verts", int(nvp.max()),
          "max fragment verts", int(fr.rec["n_verts"].max()), "event ms", float(np.median(ts)))
