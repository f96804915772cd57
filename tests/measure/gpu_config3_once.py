"""Measurement: BASELINE config 3 (10 000 pieces x 256 cells), a few events, for an ncu launch list."""
import sys
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')   # run from the repo root
import common
from surtr_b200 import FractureContext
ctx = FractureContext(0)
pieces, cells = common.voronoi(1234, 10000), common.voronoi(46354, 256)
fr = common.run_gpu(ctx, pieces, cells)
for _ in range(5):
    ctx.fracture_event()
c = ctx.counts()
print("candidates", c.n_candidates, "fragments", c.n_fragments, "event ms", ctx.last_event_ms())
