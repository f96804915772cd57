"""Measurement: the bunny mesh polyhedron (2503 vertices) x 32 cells, the global-memory tier of K3, for ncu."""
import sys, numpy as np
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')   # run from the repo root
from surtr_b200 import FractureContext
from test_oracle_port import load_polyset
m = np.load("tests/golden/bunny_mesh_x32.npz")
mesh = load_polyset(m, "mesh_")
ctx = FractureContext(0)
ctx.upload_pieces(mesh.verts, mesh.vert_off, mesh.ring_off, mesh.ring)
ctx.upload_cells(m["planes"], m["plane_off"], m["cell_verts"], m["cell_vert_off"])
for _ in range(4):
    ctx.fracture_event(); c = ctx.counts()
print("tier3 pairs", c.n_tier3, "fragments", c.n_fragments, "event ms", ctx.last_event_ms())
