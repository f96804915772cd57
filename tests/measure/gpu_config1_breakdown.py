"""Dev: where the 13 ms of config 1 (full PrepareFracture through the host classes) go."""
import sys, time, numpy as np
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')   # run from the repo root
import hostapi as H, common
from surtr_b200 import FractureContext
from test_oracle_port import load_polyset
d = np.load("tests/golden/config1_full_bunny32.npz")
def t(f, n=5):
    f(); ts = []
    for _ in range(n):
        t0 = time.perf_counter(); f(); ts.append(time.perf_counter() - t0)
    return 1e3 * float(np.median(ts))
print("full PrepareFracture", t(lambda: H.config1_full(d["verts"], d["indices"], d["seeds"])))
print("convex-only PrepareFracture", t(lambda: H.config1(d["verts"], d["seeds"])))
print("mesh rings (ExtractNeighborFromMesh)", t(lambda: H.mesh_polyhedron(d["verts"], d["indices"])))
print("DT3D neighbours", t(lambda: H.dt3d_neighbors(d["seeds"])))
print("ICH normals (limit 20)", t(lambda: H.ich_normals(d["verts"], 20)))
m = np.load("tests/golden/bunny_mesh_x32.npz")
mesh = load_polyset(m, "mesh_")
ctx = FractureContext(0)
ctx.upload_pieces(mesh.verts, mesh.vert_off, mesh.ring_off, mesh.ring)
ctx.upload_cells(m["planes"], m["plane_off"], m["cell_verts"], m["cell_vert_off"])
ctx.fracture_event(); ctx.counts()
def ev():
    ctx.fracture_event(); ctx.counts()
print("mesh x 32 cells event (wall)", t(ev), "device ms", ctx.last_event_ms())
def dl():
    ctx.download()
print("download", t(dl))
c = ctx.counts(); print("candidates", c.n_candidates, "tier3", c.n_tier3, "fragments", c.n_fragments, "verts", c.n_verts)
