"""Measurement: the launches an ncu capture should see, and nothing else inside the profiled range.

  ncu --profile-from-start off ... python tools/gpu_profile_workloads.py [config4|config3|config2|mesh] [reps]

Builds the workload (its own launches happen before cudaProfilerStart), warms the context up, then runs `reps` events
between cudaProfilerStart / cudaProfilerStop.  config4 = one batch of 64 independent events (1000 x 64 each), the unit of
bench.py's end-to-end loop; config3 = the 10 000 x 256 event; config2 = unit cube x 4096 cells; mesh = the 2503-vertex
bunny mesh x 32 cells (global tier).  Prints the event's counters and the algorithmic bytes of its K3 launch.
BLOB=1: the profiled events go through the one-copy wire format (surtr_upload_blob -> event -> surtr_download_blob_async,
pinned host buffers), so expand_blob_kernel and pack_blob_kernel are in the capture too, and the line carries the
algorithmic bytes of EVERY kernel (the formulas of DESIGN.md section 4) for the per-kernel roofline table."""
import json, os, sys
import numpy as np
sys.path.insert(0, '.')
import torch
from surtr_b200 import FractureContext, synth

what = sys.argv[1] if len(sys.argv) > 1 else "config4"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
dev = torch.device("cuda", 0)
st = torch.cuda.Stream(device=dev)
torch.cuda.set_stream(st)
ctx = FractureContext(0, st.cuda_stream)
flush = torch.empty(160 << 20, dtype=torch.uint8, device=dev)
if what == "config4":
    n_ev = int(os.environ.get("EVENTS", "64"))
    pieces, cells, ev_p, ev_c = synth.config4_events(ctx, range(n_ev))
    ctx.upload_pieces(pieces.verts, pieces.vert_off, pieces.ring_off, pieces.ring, ev_p)
    ctx.upload_cells(cells.planes, cells.plane_off, cells.verts, cells.vert_off, ev_c)
elif what == "config3":
    pieces = synth.voronoi_cells_batch(ctx, synth.seeds_uniform(1234, 10000), np.array([0, 10000], np.uint32), planes=False)
    cells = synth.voronoi_cells_batch(ctx, synth.seeds_uniform(46354, 256), np.array([0, 256], np.uint32))
    ctx.upload_pieces(pieces.verts, pieces.vert_off, pieces.ring_off, pieces.ring)
    ctx.upload_cells(cells.planes, cells.plane_off, cells.verts, cells.vert_off)
elif what == "config2":
    cells = synth.voronoi_cells_batch(ctx, synth.seeds_uniform(46354, 4096), np.array([0, 4096], np.uint32))
    cv, cvo, cro, cr = synth.unit_cube()
    pieces = synth.CellSet(cv, cvo, cro, cr, np.zeros((0, 4), np.float32), np.zeros(2, np.uint32))
    ctx.set_kdop_directions(3)
    ctx.upload_pieces(cv, cvo, cro, cr)
    ctx.upload_cells(cells.planes, cells.plane_off, cells.verts, cells.vert_off)
else:
    d = np.load("tests/golden/bunny_mesh_x32.npz")
    pieces = synth.CellSet(d["mesh_verts"], d["mesh_vert_off"], d["mesh_ring_off"], d["mesh_ring"], np.zeros((0, 4), np.float32), np.zeros(2, np.uint32))
    cells = synth.CellSet(d["cell_verts"], d["cell_vert_off"], None, None, d["planes"], d["plane_off"])
    ctx.upload_pieces(pieces.verts, pieces.vert_off, pieces.ring_off, pieces.ring)
    ctx.upload_cells(cells.planes, cells.plane_off, cells.verts, cells.vert_off)
for _ in range(3):
    ctx.fracture_event()
c = ctx.counts()
rec = ctx.download(geometry=False).rec
alg = synth.algorithmic_bytes(pieces.vert_off, pieces.ring_off, cells.plane_off, rec)
ctx.set_profiling(True)
flush.zero_()
ctx.fracture_event()
ph = ctx.last_event_phases()
ctx.set_profiling(False)
blob = os.environ.get("BLOB", "0") == "1" and what != "mesh"
per_kernel = None
if blob:
    evp = ev_p if what == "config4" else None
    evc = ev_c if what == "config4" else None
    sizes, total = FractureContext.fill_input_blob(None, pieces, cells, evp, evc)
    h_in = torch.empty(total, dtype=torch.uint8, pin_memory=True)
    FractureContext.fill_input_blob(h_in.numpy(), pieces, cells, evp, evc)
    al = lambda x: (int(x) + 255) // 256 * 256
    cap = al(64 * c.n_fragments) + al(12 * c.n_verts) + al(c.n_verts) + al(2 * c.n_ring)
    h_out = torch.empty(cap, dtype=torch.uint8, pin_memory=True)
    ctx.upload_blob_ptr(h_in.data_ptr(), sizes); ctx.fracture_event(); ctx.download_blob_into_async(h_out.data_ptr(), cap); ctx.sync()
    k = 3 if what == "config2" else 13
    nP, nC = len(pieces.vert_off) - 1, len(cells.plane_off) - 1
    nVp, nVc, nPl = len(pieces.verts), len(cells.verts), len(cells.planes)
    S, F, NV, NE = int(c.n_candidates), int(c.n_fragments), int(c.n_verts), int(c.n_ring)
    words = int(c.n_pairs) // 32
    per_kernel = {   # compulsory bytes per launch (read + write), DESIGN.md section 4
        "expand_blob_kernel": 28 * (nVp + nVc) + 5 * nVp + 8 * nP + 3 * len(pieces.ring),
        "kdop_extents_kernel": 16 * (nVp + nVc) + 8 * k * (nP + nC),
        "broadphase_mask_kernel": 8 * k * (nP + nC) + 4 * words,
        "compact_pairs_kernel": 4 * words + 8 * S,
        "clip_fast_kernel": int(alg),
        "assemble_scan_kernel": 16 * S + 16 * S + 4 * F,
        "assemble_gather_kernel": F * 64 + NV * (16 + 2) + NE + NV * (16 + 4) + 2 * NE,
        "pack_blob_kernel": F * 64 * 2 + NV * (16 + 4 + 12 + 1) + NE * 3,
    }
torch.cuda.synchronize()
torch.cuda.profiler.start()
for _ in range(reps):
    flush.zero_()
    if blob:
        ctx.upload_blob_ptr(h_in.data_ptr(), sizes)
    ctx.fracture_event()
    if blob:
        ctx.download_blob_into_async(h_out.data_ptr(), cap)
        ctx.sync()
    else:
        ctx.counts()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print(json.dumps({"workload": what, "pairs": int(c.n_pairs), "candidates": int(c.n_candidates), "fragments": int(c.n_fragments),
                  "seq_cuts": int(c.n_seq_cuts), "tier1b": int(c.n_tier1b), "tier2": int(c.n_tier2), "tier3": int(c.n_tier3),
                  "k3_algorithmic_bytes": int(alg), "kernel_ms_unprofiled": ph, "k3": os.environ.get("SURTR_K3", "fast"),
                  "algorithmic_bytes_per_kernel": per_kernel}))
