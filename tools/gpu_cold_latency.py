"""Cold single-event latency (L2 flushed before every event, CUDA events around the event) of BASELINE configs 2 and 3:
the p50 the bench reports, without the rest of the bench.  Used to A/B library variants."""
import json, sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from surtr_b200 import FractureContext, synth
import bench_secondary as B

dev = torch.device("cuda:0")
stream = torch.cuda.Stream(device=dev)
flush = torch.empty(B.FLUSH_MIB * 2 ** 20, dtype=torch.uint8, device=dev)
out = {}
with torch.cuda.stream(stream):
    ctx = FractureContext(0, stream.cuda_stream)
    cells = synth.voronoi_cells_batch(ctx, synth.seeds_uniform(46354, 4096), np.array([0, 4096], np.uint32))
    cv, cvo, cro, cr = synth.unit_cube()
    ctx.set_kdop_directions(3)
    ctx.upload_pieces(cv, cvo, cro, cr)
    ctx.upload_cells(cells.planes, cells.plane_off, cells.verts, cells.vert_off)
    for _ in range(5):
        ctx.fracture_event()
    ctx.counts()
    ms = B._latency(torch, ctx, flush, stream, 100)
    out["config2_p50_ms"] = float(np.median(ms)); out["config2_min_ms"] = float(np.min(ms))
    ctx.set_kdop_directions(13)
    pieces = synth.voronoi_cells_batch(ctx, synth.seeds_uniform(1234, 10000), np.array([0, 10000], np.uint32), planes=False)
    cells = synth.voronoi_cells_batch(ctx, synth.seeds_uniform(46354, 256), np.array([0, 256], np.uint32))
    ctx.upload_pieces(pieces.verts, pieces.vert_off, pieces.ring_off, pieces.ring)
    ctx.upload_cells(cells.planes, cells.plane_off, cells.verts, cells.vert_off)
    for _ in range(5):
        ctx.fracture_event()
    ctx.counts()
    ms = B._latency(torch, ctx, flush, stream, 100)
    out["config3_p50_ms"] = float(np.median(ms)); out["config3_min_ms"] = float(np.min(ms))
print(json.dumps(out))
