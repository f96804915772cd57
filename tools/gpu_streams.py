"""Dev/measurement: resident-input throughput of the bench workload with D events in flight (D contexts = D streams)."""
import sys, time, numpy as np, torch
sys.path.insert(0, '.')
from surtr_b200 import FractureContext, synth

dev = torch.device("cuda", 0)
N = 4096
base = FractureContext(0)
cells = synth.voronoi_cells(base, synth.seeds_uniform(46354, N))
cube = synth.unit_cube()
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
K = 400
for D in (1, 2, 3, 4, 6, 8):
    for with_flush in (True, False):
        pipes = []
        for d in range(D):
            st = torch.cuda.Stream(device=dev)
            cx = FractureContext(0, st.cuda_stream)
            cx.upload_pieces(*cube)
            cx.upload_cells(cells.planes, cells.plane_off, cells.verts, cells.vert_off)
            cx.fracture_event(); cx.counts()
            pipes.append((cx, st))
        torch.cuda.synchronize()
        s0 = pipes[0][1]
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(s0)
        for cx, st in pipes[1:]:
            st.wait_event(e0)
        for i in range(K):
            cx, st = pipes[i % D]
            if with_flush:
                with torch.cuda.stream(st):
                    flush.zero_()
            cx.fracture_event()
        for cx, st in pipes[1:]:
            e = torch.cuda.Event(); e.record(st); s0.wait_event(e)
        e1.record(s0)
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / K
        print(f"D={D} flush={with_flush}: {ms*1e3:.1f} us/event  {N/ms/1e3:.2f} M frag/s", flush=True)
        for cx, st in pipes:
            cx.counts(); cx.close()
