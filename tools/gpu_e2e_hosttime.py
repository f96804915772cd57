"""Dev/measurement: where the host thread's time goes in the pipelined end-to-end loop of bench.py."""
import sys, time, collections, numpy as np, torch
sys.path.insert(0, '.')
from surtr_b200 import FractureContext, synth, FRAGMENT_DTYPE

dev = torch.device("cuda", 0)
N, D, K = 4096, int(sys.argv[1]) if len(sys.argv) > 1 else 4, 400
base = FractureContext(0)
cells = synth.voronoi_cells(base, synth.seeds_uniform(46354, N))
cube = synth.unit_cube()
pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
h = {k: pin(v) for k, v in dict(pv=cube[0], pvo=cube[1], pro=cube[2], pr=cube[3], planes=cells.planes, plane_off=cells.plane_off,
                                cverts=cells.verts, cvo=cells.vert_off).items()}
flush = torch.empty(160 << 20, dtype=torch.uint8, device=dev)
pipes = []
for d in range(D):
    st = torch.cuda.Stream(device=dev)
    cx = FractureContext(0, st.cuda_stream)
    cx.upload_pieces_ptr(h["pv"].data_ptr(), h["pvo"].data_ptr(), h["pro"].data_ptr(), h["pr"].data_ptr(), 1)
    cx.upload_cells_ptr(h["planes"].data_ptr(), h["plane_off"].data_ptr(), h["cverts"].data_ptr(), h["cvo"].data_ptr(), N)
    cx.fracture_event(); c = cx.counts()
    ho = dict(rec=torch.empty(int(c.n_fragments) * FRAGMENT_DTYPE.itemsize, dtype=torch.uint8).pin_memory(),
              verts=torch.empty(int(c.n_verts) * 4, dtype=torch.float32).pin_memory(),
              ring_off=torch.empty(int(c.n_verts) + 1, dtype=torch.int32).pin_memory(),
              ring=torch.empty(int(c.n_ring), dtype=torch.int16).pin_memory())
    pipes.append((cx, st, ho))
T = collections.Counter()
def timed(name, f, *a):
    t0 = time.perf_counter(); f(*a); T[name] += time.perf_counter() - t0
import os
NO_UP, NO_DOWN = bool(os.environ.get('NO_UP')), bool(os.environ.get('NO_DOWN'))
NO_UP_PIECES = bool(os.environ.get('NO_UP_PIECES'))   # drops the four tiny piece copies (220 bytes): per-copy overhead?
ONE_DOWN = bool(os.environ.get('ONE_DOWN'))           # downloads only the vertex array (one copy instead of four)
def run(n, with_flush):
    for i in range(n + D):
        cx, st, ho = pipes[i % D]
        if i >= D and NO_DOWN:
            timed("counts only", cx.counts)
        elif i >= D:
            if ONE_DOWN:
                timed("download_async (verts only)", cx.download_into_async, None, ho["verts"].data_ptr(), None, None)
            else:
                timed("download_async (incl. wait for the event)", cx.download_into_async, ho["rec"].data_ptr(), ho["verts"].data_ptr(), ho["ring_off"].data_ptr(), ho["ring"].data_ptr())
        if i < n:
            if with_flush:
                def fl():
                    with torch.cuda.stream(st):
                        flush.zero_()
                timed("flush launch", fl)
            if not NO_UP and not NO_UP_PIECES: timed("upload_pieces", cx.upload_pieces_ptr, h["pv"].data_ptr(), h["pvo"].data_ptr(), h["pro"].data_ptr(), h["pr"].data_ptr(), 1)
            if not NO_UP: timed("upload_cells", cx.upload_cells_ptr, h["planes"].data_ptr(), h["plane_off"].data_ptr(), h["cverts"].data_ptr(), h["cvo"].data_ptr(), N)
            timed("fracture_event (6 launches)", cx.fracture_event)
    for cx, st, ho in pipes:
        cx.sync()
for with_flush in (False,):
    run(2 * D, with_flush); torch.cuda.synchronize(); T.clear()
    t0 = time.perf_counter(); run(K, with_flush); torch.cuda.synchronize(); tot = time.perf_counter() - t0
    print(f"depth {D} flush {with_flush}: {1e6 * tot / K:.1f} us/step wall;  host time per step:",
          {k: round(1e6 * v / K, 1) for k, v in T.items()}, "sum", round(1e6 * sum(T.values()) / K, 1))
