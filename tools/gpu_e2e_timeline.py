"""Dev/measurement (written at the end of round 1 without GPU time left -- first run pending): a DEVICE timeline of the
pipelined end-to-end loop of bench.py, from CUDA-event stamps on every event's compute stream.

Per event it records, relative to one global start stamp: when the stream reached the upload (t_start), when the uploads
had landed (t_up), when the six kernels + control-block read-back were done (t_done), and on the host when the event was
issued and when its download had been enqueued.  Printed: mean upload time, mean kernel span, mean gap between
consecutive events' t_done (= the loop's rate), how many events' kernel spans overlap at a time, and the share of the
wall time in which NO event was inside its kernel span (the GPU idling behind copies / the host).

  python tools/gpu_e2e_timeline.py [contexts=12] [lag=6] [steps=240]
  CUDA_DEVICE_MAX_CONNECTIONS=32 python tools/gpu_e2e_timeline.py      # the work-queue aliasing suspect
"""
import sys, time, numpy as np, torch
sys.path.insert(0, '.')
from surtr_b200 import FractureContext, synth, FRAGMENT_DTYPE

dev = torch.device("cuda", 0)
D = int(sys.argv[1]) if len(sys.argv) > 1 else 12
LAG = int(sys.argv[2]) if len(sys.argv) > 2 else 6
K = int(sys.argv[3]) if len(sys.argv) > 3 else 240
N = 4096
base = FractureContext(0)
cells = synth.voronoi_cells(base, synth.seeds_uniform(46354, N))
cube = synth.unit_cube()
pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
h = {k: pin(v) for k, v in dict(pv=cube[0][:, :3], pvo=cube[1], pro=cube[2], pr=cube[3], planes=cells.planes,
                                plane_off=cells.plane_off, cverts=cells.verts[:, :3], cvo=cells.vert_off).items()}
pipes = []
for d in range(D):
    st = torch.cuda.Stream(device=dev)
    cx = FractureContext(0, st.cuda_stream)
    cx.set_kdop_directions(3)
    cx.upload_pieces3_ptr(h["pv"].data_ptr(), h["pvo"].data_ptr(), h["pro"].data_ptr(), h["pr"].data_ptr(), 1)
    cx.upload_cells3_ptr(h["planes"].data_ptr(), h["plane_off"].data_ptr(), h["cverts"].data_ptr(), h["cvo"].data_ptr(), N)
    cx.fracture_event(); c = cx.counts()
    ho = dict(rec=torch.empty(int(c.n_fragments) * FRAGMENT_DTYPE.itemsize, dtype=torch.uint8).pin_memory(),
              verts=torch.empty(int(c.n_verts) * 3, dtype=torch.float32).pin_memory(),
              ring_len=torch.empty(int(c.n_verts), dtype=torch.uint8).pin_memory(),
              ring=torch.empty(int(c.n_ring), dtype=torch.int16).pin_memory())
    pipes.append((cx, st, ho))


def run(n, stamps=None, host=None):
    for i in range(n + LAG):
        if i >= LAG:
            cx, st, ho = pipes[(i - LAG) % D]
            cx.download_packed_into_async(ho["rec"].data_ptr(), ho["verts"].data_ptr(), ho["ring_len"].data_ptr(), ho["ring"].data_ptr())
            if host is not None:
                host[i - LAG][1] = time.perf_counter()
        if i < n:
            cx, st, ho = pipes[i % D]
            if host is not None:
                host[i][0] = time.perf_counter()
            if stamps is not None:
                stamps[i][0].record(st)
            cx.upload_pieces3_ptr(h["pv"].data_ptr(), h["pvo"].data_ptr(), h["pro"].data_ptr(), h["pr"].data_ptr(), 1)
            cx.upload_cells3_ptr(h["planes"].data_ptr(), h["plane_off"].data_ptr(), h["cverts"].data_ptr(), h["cvo"].data_ptr(), N)
            if stamps is not None:
                stamps[i][1].record(st)
            cx.fracture_event()
            if stamps is not None:
                stamps[i][2].record(st)
    for cx, st, ho in pipes:
        cx.sync()


run(4 * D)
torch.cuda.synchronize()
stamps = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(K)]
host = [[0.0, 0.0] for _ in range(K)]
t0_ev = torch.cuda.Event(enable_timing=True)
t0_ev.record(torch.cuda.current_stream())
torch.cuda.synchronize()
t0 = time.perf_counter()
run(K, stamps, host)
torch.cuda.synchronize()
wall = time.perf_counter() - t0
T = np.array([[t0_ev.elapsed_time(e) * 1e3 for e in s] for s in stamps])     # us since the start stamp
H = (np.array(host) - t0) * 1e6
lo = 2 * D                                                                    # skip the ramp-up
up, span = T[lo:, 1] - T[lo:, 0], T[lo:, 2] - T[lo:, 1]
done = np.sort(T[lo:, 2])
print(f"{D} contexts, lag {LAG}: {wall * 1e6 / K:.1f} us per event (wall)")
print(f"  upload on the stream     mean {up.mean():7.1f} us   p90 {np.percentile(up, 90):7.1f}")
print(f"  kernels + read-back      mean {span.mean():7.1f} us   p90 {np.percentile(span, 90):7.1f}")
print(f"  gap between completions  mean {np.diff(done).mean():7.1f} us")
print(f"  host: issue -> download enqueued  mean {(H[lo:, 1] - H[lo:, 0]).mean():7.1f} us;  issue -> stream reached it  mean {(T[lo:, 0] - H[lo:, 0]).mean():7.1f} us")
edges = np.concatenate([np.stack([T[lo:, 1], np.ones(K - lo)], 1), np.stack([T[lo:, 2], -np.ones(K - lo)], 1)])
edges = edges[np.argsort(edges[:, 0])]
depth = np.cumsum(edges[:, 1])[:-1]
dt = np.diff(edges[:, 0])
print(f"  events inside their kernel span at a time: mean {np.sum(depth * dt) / dt.sum():.2f};  no event in its kernel span for "
      f"{100 * dt[depth == 0].sum() / dt.sum():.1f} % of the time")
