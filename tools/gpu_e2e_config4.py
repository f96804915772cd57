"""Measurement: ablation + device timeline of the end-to-end loop of bench.py's headline workload (config 4 batches
through the C ABI with pinned host buffers).

  [CUDA_DEVICE_MAX_CONNECTIONS=32] python tools/gpu_e2e_config4.py [events=512] [passes=4]

For batch sizes x contexts in flight it times: compute only (inputs resident), upload + compute, compute + download, and
the full loop; then prints a device timeline (CUDA-event stamps on each context's stream and copy stream) of a few
consecutive batches of the full loop: H2D span, kernel span, D2H span."""
import ctypes as C, json, os, sys, time
import numpy as np
sys.path.insert(0, '.')
import torch
from surtr_b200 import FractureContext, synth, FRAGMENT_DTYPE, load_library

n_events = int(sys.argv[1]) if len(sys.argv) > 1 else 512
n_pass = int(sys.argv[2]) if len(sys.argv) > 2 else 4
dev = torch.device("cuda", 0)
lib = load_library()
lib.surtr_debug_copy_stream.restype = C.c_void_p
lib.surtr_debug_copy_stream.argtypes = [C.c_void_p]
main = torch.cuda.Stream(device=dev)
torch.cuda.set_stream(main)
gen = FractureContext(0, main.cuda_stream)
pieces, cells, ev_p, ev_c = synth.config4_events(gen, range(n_events))
pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory()


def views(e0, e1):
    p = synth.slice_sets(pieces, int(ev_p[e0]), int(ev_p[e1]))
    c = synth.slice_sets(cells, int(ev_c[e0]), int(ev_c[e1]))
    return p, c, (ev_p[e0:e1 + 1] - ev_p[e0]).astype(np.uint32), (ev_c[e0:e1 + 1] - ev_c[e0]).astype(np.uint32)


# per-event output sizes from one resident run
p, c, evp, evc = views(0, n_events)
gen.upload_pieces(p.verts, p.vert_off, p.ring_off, p.ring, evp)
gen.upload_cells(c.planes, c.plane_off, c.verts, c.vert_off, evc)
gen.fracture_event()
rec = gen.download(geometry=False).rec
ev_of = (rec["cell"] // 64).astype(np.int64)
ev_frags = np.bincount(ev_of, minlength=n_events)
ev_verts = np.bincount(ev_of, weights=rec["n_verts"], minlength=n_events).astype(np.int64)
ev_ring = np.bincount(ev_of, weights=rec["n_ring"], minlength=n_events).astype(np.int64)
ts = []
for _ in range(5):
    gen.fracture_event(); gen.counts(); ts.append(gen.last_event_ms()[0])
print(json.dumps({"events": n_events, "fragments": len(rec), "resident_one_batch_ms": float(np.median(ts)),
                  "max_connections": os.environ.get("CUDA_DEVICE_MAX_CONNECTIONS", "default (8)")}), flush=True)
gen.close()


class B:
    pass


def make_batches(bs):
    out = []
    for e0 in range(0, n_events, bs):
        b = B()
        b.e0, b.e1 = e0, min(n_events, e0 + bs)
        p, c, evp, evc = views(b.e0, b.e1)
        b.np, b.nc, b.ne, b.evp, b.evc = p.n, c.n, b.e1 - b.e0, np.ascontiguousarray(evp), np.ascontiguousarray(evc)
        b.h_in = {k: pin(v) for k, v in dict(pv=p.verts[:, :3], pvo=p.vert_off, pro=p.ring_off, pr=p.ring, planes=c.planes,
                                             plane_off=c.plane_off, cverts=c.verts[:, :3], cvo=c.vert_off).items()}
        # (sized for the largest batch: the ablation modes without uploads cut whatever batch the context last received)
        nf = max(int(ev_frags[e:e + bs].sum()) for e in range(0, n_events, bs))
        nv = max(int(ev_verts[e:e + bs].sum()) for e in range(0, n_events, bs))
        nr = max(int(ev_ring[e:e + bs].sum()) for e in range(0, n_events, bs))
        b.h_out = dict(rec=torch.empty(nf * 64, dtype=torch.uint8, pin_memory=True), verts=torch.empty(nv * 3, dtype=torch.float32, pin_memory=True),
                       ring_len=torch.empty(nv, dtype=torch.uint8, pin_memory=True), ring=torch.empty(nr, dtype=torch.int16, pin_memory=True))
        b.sizes, total = FractureContext.fill_input_blob(None, p, c, evp, evc)
        b.blob_in = torch.empty(total, dtype=torch.uint8, pin_memory=True)
        FractureContext.fill_input_blob(b.blob_in.numpy(), p, c, evp, evc)
        al = lambda x: (int(x) + 255) // 256 * 256
        b.cap = al(64 * nf) + al(12 * nv) + al(nv) + al(2 * nr)
        b.blob_out = torch.empty(b.cap, dtype=torch.uint8, pin_memory=True)
        b.bytes_in = sum(t.numel() * t.element_size() for t in b.h_in.values())
        b.bytes_out = sum(t.numel() * t.element_size() for t in b.h_out.values())
        out.append(b)
    return out


BLOB = True


def upload(cx, b):
    if BLOB:
        cx.upload_blob_ptr(b.blob_in.data_ptr(), b.sizes)
        return
    hi = b.h_in
    cx.upload_pieces3_ptr(hi["pv"].data_ptr(), hi["pvo"].data_ptr(), hi["pro"].data_ptr(), hi["pr"].data_ptr(), b.np, b.evp.ctypes.data, b.ne)
    cx.upload_cells3_ptr(hi["planes"].data_ptr(), hi["plane_off"].data_ptr(), hi["cverts"].data_ptr(), hi["cvo"].data_ptr(), b.nc, b.evc.ctypes.data, b.ne)


def download(cx, b):
    if BLOB:
        cx.download_blob_into_async(b.blob_out.data_ptr(), b.cap)
        return
    ho = b.h_out
    cx.download_packed_into_async(ho["rec"].data_ptr(), ho["verts"].data_ptr(), ho["ring_len"].data_ptr(), ho["ring"].data_ptr())


CHAIN = True


def loop(ctxs, batches, k, up=True, down=True, stamps=None):
    nctx = len(ctxs)
    pending = [None] * nctx
    prev_done = None
    for i in range(k * len(batches)):
        b = batches[i % len(batches)]
        s = i % nctx
        cx, st, cst = ctxs[s]
        if pending[s] is not None:
            if down:
                if stamps is not None:
                    cx.counts()                                   # the wait download() would do
                    e = torch.cuda.Event(enable_timing=True); e.record(cst); stamps[pending[s][1]]["d0"] = e
                download(cx, pending[s][0])
                if stamps is not None:
                    e = torch.cuda.Event(enable_timing=True); e.record(cst); stamps[pending[s][1]]["d1"] = e
            else:
                cx.counts()
        if stamps is not None:
            stamps.append({})
            e = torch.cuda.Event(enable_timing=True); e.record(st); stamps[i]["u0"] = e
        if up or i < nctx:
            upload(cx, b)
        if stamps is not None:
            e = torch.cuda.Event(enable_timing=True); e.record(st); stamps[i]["u1"] = e
        if CHAIN and prev_done is not None:
            st.wait_event(prev_done)          # kernels in batch order: batch i's upload overlaps batch i-1's kernels
        cx.fracture_event()
        if CHAIN:
            prev_done = torch.cuda.Event(); prev_done.record(st)
        if stamps is not None:
            e = torch.cuda.Event(enable_timing=True); e.record(st); stamps[i]["k1"] = e
        pending[s] = (b, i)
    for s in range(nctx):
        if pending[s] is not None:
            if down:
                download(ctxs[s][0], pending[s][0])
            else:
                ctxs[s][0].counts()
    for cx, _, _ in ctxs:
        cx.sync()


for blob, chain, bs, nctx in [(bl, True, b, n) for b in (32, 64, 128) for n in (2, 3, 4) for bl in (False, True)]:
    CHAIN, BLOB = chain, blob
    batches = make_batches(bs)
    for nctx in (nctx,):
        if nctx > len(batches):
            continue
        ctxs = []
        for _ in range(nctx):
            st = torch.cuda.Stream(device=dev)
            cx = FractureContext(0, st.cuda_stream)
            ctxs.append((cx, st, torch.cuda.ExternalStream(lib.surtr_debug_copy_stream(cx._h), device=dev)))
        loop(ctxs, batches, 2)
        row = {"blob": blob, "chain": chain, "batch_events": bs, "contexts": nctx}
        for name, up, down in (("compute", False, False), ("up+compute", True, False), ("compute+down", False, True), ("full", True, True)):
            if not up:      # same-size batch resident in every context: compute on whatever was uploaded last
                pass
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            loop(ctxs, batches, n_pass, up, down)
            torch.cuda.synchronize()
            row[name + "_ms_per_batch"] = round(1e3 * (time.perf_counter() - t0) / (n_pass * len(batches)), 4)
        row["h2d_mb"] = round(batches[0].bytes_in / 1e6, 2)
        row["d2h_mb"] = round(batches[0].bytes_out / 1e6, 2)
        print(json.dumps(row), flush=True)
        if bs == 64 and nctx == 3 and blob:
            stamps = []
            g0 = torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            g0.record(main)
            for _, st, cst in ctxs:
                st.wait_event(g0); cst.wait_event(g0)
            loop(ctxs, batches, 2, True, True, stamps)
            torch.cuda.synchronize()
            print("timeline (ms since start): batch ctx | H2D start..end | kernels end | D2H start..end")
            for i in range(len(batches), min(len(stamps), len(batches) + 12)):
                s = stamps[i]
                f = lambda k: f"{g0.elapsed_time(s[k]):8.3f}" if k in s else "    -   "
                print(f"  {i:4d} {i % nctx} | {f('u0')} .. {f('u1')} | {f('k1')} | {f('d0')} .. {f('d1')}")
        for cx, _, _ in ctxs:
            cx.close()
