"""Dev: per-candidate cycle histogram of K3 (surtr_debug.h)."""
import ctypes as C, sys, numpy as np
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')   # run from the repo root
from surtr_b200 import FractureContext, engine
import common
lib = engine.load_library()
lib.surtr_debug_enable.argtypes = [C.c_void_p, C.c_int]
lib.surtr_debug_read.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64]
ctx = FractureContext(0)
lib.surtr_debug_enable(ctx._h, 1)
for name, pieces, cells in [("cube x 4096", common.unit_cube(), common.voronoi(46354, 4096)),
                            ("10000 x 256", common.voronoi(1234, 10000), common.voronoi(46354, 256))]:
    for rep in range(2):
        got = common.run_gpu(ctx, pieces, cells)
    c = ctx.counts()
    n = int(c.n_candidates)
    d = np.zeros((n, 8), np.uint32)
    lib.surtr_debug_read(ctx._h, d.ctypes.data_as(C.c_void_p), n)
    tot = d[:, :4].sum(1)
    print(name, "cands", n, "event ms", ctx.last_event_ms())
    print("  cycles per pair: total p50 %d p90 %d p99 %d max %d" % tuple(np.percentile(tot, [50, 90, 99, 100])))
    for k, nm in enumerate(["load", "clip", "moments", "write"]):
        x = d[:, k]
        print("   %-8s mean %8.0f p50 %8.0f p99 %8.0f max %8d" % (nm, x.mean(), np.percentile(x, 50), np.percentile(x, 99), x.max()))
    cuts = d[:, 5]; seq = d[:, 4]
    m = cuts > 0
    print("   cuts/pair mean %.2f; cycles per cut (pairs w/o seq): %.0f ; pairs with seq cuts: %d" % (cuts.mean(), (d[m & (seq == 0), 1] / np.maximum(1, cuts[m & (seq == 0)])).mean(), (seq > 0).sum()))
    s1 = seq > 0
    if s1.any():
        print("   seq pairs: clip cycles mean %.0f max %d ; non-seq clip mean %.0f max %d" % (d[s1, 1].mean(), d[s1, 1].max(), d[~s1, 1].mean(), d[~s1, 1].max()))
    nocut = (cuts == 0)
    if nocut.any():
        print("   pairs with 0 cuts: %d, clip cycles mean %.0f (planes mean %.1f)" % (nocut.sum(), d[nocut, 1].mean(), d[nocut, 7].mean()))
