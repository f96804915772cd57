"""The DEVICE code of the hot path checked WITHOUT a GPU: surtr_b200/csrc/clip_sub.cuh (K3's small tier,
sub_clip_by_planes<32>, and K4's sub_fragment_moments<16>) is compiled for the host over a SIMT shim (tests/emu/: 32
lanes as coroutines that meet at every warp collective) and run pair by pair against the committed outputs of the
reference build and against the oracle port -- bitwise, like the GPU parity tests.  This is test infrastructure: it
executes the kernel SOURCE on the CPU to catch a regression before GPU time is spent; the product never uses it
(surtr_b200/ has no CPU path) and the -m gpu tests remain the parity proof for the compiled kernels."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import common
from common import GOLDEN, bits
from oracle import portapi as P
from test_oracle_port import load_polyset

HERE = os.path.dirname(os.path.abspath(__file__))
EMU = os.path.join(HERE, "emu")
LIB = os.path.join(EMU, "_build", "libk3emu.so")
SRC = [os.path.join(EMU, "k3_emu.cpp"), os.path.join(EMU, "simt_shim.h")] + \
      [os.path.join(HERE, "..", "surtr_b200", "csrc", f) for f in ("clip_sub.cuh", "clip_fast.cuh", "clip_duo.cuh", "clip_global.cuh", "clip_warp.cuh", "surtr_math.cuh")]
CUDA_INC = os.environ.get("CUDA_HOME", "/usr/local/cuda") + "/include"


@pytest.fixture(scope="module")
def emu():
    if not os.path.exists(os.path.join(CUDA_INC, "cuda_runtime.h")):
        pytest.skip("CUDA headers not found (vector types for the host compile)")
    if not os.path.exists(LIB) or any(os.path.getmtime(s) > os.path.getmtime(LIB) for s in SRC):
        os.makedirs(os.path.dirname(LIB), exist_ok=True)
        r = subprocess.run(["g++", "-std=c++17", "-O1", "-ffp-contract=off", "-fPIC", "-shared", "-Wall", "-Wno-unknown-pragmas",
                            "-I", CUDA_INC, SRC[0], "-o", LIB], capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
    lib = C.CDLL(LIB)
    lib.k3emu_pair.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int] + [C.c_void_p] * 7
    lib.k3emu_pair_large.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int] + [C.c_void_p] * 7
    return lib


@pytest.fixture(params=[2, 22, 16, 4, 0], ids=["fast64", "fast64lat", "duo", "fast128", "round1"])
def small(emu, request):
    """The small-tier clipper k3emu_pair runs: clip_fast.cuh with 64 slots in its throughput and its latency build (clip_fast_kernel<2,false>, the main K3
    launch), with 128 slots (clip_fast_kernel<4,true>), and round 1's clip_sub.cuh (kept for A/B profiles)."""
    emu.k3emu_set_variant(request.param)
    yield emu
    emu.k3emu_set_variant(2)


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


class Out:
    def __init__(self, cap=64, deg=8):
        self.verts = np.zeros((cap, 4), np.float32)
        self.ring_off = np.zeros(cap + 1, np.uint32)
        self.ring = np.zeros(cap * deg, np.uint16)
        self.info = np.zeros(8, np.int32)
        self.volume = np.zeros(1, np.float64)
        self.centroid = np.zeros(3, np.float32)
        self.inertia = np.zeros(6, np.float32)


def run_pair(lib, ps, p, planes, tier=None):
    """Piece p of the PolySet through the emulated kernels.  tier = None: the small tier (one warp, clip_sub.cuh);
    tier = (warps, vertex slots): the large / unbounded tier of clip_global.cuh as a block of that many warps.
    Returns (status, Out)."""
    v0, v1 = int(ps.vert_off[p]), int(ps.vert_off[p + 1])
    r0, r1 = int(ps.ring_off[v0]), int(ps.ring_off[v1])
    verts = np.ascontiguousarray(ps.verts[v0:v1], np.float32)
    roff = np.ascontiguousarray(ps.ring_off[v0:v1 + 1] - r0, np.uint32)
    ring = np.ascontiguousarray(ps.ring[r0:r1], np.uint16)
    planes = np.ascontiguousarray(planes, np.float32).reshape(-1, 4)
    if tier is None:
        o = Out()
        rc = lib.k3emu_pair(_p(verts), _p(roff), _p(ring), v1 - v0, _p(planes), len(planes), _p(o.verts), _p(o.ring_off), _p(o.ring),
                            _p(o.info), _p(o.volume), _p(o.centroid), _p(o.inertia))
    else:
        o = Out(tier[1], 16)
        rc = lib.k3emu_pair_large(tier[0], tier[1], _p(verts), _p(roff), _p(ring), v1 - v0, _p(planes), len(planes), _p(o.verts),
                                  _p(o.ring_off), _p(o.ring), _p(o.info), _p(o.volume), _p(o.centroid), _p(o.inertia))
    assert rc == 0, f"emulation inconsistency {rc}"
    return int(o.info[0]), o


def check_against(want, i, o):
    """Emulated fragment == fragment i of the expected PolySet, bit for bit."""
    v0, v1 = int(want.vert_off[i]), int(want.vert_off[i + 1])
    nv, ne = int(o.info[1]), int(o.info[2])
    assert nv == v1 - v0, "vertex count"
    assert np.array_equal(bits(o.verts[:nv, :3]), bits(want.verts[v0:v1, :3])), "vertex positions (bitwise)"
    r0 = int(want.ring_off[v0])
    assert np.array_equal(o.ring_off[:nv + 1], want.ring_off[v0:v1 + 1] - r0), "ring offsets"
    assert np.array_equal(o.ring[:ne], want.ring[r0:r0 + ne]), "rings"
    assert int(o.info[6]) == int(want.nfaces[i]), "face count"
    if want.volume is not None:
        assert bits(o.volume)[0] == bits(want.volume[i:i + 1])[0], "volume (bitwise)"
        assert np.array_equal(bits(o.centroid), bits(want.centroid[i])), "centroid (bitwise)"


def run_event(lib, pieces, planes, plane_off, want, stats, tier=None, cells=None):
    index = {(int(c), int(p)): i for i, (c, p) in enumerate(zip(want.cell, want.piece))}
    seen = 0
    for c in (range(len(plane_off) - 1) if cells is None else cells):
        pl = planes[int(plane_off[c]):int(plane_off[c + 1])]
        for p in range(pieces.n):
            status, o = run_pair(lib, pieces, p, pl, tier)
            stats["pairs"] += 1
            stats["seq_cuts"] += int(o.info[3])
            stats["cuts"] += int(o.info[4])
            if status != 0:
                stats["overflow"] += 1      # the pair belongs to the large tier: not this code
                seen += (c, p) in index
                continue
            if (c, p) in index:
                assert o.info[1] > 0, f"pair ({c}, {p}) lost its fragment"
                check_against(want, index[(c, p)], o)
                seen += 1
            else:
                assert o.info[1] == 0, f"pair ({c}, {p}) produced a fragment the reference does not have"
    assert seen == (want.n if cells is None else sum(1 for c in want.cell if int(c) in set(cells)))


@pytest.mark.parametrize("name", ["cube_x64", "pieces200_x32"])
def test_emulated_kernels_on_reference_fixtures(small, name):
    """Committed outputs of the REFERENCE build (tests/golden/make_golden.py): every (piece, cell) pair of the event."""
    d = np.load(os.path.join(GOLDEN, name + ".npz"))
    cells, pieces, want = load_polyset(d, "cells_"), load_polyset(d, "pieces_"), load_polyset(d, "frag_")
    stats = dict(pairs=0, seq_cuts=0, cuts=0, overflow=0)
    run_event(small, pieces, cells.planes, cells.plane_off, want, stats)
    assert stats["pairs"] == pieces.n * cells.n and stats["overflow"] == 0 and stats["cuts"] > 2 * want.n   # (pairs the plane prefilter kills never cut)


def test_emulated_kernels_on_degenerate_cuts(small):
    """Planes through vertices, along edges and coincident with faces (the in-plane band, the sequential replay, the
    degree-2 splice, the all-in-plane box test): expected = the reference build's output."""
    d = np.load(os.path.join(GOLDEN, "degenerate_x400.npz"))
    pieces, want = load_polyset(d, "pieces_"), load_polyset(d, "frag_")
    stats = dict(pairs=0, seq_cuts=0, cuts=0, overflow=0)
    run_event(small, pieces, d["planes"], d["plane_off"], want, stats)
    assert stats["seq_cuts"] > 100 and stats["overflow"] == 0


def test_emulated_kernels_against_the_oracle_port(small):
    """Config-4-shaped pairs (Voronoi pieces x Voronoi cells) and config 2's first cells against the oracle port, plus
    the lazy compaction: a 40-plane cell drives the slot counter past 64 and forces the in-kernel renumbering."""
    pieces, cells = common.voronoi(1234, 150), common.voronoi(46354, 24)
    want = P.apply_fracture(pieces, cells.planes, cells.plane_off)
    stats = dict(pairs=0, seq_cuts=0, cuts=0, overflow=0)
    run_event(small, pieces, cells.planes, cells.plane_off, want, stats)
    assert stats["overflow"] == 0

    cube = common.unit_cube()
    big = common.voronoi(46354, 4096)
    sel = np.argsort(-np.diff(big.plane_off.astype(np.int64)))[:48]      # the cells with the most planes
    planes = np.concatenate([big.planes[big.plane_off[c]:big.plane_off[c + 1]] for c in sel])
    off = np.concatenate([[0], np.cumsum([big.plane_off[c + 1] - big.plane_off[c] for c in sel])]).astype(np.uint32)
    want = P.apply_fracture(cube, planes, off)
    stats = dict(pairs=0, seq_cuts=0, cuts=0, overflow=0)
    run_event(small, cube, planes, off, want, stats)
    assert want.n == 48 and stats["cuts"] > 48 * 20


def test_emulated_plane_queue_beyond_one_and_two_words(small):
    """Cells of 40, 64, 65 and 150 planes (tangent planes of a sphere, most of them far from the piece): the plane
    prefilter keeps one bit per plane in two 32-bit words and visits planes beyond the 64th one after the other
    (clip_fast.cuh, PlaneQueue); expected = the oracle port."""
    pieces = common.voronoi(1234, 40)
    rng = np.random.default_rng(5)
    planes, off = [], [0]
    for n in (40, 64, 65, 150):
        d = rng.normal(size=(n, 3))
        d /= np.linalg.norm(d, axis=1, keepdims=True)
        r = rng.uniform(0.25, 0.9, size=n)           # the sphere sits at the box centre; small r cuts, large r only touches the prefilter
        c = np.array([0.5, 0.5, 0.5])
        pl = np.concatenate([d, -(d @ c + r)[:, None]], axis=1).astype(np.float32)
        planes.append(pl)
        off.append(off[-1] + n)
    planes, off = np.concatenate(planes), np.array(off, np.uint32)
    want = P.apply_fracture(pieces, planes, off)
    stats = dict(pairs=0, seq_cuts=0, cuts=0, overflow=0)
    run_event(small, pieces, planes, off, want, stats)
    assert want.n >= 10 and stats["cuts"] > 50


@pytest.mark.parametrize("warps", [1, 4, 8])
def test_emulated_large_tiers_on_reference_fixtures(emu, warps):
    """clip_global.cuh (global_clip_by_planes + global_fragment_moments) as one warp, as the four-warp block of
    clip_shared_kernel and as the eight-warp block of clip_global_kernel: the cube event and the degenerate sequences
    (this tier has its own sequential patch / splice code) against the reference build's outputs."""
    d = np.load(os.path.join(GOLDEN, "cube_x64.npz"))
    cells, pieces, want = load_polyset(d, "cells_"), load_polyset(d, "pieces_"), load_polyset(d, "frag_")
    stats = dict(pairs=0, seq_cuts=0, cuts=0, overflow=0)
    run_event(emu, pieces, cells.planes, cells.plane_off, want, stats, tier=(warps, 256), cells=range(0, 64, 4))
    d = np.load(os.path.join(GOLDEN, "degenerate_x400.npz"))
    pieces, want = load_polyset(d, "pieces_"), load_polyset(d, "frag_")
    run_event(emu, pieces, d["planes"], d["plane_off"], want, stats, tier=(warps, 256), cells=range(0, 400, 5))
    assert stats["overflow"] == 0 and stats["seq_cuts"] > 20


def test_emulated_global_tier_on_the_bunny_mesh(emu):
    """Row f-1: the 2503-vertex non-convex mesh polyhedron x cells of the 32-cell pattern, eight warps per pair and a
    workspace as the engine sizes it -- expected fragments from the REFERENCE build (bunny_mesh_x32.npz); and the
    107-vertex ACH in the four-warp large tier against the oracle port."""
    d = np.load(os.path.join(GOLDEN, "bunny_mesh_x32.npz"))
    mesh, want = load_polyset(d, "mesh_"), load_polyset(d, "frag_")
    stats = dict(pairs=0, seq_cuts=0, cuts=0, overflow=0)
    run_event(emu, mesh, d["planes"], d["plane_off"], want, stats, tier=(8, 4096), cells=[0, 5, 11, 17, 23, 31])
    assert stats["overflow"] == 0
    a = np.load(os.path.join(GOLDEN, "config1_bunny32.npz"))
    ach = common.PolySet(a["ach_verts"], a["ach_vert_off"], a["ach_ring_off"], a["ach_ring"])
    want = P.apply_fracture(ach, d["planes"], d["plane_off"])
    run_event(emu, ach, d["planes"], d["plane_off"], want, stats, tier=(4, 256))
    assert stats["overflow"] == 0 and want.n > 20


def test_emulated_global_tier_with_wide_rings_and_lazy_compaction(emu):
    """The global tier's ring stride is a run-time value (a vertex may have any number of neighbours) and its slots are
    renumbered lazily: (a) the degenerate sequences again with 24 ring slots per vertex (not a power of two) and a
    workspace so small (40 slots for 8-vertex pieces cut up to four times) that the renumbering runs inside the plane
    loop; (b) a cone whose apex has 40 neighbours, cut by Voronoi cells, with 48 ring slots -- against the oracle port."""
    emu.k3emu_set_gd.argtypes = [C.c_int]
    d = np.load(os.path.join(GOLDEN, "degenerate_x400.npz"))
    pieces, want = load_polyset(d, "pieces_"), load_polyset(d, "frag_")
    stats = dict(pairs=0, seq_cuts=0, cuts=0, overflow=0)
    try:
        emu.k3emu_set_gd(24)
        run_event(emu, pieces, d["planes"], d["plane_off"], want, stats, tier=(4, 40), cells=range(0, 400, 7))
        d2 = np.load(os.path.join(GOLDEN, "cube_x64.npz"))
        cells, cube, want2 = load_polyset(d2, "cells_"), load_polyset(d2, "pieces_"), load_polyset(d2, "frag_")
        run_event(emu, cube, cells.planes, cells.plane_off, want2, stats, tier=(1, 64), cells=range(0, 64, 3))   # 20+ cuts in 64 slots
        assert stats["overflow"] == 0 and stats["seq_cuts"] > 10
        import test_gpu_parity as T
        import hostapi as H
        verts, tri = T._cone_mesh(40)
        mesh = H.mesh_polyhedron(verts, tri)
        vc = common.voronoi(46354, 24)
        cp = vc.subset(range(vc.n))
        cp.verts = vc.verts.copy()
        cp.verts[:, :3] = vc.verts[:, :3] * np.float32(2.2) + np.array([0.0, 0.6, 0.0], np.float32)
        planes, off = P.face_planes(cp)
        wantc = P.apply_fracture(mesh, planes, off, cap_frags=256, cap_verts=20000)
        emu.k3emu_set_gd(48)
        run_event(emu, mesh, planes, off, wantc, stats, tier=(8, 256), cells=range(0, 24, 2))
        assert stats["overflow"] == 0 and wantc.n >= 8
    finally:
        emu.k3emu_set_gd(16)


def test_emulated_moments_of_two_different_fragments_per_warp(emu):
    """assemble_gather_kernel gives every fragment 16 lanes: two NEIGHBOURING fragments of different size share a warp
    and run sub_fragment_moments<16> in lock step.  Every consecutive pair of the reference's pieces200_x32 fragments
    (and a lone last fragment next to an idle half-warp): face count, volume and centroid bitwise; inertia against the
    oracle's double-precision integral."""
    d = np.load(os.path.join(GOLDEN, "pieces200_x32.npz"))
    cells, pieces = load_polyset(d, "cells_"), load_polyset(d, "pieces_")
    want = P.apply_fracture(pieces, cells.planes, cells.plane_off, inertia=True)
    ref = load_polyset(d, "frag_")
    assert np.array_equal(bits(want.volume), bits(ref.volume)) and np.array_equal(want.nfaces, ref.nfaces)
    emu.k3emu_moments_two.argtypes = [C.c_void_p] * 8
    pairs = [(i, i + 1) for i in range(0, want.n - 1, 2)] + [(want.n - 1, None), (None, 0)]
    worst = 0.0
    for a, b in pairs:
        vs, ros, rs, ns = [], [], [], []
        for i in (a, b):
            if i is None:
                vs.append(np.zeros((1, 4), np.float32)); ros.append(np.zeros(2, np.uint32)); rs.append(np.zeros(1, np.uint16)); ns.append(0)
                continue
            v0, v1 = int(want.vert_off[i]), int(want.vert_off[i + 1])
            r0, r1 = int(want.ring_off[v0]), int(want.ring_off[v1])
            vs.append(np.ascontiguousarray(want.verts[v0:v1], np.float32))
            ros.append(np.ascontiguousarray(want.ring_off[v0:v1 + 1] - r0, np.uint32))
            rs.append(np.ascontiguousarray(want.ring[r0:r1], np.uint16))
            ns.append(v1 - v0)
        ptr = lambda arrs: (C.c_void_p * 2)(*[x.ctypes.data for x in arrs])
        n = np.asarray(ns, np.int32)
        faces, vol, cen, ine = np.zeros(2, np.int32), np.zeros(2, np.float64), np.zeros((2, 3), np.float32), np.zeros((2, 6), np.float32)
        assert emu.k3emu_moments_two(ptr(vs), ptr(ros), ptr(rs), _p(n), _p(faces), _p(vol), _p(cen), _p(ine)) == 0
        for h, i in enumerate((a, b)):
            if i is None:
                continue
            assert faces[h] == want.nfaces[i] and bits(vol[h:h + 1])[0] == bits(want.volume[i:i + 1])[0], (a, b)
            assert np.array_equal(bits(cen[h]), bits(want.centroid[i])), (a, b)
            if want.volume[i] > 1e-9:
                scale = max(np.abs(want.inertia[i, :3]).max(), 1e-300)
                worst = max(worst, float(np.abs(ine[h].astype(np.float64) - want.inertia[i]).max() / scale))
    assert worst < 1e-4


@pytest.mark.parametrize("schedule", [1, 7, 2026])
def test_emulated_kernels_do_not_depend_on_the_lane_order(emu, schedule):
    """Between two collectives the emulated lanes run one after the other; on the GPU they run in any order.  Code whose
    conflicting shared-memory accesses are all separated by __syncwarp / __syncthreads gives the same result for every
    order: descending lanes and two random permutations per pass must reproduce the reference fixtures as well (small
    tier incl. the concurrent ring-byte updates of the vertex insertion, the four-warp and the eight-warp tier)."""
    emu.k3emu_set_schedule(schedule)
    try:
        stats = dict(pairs=0, seq_cuts=0, cuts=0, overflow=0)
        d = np.load(os.path.join(GOLDEN, "cube_x64.npz"))
        cells, pieces, want = load_polyset(d, "cells_"), load_polyset(d, "pieces_"), load_polyset(d, "frag_")
        run_event(emu, pieces, cells.planes, cells.plane_off, want, stats)
        run_event(emu, pieces, cells.planes, cells.plane_off, want, stats, tier=(4, 256), cells=range(0, 64, 8))
        d = np.load(os.path.join(GOLDEN, "degenerate_x400.npz"))
        pieces, want = load_polyset(d, "pieces_"), load_polyset(d, "frag_")
        run_event(emu, pieces, d["planes"], d["plane_off"], want, stats, cells=range(0, 400, 3))
        run_event(emu, pieces, d["planes"], d["plane_off"], want, stats, tier=(8, 256), cells=range(1, 400, 16))
        d = np.load(os.path.join(GOLDEN, "pieces200_x32.npz"))
        cells, pieces, want = load_polyset(d, "cells_"), load_polyset(d, "pieces_"), load_polyset(d, "frag_")
        run_event(emu, pieces, cells.planes, cells.plane_off, want, stats, cells=[schedule % 32])
        assert stats["overflow"] == 0 and stats["seq_cuts"] > 30
    finally:
        emu.k3emu_set_schedule(0)


def test_shim_collectives():
    """The shim's own semantics on a known case: widths, segment boundaries and byte intrinsics are what CUDA documents."""
    src = os.path.join(EMU, "_build", "shim_selftest.cpp")
    os.makedirs(os.path.dirname(src), exist_ok=True)
    open(src, "w").write(r'''
#include "../simt_shim.h"
int main()
{
    unsigned ballots[32]; int up[32], xr[32], idx[32], mx[32];
    simt::run_warp([&](int lane) {
        ballots[lane] = __ballot_sync(0xffffffffu, lane % 3 == 0);
        up[lane] = __shfl_up_sync(0xffffffffu, lane, 2, 16);
        xr[lane] = __shfl_xor_sync(0xffffffffu, lane, 4, 16);
        idx[lane] = __shfl_sync(0xffffffffu, lane * 10, 5, 16);
        if (lane & 1) __syncwarp(); else __syncwarp();
        mx[lane] = __reduce_max_sync(0xffffffffu, lane ^ 21);
    });
    /* a block of four warps: per-warp collectives interleaved with block barriers */
    int wsum[128], any[128], shared[4] = { 0, 0, 0, 0 };
    simt::run_block(128, [&](int tid) {
        const int w = tid >> 5, l = tid & 31;
        int v = l;
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (l == 0) shared[w] = v + w;
        __syncthreads();
        wsum[tid] = shared[0] + shared[1] + shared[2] + shared[3];
        any[tid] = __syncthreads_or(tid == 77);
        if (w == 2) __syncwarp();          /* only one warp takes this collective: legal */
        __syncthreads();
    });
    for (int t = 0; t < 128; t++)
        if (wsum[t] != 4 * 496 + 6 || any[t] != 1) return 10;
    for (int l = 0; l < 32; l++)
    {
        if (ballots[l] != 0x49249249u) return 1;
        if (up[l] != ((l & 15) >= 2 ? l - 2 : l)) return 2;
        if (xr[l] != (l ^ 4)) return 3;
        if (idx[l] != ((l & 16) | 5) * 10) return 4;
        if (mx[l] != 31) return 5;
    }
    if (__byte_perm(0x33221100u, 0x77665544u, 0x6u) != 0x00000066u) return 6;      /* byte 6, then byte 0 three times */
    if (__byte_perm(0x33221100u, 0x77665544u, 0x3210u) != 0x33221100u) return 7;
    if (__vcmpeq4(0x11223344u, 0x11AA33BBu) != 0xff00ff00u) return 8;
    if (__ffs(0) != 0 || __ffs(8) != 4 || __clzll(0) != 64 || __clzll(1) != 63 || __popcll(~0ull) != 64) return 9;
    return 0;
}
''')
    exe = os.path.join(EMU, "_build", "shim_selftest")
    if not os.path.exists(os.path.join(CUDA_INC, "cuda_runtime.h")):
        pytest.skip("CUDA headers not found")
    r = subprocess.run(["g++", "-std=c++17", "-O1", "-I", CUDA_INC, src, "-o", exe], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert subprocess.run([exe]).returncode == 0
