"""Shared builders for the tests: synthetic inputs of the BASELINE.json shapes and exact comparison helpers.

Inputs are built with the oracle's C port (plus qhull for neighbour lists) -- the oracle is the checker and the
input generator only; the thing under test is always the CUDA path behind the C ABI.
"""
from __future__ import annotations

import hashlib
import os

import numpy as np

from oracle import portapi, refapi
from oracle.refapi import PolySet

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def have_ref() -> bool:
    return refapi.available()


def seeds_uniform(seed: int, n: int) -> np.ndarray:
    """Surtr.cpp:1988-1998: std::mt19937(seed) + std::uniform_real_distribution<double>(-0.5, 0.5), x/y/z order,
    narrowed to float.  libstdc++'s generate_canonical<double,53> takes two 32-bit draws per value
    (lo + hi*2^32) / 2^64; RandomState(seed) is the same init_genrand stream.  Checked against the reference build
    in tests/test_oracle_port.py."""
    rs = np.random.RandomState(seed)
    raw = rs.randint(0, 2 ** 32, size=6 * n, dtype=np.uint64)
    lo = raw[0::2].astype(np.float64)
    hi = raw[1::2].astype(np.float64)
    r = (lo + hi * 4294967296.0) / 18446744073709551616.0
    r = np.minimum(r, np.nextafter(1.0, 0.0))
    return (r * 1.0 + (-0.5)).reshape(n, 3).astype(np.float32)


def scipy_neighbors(seeds: np.ndarray):
    """Delaunay neighbour CSR (ascending) from qhull; a superset of DT3D::Triangulate's (Inc/DT3D.h:159-267)
    neighbour sets that yields the same cells in (V, F) -- checked in tests/test_oracle_port.py."""
    from scipy.spatial import Delaunay
    d = Delaunay(seeds.astype(np.float64))
    indptr, indices = d.vertex_neighbor_vertices
    off = indptr.astype(np.uint32)
    idx = np.concatenate([np.sort(indices[indptr[i]:indptr[i + 1]]) for i in range(len(seeds))]).astype(np.uint32)
    return off, idx


_voro_cache = {}


def voronoi(seed: int, n: int) -> PolySet:
    """Voronoi cell set of n uniform seeds in the unit box (cells double as convex pieces)."""
    key = (seed, n)
    if key not in _voro_cache:
        s = seeds_uniform(seed, n)
        off, idx = scipy_neighbors(s)
        _voro_cache[key] = portapi.voronoi_cells(s, off, idx)
    return _voro_cache[key]


def unit_cube() -> PolySet:
    """Poly::GetBB() (Poly.cpp:587-617)."""
    v = np.array([[-.5, -.5, -.5], [.5, -.5, -.5], [.5, .5, -.5], [-.5, .5, -.5],
                  [-.5, -.5, .5], [.5, -.5, .5], [.5, .5, .5], [-.5, .5, .5]], np.float32)
    nb = np.array([[1, 4, 3], [5, 0, 2], [3, 6, 1], [7, 2, 0], [5, 7, 0], [1, 6, 4], [5, 2, 7], [4, 6, 3]], np.uint16)
    verts = np.zeros((8, 4), np.float32)
    verts[:, :3] = v
    return PolySet(verts, np.array([0, 8], np.uint32), np.arange(0, 25, 3, dtype=np.uint32), nb.reshape(-1).copy())


def concat(sets):
    """Concatenate polysets (with planes) -> (PolySet, piece event offsets)."""
    verts = np.concatenate([s.verts for s in sets])
    ring = np.concatenate([s.ring for s in sets])
    vo, ro, po, pl = [np.zeros(1, np.uint32)], [np.zeros(1, np.uint32)], [np.zeros(1, np.uint32)], []
    ev = [0]
    for s in sets:
        vo.append(s.vert_off[1:] + vo[-1][-1])
        ro.append(s.ring_off[1:] + ro[-1][-1])
        if s.planes is not None:
            po.append(s.poly_face_off[1:] + po[-1][-1])
            pl.append(s.planes)
        ev.append(ev[-1] + s.n)
    out = PolySet(verts, np.concatenate(vo).astype(np.uint32), np.concatenate(ro).astype(np.uint32), ring)
    if pl:
        out.planes = np.concatenate(pl)
        out.poly_face_off = np.concatenate(po).astype(np.uint32)
    return out, np.asarray(ev, np.uint32)


def run_gpu(ctx, pieces, cells, ev_piece_off=None, ev_cell_off=None, bounded=True):
    ctx.upload_pieces(pieces.verts, pieces.vert_off, pieces.ring_off, pieces.ring, ev_piece_off)
    if bounded:
        ctx.upload_cells(cells.planes, cells.plane_off, cells.verts, cells.vert_off, ev_cell_off)
    else:
        ctx.upload_cells(cells.planes, cells.plane_off, None, None, ev_cell_off)
    ctx.fracture_event()
    return ctx.download()


def bits(a):
    a = np.ascontiguousarray(a)
    return a.view(np.uint8) if a.dtype.kind == "f" else a


def assert_fragments_equal(got, want, moments=True, exact_order=True):
    """got: surtr_b200.Fragments, want: oracle PolySet (cell, piece, nfaces, volume, centroid)."""
    assert got.n == want.n, f"fragment count {got.n} != {want.n}"
    assert np.array_equal(got.rec["cell"], want.cell), "piece-to-cell assignment (cell)"
    assert np.array_equal(got.rec["piece"], want.piece), "piece-to-cell assignment (piece)"
    assert np.array_equal(got.rec["n_verts"], want.nverts), "vertex counts"
    assert np.array_equal(got.rec["n_faces"], want.nfaces), "face counts"
    assert np.array_equal(got.vert_off, want.vert_off)
    if exact_order:
        assert np.array_equal(bits(got.verts), bits(want.verts)), "vertex positions (bitwise)"
        assert np.array_equal(got.ring_off, want.ring_off), "ring offsets"
        assert np.array_equal(got.ring, want.ring), "rings"
    if moments:
        assert np.array_equal(bits(got.rec["volume"]), bits(want.volume)), "volumes (bitwise)"
        assert np.array_equal(bits(got.rec["centroid"]), bits(want.centroid)), "centroids (bitwise)"


def digest(*arrays) -> str:
    h = hashlib.sha256()
    for a in arrays:
        h.update(np.ascontiguousarray(a).tobytes())
    return h.hexdigest()


def summary_of(cell, piece, nverts, nfaces, verts, volume) -> dict:
    """Size-independent fingerprint of a fragment set (what tests/golden/summaries.json stores)."""
    return {
        "n": int(len(cell)),
        "n_verts": int(np.sum(nverts)),
        "assign": digest(np.asarray(cell, np.uint32), np.asarray(piece, np.uint32)),
        "vf": digest(np.asarray(nverts, np.uint32), np.asarray(nfaces, np.uint32)),
        "verts": digest(np.asarray(verts, np.float32)),
        "volume": digest(np.asarray(volume, np.float64)),
        "sum_volume": float(np.sum(volume)),
    }


def summary_of_polyset(ps: PolySet) -> dict:
    return summary_of(ps.cell, ps.piece, ps.nverts, ps.nfaces, ps.verts, ps.volume)


def summary_of_fragments(fr) -> dict:
    return summary_of(fr.rec["cell"], fr.rec["piece"], fr.rec["n_verts"], fr.rec["n_faces"], fr.verts, fr.rec["volume"])


def recursion_levels(depth=3, seeds_per_level=64, base_seed=1000):
    """Config 5 cell sets: level l uses mt19937(base_seed + l)."""
    return [voronoi(base_seed + l, seeds_per_level) for l in range(depth)]


def fragments_as_polyset(fr) -> PolySet:
    return PolySet(fr.verts, fr.vert_off, fr.ring_off, fr.ring)


def degenerate_large_inputs():
    """In-plane cuts for the large tiers: the bunny ACH (107 vertices: shared-memory large tier) and the bunny mesh
    polyhedron (2503 vertices: global-memory tier) against 48 sequences of 1-4 axis-aligned planes, each through the
    exact coordinate of a random vertex of one of the two pieces (signed distance exactly 0 there)."""
    a = np.load(os.path.join(GOLDEN, "config1_bunny32.npz"))
    m = np.load(os.path.join(GOLDEN, "bunny_mesh_x32.npz"))
    ach = PolySet(a["ach_verts"], a["ach_vert_off"], a["ach_ring_off"], a["ach_ring"])
    mesh = PolySet(m["mesh_verts"], m["mesh_vert_off"], m["mesh_ring_off"], m["mesh_ring"])
    pieces, _ = concat([ach, mesh])
    rng = np.random.RandomState(5)
    planes, off = [], [0]
    for _ in range(48):
        for _ in range(rng.randint(1, 5)):
            src = ach if rng.rand() < 0.5 else mesh
            v = src.verts[rng.randint(len(src.verts)), :3]
            ax = rng.randint(3)
            sgn = np.float32(1.0 if rng.rand() < 0.5 else -1.0)
            n = np.zeros(3, np.float32)
            n[ax] = sgn
            planes.append([n[0], n[1], n[2], -sgn * v[ax]])
        off.append(len(planes))
    return pieces, np.asarray(planes, np.float32), np.asarray(off, np.uint32)
