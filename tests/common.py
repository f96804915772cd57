"""Shared builders for the tests: synthetic inputs of the BASELINE.json shapes and exact comparison helpers.

Inputs are built with the oracle (reference build when present, else the C port) -- the oracle is the checker and
the input generator only; the thing under test is always the CUDA path behind the C ABI.
"""
from __future__ import annotations

import numpy as np

from oracle import portapi, refapi


def have_ref() -> bool:
    return refapi.available()


def scipy_neighbors(seeds: np.ndarray):
    """Delaunay neighbour CSR (ascending) from qhull -- used for big seed sets where the reference's O(n^2)
    DT3D::Triangulate (Inc/DT3D.h:159-267) takes minutes; cross-checked against it for small n."""
    from scipy.spatial import Delaunay
    d = Delaunay(seeds.astype(np.float64))
    indptr, indices = d.vertex_neighbor_vertices
    off = indptr.astype(np.uint32)
    idx = np.concatenate([np.sort(indices[indptr[i]:indptr[i + 1]]) for i in range(len(seeds))]).astype(np.uint32)
    return off, idx


def voronoi(seed: int, n: int):
    """Voronoi cell set of n uniform seeds (Surtr.cpp:1988-1998) in the unit box, built with the reference clipper."""
    s = refapi.seeds_uniform(seed, n)
    if n <= 512:
        off, idx, _ = refapi.dt3d_neighbors(s)
    else:
        off, idx = scipy_neighbors(s)
    return refapi.voronoi_cells(s, off, idx)


def cell_bounds_arrays(cells):
    """Cell vertex streams for the broad phase: the cell polyhedron's own vertices."""
    return cells.verts, cells.vert_off


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
