"""Synthetic workloads of the BASELINE.json shapes, built with the PRODUCT path only (no oracle).

Voronoi cell sets are derived from the seeds' Delaunay neighbours: cell i = the container box Poly::GetBB()
(Poly.cpp:587-617) cut by the bisector half-spaces Plane((Si+Sj)*0.5, Sj-Si) towards its neighbours j (ascending)
-- the cutting itself runs on the GPU through the C ABI (one "cell" per plane list, one unit-cube piece).  Cell
faces then get their planes the way a VMACH::PolygonFace does (VMACH.cpp:289-310): Plane(v0, v1, v2) of the first
three distinct loop vertices.  This replaces the voro++ call sites Surtr.cpp:2007-2067 (dependency not vendored).
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

f32 = np.float32


@dataclass
class CellSet:
    """A Voronoi pattern: polyhedra (usable as convex pieces) + outward face planes (usable as cells)."""
    verts: np.ndarray       # float32 [NV,4]
    vert_off: np.ndarray    # uint32 [n+1]
    ring_off: np.ndarray    # uint32 [NV+1]
    ring: np.ndarray        # uint16 [NE]
    planes: np.ndarray      # float32 [NF,4]
    plane_off: np.ndarray   # uint32 [n+1]

    @property
    def n(self) -> int:
        return len(self.vert_off) - 1


def seeds_uniform(seed: int, n: int) -> np.ndarray:
    """Surtr::GenerateVoronoi seeds (Surtr.cpp:1988-1998): std::mt19937(seed), uniform_real_distribution<double>
    (-0.5, 0.5) in x, y, z order, narrowed to float.  libstdc++ draws two 32-bit words per double."""
    rs = np.random.RandomState(seed)
    raw = rs.randint(0, 2 ** 32, size=6 * n, dtype=np.uint64)
    r = (raw[0::2].astype(np.float64) + raw[1::2].astype(np.float64) * 4294967296.0) / 18446744073709551616.0
    r = np.minimum(r, np.nextafter(1.0, 0.0))
    return (r * 1.0 + (-0.5)).reshape(n, 3).astype(f32)


def unit_cube():
    """Poly::GetBB() in the flat layout."""
    v = np.array([[-.5, -.5, -.5], [.5, -.5, -.5], [.5, .5, -.5], [-.5, .5, -.5],
                  [-.5, -.5, .5], [.5, -.5, .5], [.5, .5, .5], [-.5, .5, .5]], f32)
    nb = np.array([[1, 4, 3], [5, 0, 2], [3, 6, 1], [7, 2, 0], [5, 7, 0], [1, 6, 4], [5, 2, 7], [4, 6, 3]], np.uint16)
    verts = np.zeros((8, 4), f32)
    verts[:, :3] = v
    return verts, np.array([0, 8], np.uint32), np.arange(0, 25, 3, dtype=np.uint32), nb.reshape(-1).copy()


def delaunay_neighbors(seeds: np.ndarray):
    """Neighbour CSR (ascending j) of the seeds' 3-D Delaunay triangulation."""
    from scipy.spatial import Delaunay
    d = Delaunay(seeds.astype(np.float64))
    indptr, indices = d.vertex_neighbor_vertices
    off = indptr.astype(np.uint32)
    idx = np.concatenate([np.sort(indices[indptr[i]:indptr[i + 1]]) for i in range(len(seeds))]).astype(np.uint32)
    return off, idx


def bisector_planes(seeds: np.ndarray, nb_off: np.ndarray, nb_idx: np.ndarray) -> np.ndarray:
    """Plane((Si+Sj)*0.5, Sj-Si) = (n, -Dot(mid, n)) with SimpleMath's op order ((x*x' + y*y') + z*z')."""
    i = np.repeat(np.arange(len(seeds)), np.diff(nb_off).astype(np.int64))
    si, sj = seeds[i].astype(f32), seeds[nb_idx].astype(f32)
    mid = (si + sj) * f32(0.5)
    n = sj - si
    w = -((mid[:, 0] * n[:, 0] + mid[:, 1] * n[:, 1]) + mid[:, 2] * n[:, 2])
    return np.concatenate([n, w[:, None]], axis=1).astype(f32)


def _plane_from_points(p1, p2, p3):
    """SimpleMath Plane(p1,p2,p3) (SimpleMath.inl:2773-2780) in float32, one rounding per operation."""
    a, b = (p1 - p2).astype(f32), (p1 - p3).astype(f32)
    n = np.array([f32(a[1] * b[2]) - f32(a[2] * b[1]), f32(a[2] * b[0]) - f32(a[0] * b[2]),
                  f32(a[0] * b[1]) - f32(a[1] * b[0])], f32)
    lsq = f32(f32(f32(n[0] * n[0]) + f32(n[1] * n[1])) + f32(n[2] * n[2]))
    if lsq == 0:
        n = np.zeros(3, f32)
    else:
        n = (n / np.sqrt(lsq, dtype=f32)).astype(f32)
    w = -f32(f32(f32(n[0] * p1[0]) + f32(n[1] * p1[1])) + f32(n[2] * p1[2]))
    return np.array([n[0], n[1], n[2], w], f32)


def _face_vertex_lists(verts, vert_off, ring_off, ring):
    """Per polyhedron, the vertex list of every face as VMACH::PolygonFace would hold it: loops in Poly::ExtractFaces
    order (Poly.cpp:89-126) fed through PolygonFace::AddVertex, which drops a vertex closer than 1e-12 to one already
    in the face (VMACH.cpp:289-301)."""
    ring = ring.astype(np.int64)
    for i in range(len(vert_off) - 1):
        v0, v1 = int(vert_off[i]), int(vert_off[i + 1])
        rings = [ring[ring_off[v]:ring_off[v + 1]].tolist() for v in range(v0, v1)]
        pos = verts[v0:v1, :3]
        visited = set()
        faces = []
        for a in range(v1 - v0):
            for b in rings[a]:
                if (a, b) in visited:
                    continue
                loop, prev, cur = [a], a, b
                while cur != a:
                    visited.add((prev, cur))
                    loop.append(cur)
                    r = rings[cur]
                    k = r.index(prev)
                    prev, cur = cur, r[k - 1]
                visited.add((prev, cur))
                kept = []
                for v in loop:
                    p = pos[v]
                    if any(float(np.sqrt(np.sum((p - q) ** 2, dtype=f32))) < 1e-12 for q in kept):
                        continue
                    kept.append(p)
                faces.append(kept)
        yield faces


def face_planes(verts, vert_off, ring_off, ring):
    """Faces in Poly::ExtractFaces order (Poly.cpp:89-126), plane per face by the PolygonFace::AddVertex route."""
    planes, plane_off = [], [0]
    for faces in _face_vertex_lists(verts, vert_off, ring_off, ring):
        for kept in faces:
            planes.append(_plane_from_points(*kept[:3]) if len(kept) >= 3 else np.array([0, 1, 0, 0], f32))
        plane_off.append(len(planes))
    return np.asarray(planes, f32).reshape(-1, 4), np.asarray(plane_off, np.uint32)


def pattern_arrays(verts, vert_off, ring_off, ring):
    """A cell set as the flat pattern of surtr_upload_pattern: (face_verts4, face_vert_off, cell_face_off)."""
    fv, fvo, cfo = [], [0], [0]
    for faces in _face_vertex_lists(verts, vert_off, ring_off, ring):
        for kept in faces:
            fv.extend(kept)
            fvo.append(len(fv))
        cfo.append(len(fvo) - 1)
    out = np.zeros((len(fv), 4), f32)
    out[:, :3] = np.asarray(fv, f32).reshape(-1, 3)
    return out, np.asarray(fvo, np.uint32), np.asarray(cfo, np.uint32)


def voronoi_cells(ctx, seeds: np.ndarray, nb_off=None, nb_idx=None) -> CellSet:
    """Voronoi cells of `seeds` in the unit container box, cut on the GPU (ctx: surtr_b200.FractureContext)."""
    if nb_off is None:
        nb_off, nb_idx = delaunay_neighbors(seeds)
    cv, cvo, cro, cr = unit_cube()
    ctx.upload_pieces(cv, cvo, cro, cr)
    ctx.upload_cells(bisector_planes(seeds, nb_off, nb_idx), nb_off)     # unbounded cells: every pair is clipped
    ctx.fracture_event()
    fr = ctx.download()
    if fr.n != len(seeds):
        raise RuntimeError("degenerate seed set: some Voronoi cell is empty")
    planes, plane_off = face_planes(fr.verts, fr.vert_off, fr.ring_off, fr.ring)
    return CellSet(fr.verts, fr.vert_off, fr.ring_off, fr.ring, planes, plane_off)


def roll_cells(cs: CellSet, r: int) -> CellSet:
    """The same pattern with its cells renumbered: new cell i = old cell (i + r) mod n (CSR arrays rotated).
    bench.py uses it to lay out many distinct resident input sets (more input than the L2 holds)."""
    n = cs.n
    r %= n
    if r == 0:
        return cs

    def rot(data, off):
        off = off.astype(np.int64)
        cut = int(off[r])
        new_off = np.concatenate([off[r:] - cut, off[1:r + 1] + (int(off[-1]) - cut)]).astype(np.uint32)
        return np.concatenate([data[cut:], data[:cut]]), new_off

    verts, vert_off = rot(cs.verts, cs.vert_off)
    ro = cs.ring_off.astype(np.int64)
    lens = np.diff(ro)
    cutv = int(cs.vert_off[r])
    ring_len = np.concatenate([lens[cutv:], lens[:cutv]])
    ring_off = np.concatenate([[0], np.cumsum(ring_len)]).astype(np.uint32)
    cute = int(ro[cutv])
    ring = np.concatenate([cs.ring[cute:], cs.ring[:cute]])
    planes, plane_off = rot(cs.planes, cs.plane_off)
    return CellSet(np.ascontiguousarray(verts), vert_off, ring_off, np.ascontiguousarray(ring),
                   np.ascontiguousarray(planes), plane_off)


def algorithmic_bytes(pieces_vert_off, pieces_ring_off, plane_off, rec) -> int:
    """compulsory traffic of the clip + assembly per surviving pair:
    16*V_in + 4*E2_in + 16*P_cell + 16*V_out + 4*E2_out + 64."""
    p = rec["piece"].astype(np.int64)
    c = rec["cell"].astype(np.int64)
    v_in = (pieces_vert_off[p + 1] - pieces_vert_off[p]).astype(np.int64)
    e_in = (pieces_ring_off[pieces_vert_off[p + 1]] - pieces_ring_off[pieces_vert_off[p]]).astype(np.int64)
    pl = (plane_off[c + 1] - plane_off[c]).astype(np.int64)
    v_out = rec["n_verts"].astype(np.int64)
    e_out = rec["n_ring"].astype(np.int64)
    return int(np.sum(16 * v_in + 4 * e_in + 16 * pl + 16 * v_out + 4 * e_out + 64))


def algorithmic_flops(pieces_vert_off, plane_off, rec) -> int:
    """FP32 operations the path needs per surviving pair -- classification 6*P*V (three
    multiplies and three adds per vertex per plane, V = mean of the piece's and the fragment's vertex counts),
    13 per new vertex (one intersection; about one new vertex per result vertex and cut generation, counted as
    2*V_out), and 24 + 60 per fan triangle for volume / centroid and inertia over 2*V_out - 4 triangles."""
    p = rec["piece"].astype(np.int64)
    c = rec["cell"].astype(np.int64)
    v_in = (pieces_vert_off[p + 1] - pieces_vert_off[p]).astype(np.int64)
    pl = (plane_off[c + 1] - plane_off[c]).astype(np.int64)
    v_out = rec["n_verts"].astype(np.int64)
    return int(np.sum(6 * pl * (v_in + v_out) // 2 + 13 * 2 * v_out + 84 * np.maximum(2 * v_out - 4, 0)))


# ------------------------------------------------------------------------------------------------ batch generation
def voronoi_cells_batch(ctx, seeds: np.ndarray, set_off: np.ndarray, planes: bool = True):
    """Voronoi cell sets of MANY seed sets at once (BASELINE config 4: 4096 x (1000 + 64) cells).  Product path only:
    Delaunay neighbours on the host worker pool (host/DT3D.cpp through hostlib), the container box cut by every cell's
    bisector half-spaces in ONE GPU event, face planes on the host worker pool (VMACH::PolygonFace semantics).
    Returns one CellSet holding all sets back to back; set s owns cells [set_off[s], set_off[s+1])."""
    from . import hostlib
    seeds = np.ascontiguousarray(seeds, f32).reshape(-1, 3)
    set_off = np.asarray(set_off, np.uint32)
    nb_off, nb_local = hostlib.dt3d_neighbors_batch(seeds, set_off)
    per_seed_base = np.repeat(set_off[:-1].astype(np.int64), np.diff(set_off.astype(np.int64)))
    nb_idx = (nb_local.astype(np.int64) + np.repeat(per_seed_base, np.diff(nb_off.astype(np.int64)))).astype(np.uint32)
    cv, cvo, cro, cr = unit_cube()
    ctx.upload_pieces(cv, cvo, cro, cr)
    ctx.upload_cells(bisector_planes(seeds, nb_off, nb_idx), nb_off)     # unbounded cells: every pair is clipped
    ctx.fracture_event()
    fr = ctx.download()
    if fr.n != len(seeds):
        raise RuntimeError("degenerate seed set: some Voronoi cell is empty")
    if planes:
        pl, po = hostlib.face_planes(fr.verts, fr.vert_off, fr.ring_off, fr.ring)
    else:
        pl, po = np.zeros((0, 4), f32), np.zeros(len(seeds) + 1, np.uint32)
    return CellSet(fr.verts, fr.vert_off, fr.ring_off, fr.ring, pl, po)


def slice_sets(cs: CellSet, c0: int, c1: int) -> CellSet:
    """Cells [c0, c1) of a CellSet as a CellSet of its own (offsets rebased; views where possible)."""
    v0, v1 = int(cs.vert_off[c0]), int(cs.vert_off[c1])
    e0, e1 = int(cs.ring_off[v0]), int(cs.ring_off[v1])
    has_planes = len(cs.planes) > 0
    p0, p1 = (int(cs.plane_off[c0]), int(cs.plane_off[c1])) if has_planes else (0, 0)
    return CellSet(cs.verts[v0:v1], (cs.vert_off[c0:c1 + 1] - np.uint32(v0)).astype(np.uint32),
                   (cs.ring_off[v0:v1 + 1] - np.uint32(e0)).astype(np.uint32), cs.ring[e0:e1], cs.planes[p0:p1],
                   (cs.plane_off[c0:c1 + 1] - np.uint32(p0)).astype(np.uint32) if has_planes else np.zeros(c1 - c0 + 1, np.uint32))


def config4_events(ctx, event_ids, n_pieces: int = 1000, n_cells: int = 64, chunk: int = 128):
    """BASELINE config 4 inputs for the given global event ids: event e cuts the Voronoi cells
    of mt19937(1234 + e) (n_pieces seeds) by the Voronoi cells of mt19937(46354 + e) (n_cells seeds).  Returns
    (pieces: CellSet without planes, cells: CellSet with planes, ev_piece_off, ev_cell_off)."""
    parts_p, parts_c = [], []
    event_ids = list(event_ids)
    for i in range(0, len(event_ids), chunk):
        ids = event_ids[i:i + chunk]
        sp = np.concatenate([seeds_uniform(1234 + e, n_pieces) for e in ids])
        sc = np.concatenate([seeds_uniform(46354 + e, n_cells) for e in ids])
        parts_p.append(voronoi_cells_batch(ctx, sp, np.arange(len(ids) + 1, dtype=np.uint32) * n_pieces, planes=False))
        parts_c.append(voronoi_cells_batch(ctx, sc, np.arange(len(ids) + 1, dtype=np.uint32) * n_cells, planes=True))
    pieces, cells = concat_sets(parts_p), concat_sets(parts_c)
    ev_p = (np.arange(len(event_ids) + 1, dtype=np.uint64) * n_pieces).astype(np.uint32)
    ev_c = (np.arange(len(event_ids) + 1, dtype=np.uint64) * n_cells).astype(np.uint32)
    return pieces, cells, ev_p, ev_c


def concat_sets(parts) -> CellSet:
    """Concatenate CellSets (offsets shifted)."""
    if len(parts) == 1:
        return parts[0]
    vo, ro, po = [np.zeros(1, np.uint64)], [np.zeros(1, np.uint64)], [np.zeros(1, np.uint64)]
    for p in parts:
        vo.append(p.vert_off[1:].astype(np.uint64) + vo[-1][-1])
        ro.append(p.ring_off[1:].astype(np.uint64) + ro[-1][-1])
        po.append(p.plane_off[1:].astype(np.uint64) + po[-1][-1])
    return CellSet(np.concatenate([p.verts for p in parts]), np.concatenate(vo).astype(np.uint32),
                   np.concatenate(ro).astype(np.uint32), np.concatenate([p.ring for p in parts]),
                   np.concatenate([p.planes for p in parts]), np.concatenate(po).astype(np.uint32))
