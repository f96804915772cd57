"""ctypes binding of oracle/_ref/libsurtr_ref.so -- TEST INFRASTRUCTURE ONLY.

The library is the REFERENCE's own geometry code (Src/Poly.cpp, Src/Kdop.cpp, Src/VMACH.cpp,
Inc/DT3D.h) compiled headless by oracle/Makefile plus the thin driver oracle/ref_driver.cpp.
Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may import this module.
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass, field

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_ref", "libsurtr_ref.so")

_lib = None


def available() -> bool:
    return os.path.exists(LIB_PATH)


def lib():
    global _lib
    if _lib is None:
        if not available():
            raise RuntimeError(f"{LIB_PATH} missing: run `make -C oracle ref` where /root/reference exists")
        _lib = C.CDLL(LIB_PATH)
        _lib.ref_polyset_new.restype = C.c_void_p
        _lib.ref_polyset_free.argtypes = [C.c_void_p]
        _lib.ref_polyset_sizes.argtypes = [C.c_void_p, C.c_void_p]
        _lib.ref_polyset_seconds.argtypes = [C.c_void_p]
        _lib.ref_polyset_seconds.restype = C.c_double
        _lib.ref_polyset_export.argtypes = [C.c_void_p] * 14
        _lib.ref_dt3d_neighbors.restype = C.c_uint64
        _lib.ref_dt3d_neighbors.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p]
        _lib.ref_voronoi_cells.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p]
        _lib.ref_clip_each.argtypes = [C.c_void_p] * 4 + [C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p]
        _lib.ref_apply_fracture.argtypes = ([C.c_void_p] * 4 + [C.c_uint32, C.c_void_p, C.c_void_p, C.c_uint32,
                                                                 C.c_uint32, C.c_int, C.c_void_p])
        _lib.ref_refit.argtypes = [C.c_void_p] * 4 + [C.c_uint32, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
        _lib.ref_apply_fracture_mesh.argtypes = [C.c_void_p] * 8 + [C.c_uint32, C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p]
        _lib.ref_mesh_polyhedron.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32, C.c_void_p]
        _lib.ref_config1_full.argtypes = ([C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32, C.c_int, C.c_float, C.c_int, C.c_void_p,
                                           C.c_uint32] + [C.c_void_p] * 5)
        _lib.ref_do_fracture.argtypes = ([C.c_void_p] * 8 + [C.c_uint32, C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p,
                                          C.c_uint32, C.c_void_p, C.c_float, C.c_float, C.c_int, C.c_int, C.c_void_p, C.c_void_p,
                                          C.c_void_p])
        _lib.ref_seeds_uniform.argtypes = [C.c_uint32, C.c_uint32, C.c_void_p]
        _lib.ref_seeds_radial.argtypes = [C.c_uint32, C.c_uint32, C.c_double, C.c_void_p]
        _lib.ref_unit_cube.argtypes = [C.c_void_p]
        _lib.ref_kdop_calc_poly.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32] + [C.c_void_p] * 3
        _lib.ref_kdop_calc_gap.argtypes = ([C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32, C.c_double, C.c_float]
                                           + [C.c_void_p] * 3)
        _lib.ref_ich_normals.argtypes = [C.c_void_p, C.c_uint32, C.c_int, C.c_void_p, C.c_uint32]
        _lib.ref_ich_normals.restype = C.c_uint32
        _lib.ref_compare_plane_point.argtypes = [C.c_void_p, C.c_void_p]
        _lib.ref_compare_plane_point.restype = C.c_int
        _lib.ref_plane_line_intersection.argtypes = [C.c_void_p] * 4
        _lib.ref_plane_from_points.argtypes = [C.c_void_p] * 4
        _lib.ref_plane_from_point_normal.argtypes = [C.c_void_p] * 3
        _lib.ref_box_planes.argtypes = [C.c_void_p]
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


@dataclass
class PolySet:
    """Flat set of polyhedra (layout documented in include/surtr_b200.h)."""
    verts: np.ndarray      # float32 [NV,4]
    vert_off: np.ndarray   # uint32 [n+1]
    ring_off: np.ndarray   # uint32 [NV+1]
    ring: np.ndarray       # uint16 [NE]
    cell: np.ndarray = None
    piece: np.ndarray = None
    nfaces: np.ndarray = None
    volume: np.ndarray = None
    centroid: np.ndarray = None
    poly_face_off: np.ndarray = None
    face_off: np.ndarray = None
    face_idx: np.ndarray = None
    planes: np.ndarray = None   # float32 [NF,4] : one plane per face, PolygonFace::AddVertex route
    seconds: float = 0.0

    @property
    def n(self) -> int:
        return len(self.vert_off) - 1

    @property
    def nverts(self) -> np.ndarray:
        return np.diff(self.vert_off).astype(np.uint32)

    @property
    def plane_off(self) -> np.ndarray:
        return self.poly_face_off

    def poly(self, i):
        v0, v1 = int(self.vert_off[i]), int(self.vert_off[i + 1])
        rings = [self.ring[self.ring_off[v]:self.ring_off[v + 1]].tolist() for v in range(v0, v1)]
        return self.verts[v0:v1, :3].copy(), rings

    def faces(self, i):
        f0, f1 = int(self.poly_face_off[i]), int(self.poly_face_off[i + 1])
        return [self.face_idx[self.face_off[f]:self.face_off[f + 1]].tolist() for f in range(f0, f1)]

    def subset(self, idx) -> "PolySet":
        """Polyhedra `idx` re-packed (geometry only)."""
        verts, vo, ro, ring = [], [0], [0], []
        for i in idx:
            v0, v1 = int(self.vert_off[i]), int(self.vert_off[i + 1])
            verts.append(self.verts[v0:v1])
            r0, r1 = int(self.ring_off[v0]), int(self.ring_off[v1])
            ring.append(self.ring[r0:r1])
            ro.extend((self.ring_off[v0 + 1:v1 + 1] - r0 + ro[-1]).tolist())
            vo.append(vo[-1] + v1 - v0)
        return PolySet(np.concatenate(verts) if verts else np.zeros((0, 4), np.float32),
                       np.asarray(vo, np.uint32), np.asarray(ro, np.uint32),
                       np.concatenate(ring) if ring else np.zeros(0, np.uint16))


def _export(h) -> PolySet:
    L = lib()
    sizes = np.zeros(5, np.uint64)
    L.ref_polyset_sizes(h, _p(sizes))
    n, nv, ne, nf, nfi = (int(x) for x in sizes)
    ps = PolySet(
        verts=np.zeros((nv, 4), np.float32), vert_off=np.zeros(n + 1, np.uint32),
        ring_off=np.zeros(nv + 1, np.uint32), ring=np.zeros(ne, np.uint16),
        cell=np.zeros(n, np.uint32), piece=np.zeros(n, np.uint32), nfaces=np.zeros(n, np.uint32),
        volume=np.zeros(n, np.float64), centroid=np.zeros((n, 3), np.float32),
        poly_face_off=np.zeros(n + 1, np.uint32), face_off=np.zeros(nf + 1, np.uint32),
        face_idx=np.zeros(nfi, np.uint16), planes=np.zeros((nf, 4), np.float32))
    L.ref_polyset_export(h, _p(ps.verts), _p(ps.vert_off), _p(ps.ring_off), _p(ps.ring), _p(ps.cell), _p(ps.piece),
                         _p(ps.nfaces), _p(ps.volume), _p(ps.centroid), _p(ps.poly_face_off), _p(ps.face_off),
                         _p(ps.face_idx), _p(ps.planes))
    ps.seconds = L.ref_polyset_seconds(h)
    L.ref_polyset_free(h)
    return ps


def seeds_uniform(seed: int, n: int) -> np.ndarray:
    out = np.zeros((n, 3), np.float32)
    lib().ref_seeds_uniform(seed, n, _p(out))
    return out


def seeds_radial(seed: int, n: int, mean: float) -> np.ndarray:
    out = np.zeros((n, 3), np.float32)
    lib().ref_seeds_radial(seed, n, mean, _p(out))
    return out


def unit_cube() -> PolySet:
    h = lib().ref_polyset_new()
    lib().ref_unit_cube(h)
    return _export(h)


def dt3d_neighbors(seeds: np.ndarray):
    seeds = np.ascontiguousarray(seeds, np.float32)
    n = len(seeds)
    off = np.zeros(n + 1, np.uint32)
    sec = C.c_double(0)
    total = lib().ref_dt3d_neighbors(_p(seeds), n, _p(off), None, C.byref(sec))
    idx = np.zeros(int(total), np.uint32)
    lib().ref_dt3d_neighbors(_p(seeds), n, _p(off), _p(idx), None)
    return off, idx, sec.value


def voronoi_cells(seeds: np.ndarray, nb_off=None, nb_idx=None) -> PolySet:
    seeds = np.ascontiguousarray(seeds, np.float32)
    h = lib().ref_polyset_new()
    lib().ref_voronoi_cells(_p(seeds), len(seeds), _p(nb_off), _p(nb_idx), h)
    return _export(h)


def clip_each(ps: PolySet, planes: np.ndarray, pl_off: np.ndarray) -> PolySet:
    planes = np.ascontiguousarray(planes, np.float32)
    pl_off = np.ascontiguousarray(pl_off, np.uint32)
    h = lib().ref_polyset_new()
    lib().ref_clip_each(_p(ps.verts), _p(ps.vert_off), _p(ps.ring_off), _p(ps.ring), ps.n, _p(planes), _p(pl_off), h)
    return _export(h)


def apply_fracture(pieces: PolySet, planes: np.ndarray, plane_off: np.ndarray, nthreads: int = 0,
                   moments: bool = True) -> PolySet:
    planes = np.ascontiguousarray(planes, np.float32)
    plane_off = np.ascontiguousarray(plane_off, np.uint32)
    h = lib().ref_polyset_new()
    lib().ref_apply_fracture(_p(pieces.verts), _p(pieces.vert_off), _p(pieces.ring_off), _p(pieces.ring), pieces.n,
                             _p(planes), _p(plane_off), len(plane_off) - 1, nthreads, int(moments), h)
    return _export(h)


def kdop_calc_poly(verts4: np.ndarray, normals: np.ndarray):
    verts4 = np.ascontiguousarray(verts4, np.float32)
    normals = np.ascontiguousarray(normals, np.float32)
    k = len(normals)
    dist = np.zeros((k, 2), np.float64)
    planes = np.zeros((k, 2, 4), np.float32)
    vtx = np.zeros((k, 2, 3), np.float32)
    lib().ref_kdop_calc_poly(_p(verts4), len(verts4), _p(normals), k, _p(dist), _p(planes), _p(vtx))
    return dist, planes, vtx


def kdop_calc_gap(verts4: np.ndarray, normals: np.ndarray, max_axis_scale: float, plane_gap_inv: float):
    verts4 = np.ascontiguousarray(verts4, np.float32)
    normals = np.ascontiguousarray(normals, np.float32)
    k = len(normals)
    dist = np.zeros((k, 2), np.float64)
    planes = np.zeros((k, 2, 4), np.float32)
    vtx = np.zeros((k, 2, 3), np.float32)
    lib().ref_kdop_calc_gap(_p(verts4), len(verts4), _p(normals), k, max_axis_scale, plane_gap_inv,
                            _p(dist), _p(planes), _p(vtx))
    return dist, planes, vtx


def ich_normals(verts4: np.ndarray, limit: int) -> np.ndarray:
    verts4 = np.ascontiguousarray(verts4, np.float32)
    out = np.zeros((4096, 3), np.float32)
    n = lib().ref_ich_normals(_p(verts4), len(verts4), limit, _p(out), len(out))
    return out[:n].copy()


def compare_plane_point(plane, p) -> int:
    plane = np.ascontiguousarray(plane, np.float32)
    p = np.ascontiguousarray(p, np.float32)
    return lib().ref_compare_plane_point(_p(plane), _p(p))


def plane_line_intersection(a, b, plane) -> np.ndarray:
    a, b, plane = (np.ascontiguousarray(x, np.float32) for x in (a, b, plane))
    out = np.zeros(3, np.float32)
    lib().ref_plane_line_intersection(_p(a), _p(b), _p(plane), _p(out))
    return out


def plane_from_points(a, b, c) -> np.ndarray:
    a, b, c = (np.ascontiguousarray(x, np.float32) for x in (a, b, c))
    out = np.zeros(4, np.float32)
    lib().ref_plane_from_points(_p(a), _p(b), _p(c), _p(out))
    return out


def plane_from_point_normal(a, n) -> np.ndarray:
    a, n = (np.ascontiguousarray(x, np.float32) for x in (a, n))
    out = np.zeros(4, np.float32)
    lib().ref_plane_from_point_normal(_p(a), _p(n), _p(out))
    return out


def box_planes() -> np.ndarray:
    out = np.zeros((6, 4), np.float32)
    lib().ref_box_planes(_p(out))
    return out


def refit(convex: PolySet, mesh_verts4: np.ndarray, mesh_vert_off: np.ndarray, limit: int = 4) -> PolySet:
    """m_refittingTask (Surtr.cpp:1449-1455) per piece with the reference's own ConvexHull / Kdop / clipper."""
    mesh_verts4 = np.ascontiguousarray(mesh_verts4, np.float32)
    mesh_vert_off = np.ascontiguousarray(mesh_vert_off, np.uint32)
    h = lib().ref_polyset_new()
    lib().ref_refit(_p(convex.verts), _p(convex.vert_off), _p(convex.ring_off), _p(convex.ring), convex.n,
                    _p(mesh_verts4), _p(mesh_vert_off), limit, h)
    return _export(h)


def config1_convex(verts4, seeds, nb_off, nb_idx, ich_limit=20, gap_inv=2000.0):
    """PrepareFracture's convex branch on a vertex cloud (see ref_driver.cpp).  Returns (ACH, fragments)."""
    verts4 = np.ascontiguousarray(verts4, np.float32)
    seeds = np.ascontiguousarray(seeds, np.float32)
    nb_off = np.ascontiguousarray(nb_off, np.uint32)
    nb_idx = np.ascontiguousarray(nb_idx, np.uint32)
    L = lib()
    L.ref_config1_convex.argtypes = [C.c_void_p, C.c_uint32, C.c_int, C.c_float, C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p,
                                     C.c_void_p, C.c_void_p]
    h1, h2 = L.ref_polyset_new(), L.ref_polyset_new()
    L.ref_config1_convex(_p(verts4), len(verts4), ich_limit, gap_inv, _p(seeds), len(seeds), _p(nb_off), _p(nb_idx), h1, h2)
    return _export(h1), _export(h2)


def mesh_polyhedron(verts4, indices) -> PolySet:
    """Poly::ExtractNeighborFromMesh + InitPolyhedron (Surtr.cpp:1788-1795) on a triangle mesh."""
    verts4 = np.ascontiguousarray(verts4, np.float32)
    indices = np.ascontiguousarray(indices, np.int32).reshape(-1)
    L = lib()
    h = L.ref_polyset_new()
    if L.ref_mesh_polyhedron(_p(verts4), len(verts4), _p(indices), len(indices), h):
        L.ref_polyset_free(h)
        raise RuntimeError("ExtractNeighborFromMesh: asymmetric adjacency")
    return _export(h)


def apply_fracture_mesh(convex: PolySet, mesh: PolySet, planes, plane_off):
    """Full m_fractureTask (Surtr.cpp:1457-1504) per cell: convex clip, mesh clip, island split.
    Returns (Piece::Convex set, Piece::Mesh set), both in PieceVec order."""
    assert convex.n == mesh.n
    planes = np.ascontiguousarray(planes, np.float32)
    plane_off = np.ascontiguousarray(plane_off, np.uint32)
    L = lib()
    hc, hm = L.ref_polyset_new(), L.ref_polyset_new()
    L.ref_apply_fracture_mesh(_p(convex.verts), _p(convex.vert_off), _p(convex.ring_off), _p(convex.ring),
                              _p(mesh.verts), _p(mesh.vert_off), _p(mesh.ring_off), _p(mesh.ring), convex.n,
                              _p(planes), _p(plane_off), len(plane_off) - 1, hc, hm)
    return _export(hc), _export(hm)


def config1_full(verts4, indices, seeds, nb_off, nb_idx, ich_limit=20, gap_inv=2000.0, refit_limit=4):
    """Surtr::PrepareFracture in full (Surtr.cpp:1747-1827).  Returns (ACH, Piece::Convex set after Refitting,
    Piece::Mesh set)."""
    verts4, seeds = np.ascontiguousarray(verts4, np.float32), np.ascontiguousarray(seeds, np.float32)
    indices = np.ascontiguousarray(indices, np.int32).reshape(-1)
    nb_off, nb_idx = np.ascontiguousarray(nb_off, np.uint32), np.ascontiguousarray(nb_idx, np.uint32)
    L = lib()
    ha, hc, hm = L.ref_polyset_new(), L.ref_polyset_new(), L.ref_polyset_new()
    L.ref_config1_full(_p(verts4), len(verts4), _p(indices), len(indices), ich_limit, gap_inv, refit_limit, _p(seeds), len(seeds),
                       _p(nb_off), _p(nb_idx), ha, hc, hm)
    return _export(ha), _export(hc), _export(hm)


def do_fracture(convex: PolySet, mesh: PolySet, seeds, nb_off, nb_idx, cloud, impact, impact_radius, max_axis_scale,
                partial: bool, refit_limit: int = 4):
    """Surtr::DoFracture (Surtr.cpp:1885-1959) on one compound.  Returns (Piece::Convex set, Piece::Mesh set,
    number of compounds); .cell = compound index of every piece, .piece = 1 for the caller's untouched pieces."""
    seeds, cloud = np.ascontiguousarray(seeds, np.float32), np.ascontiguousarray(cloud, np.float32)
    nb_off, nb_idx = np.ascontiguousarray(nb_off, np.uint32), np.ascontiguousarray(nb_idx, np.uint32)
    impact = np.ascontiguousarray(impact, np.float32)
    L = lib()
    hc, hm = L.ref_polyset_new(), L.ref_polyset_new()
    ncomp = C.c_uint32(0)
    L.ref_do_fracture(_p(convex.verts), _p(convex.vert_off), _p(convex.ring_off), _p(convex.ring),
                      _p(mesh.verts), _p(mesh.vert_off), _p(mesh.ring_off), _p(mesh.ring), convex.n,
                      _p(seeds), len(seeds), _p(nb_off), _p(nb_idx), _p(cloud), len(cloud), _p(impact),
                      impact_radius, max_axis_scale, int(partial), refit_limit, hc, hm, C.byref(ncomp))
    return _export(hc), _export(hm), ncomp.value


def transform(verts4, matrix16) -> np.ndarray:
    """Poly::Transform (Poly.cpp:580-585): row-major world matrix, transposed inside like the reference does."""
    verts4 = np.ascontiguousarray(verts4, np.float32)
    m = np.ascontiguousarray(matrix16, np.float32).reshape(16)
    out = np.zeros_like(verts4)
    L = lib()
    L.ref_transform.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p]
    L.ref_transform(_p(verts4), len(verts4), _p(m), _p(out))
    return out
