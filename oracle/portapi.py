"""ctypes binding of oracle/_build/libsurtr_oracle.so (the plain-C restatement) -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from .refapi import PolySet

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_build", "libsurtr_oracle.so")
_lib = None


class _SoOut(C.Structure):
    _fields_ = [("verts4", C.c_void_p), ("cap_verts", C.c_uint64), ("vert_off", C.c_void_p),
                ("ring_off", C.c_void_p), ("ring", C.c_void_p), ("cap_ring", C.c_uint64),
                ("rec", C.c_void_p), ("cap_frags", C.c_uint64), ("volume", C.c_void_p),
                ("centroid", C.c_void_p), ("inertia", C.c_void_p)]


def build() -> str:
    src = [os.path.join(_HERE, f) for f in ("surtr_oracle.c", "surtr_oracle.h", "Makefile")]
    if (not os.path.exists(LIB_PATH)) or any(os.path.getmtime(s) > os.path.getmtime(LIB_PATH) for s in src):
        subprocess.run(["make", "-C", _HERE, "port"], check=True, capture_output=True)
    return LIB_PATH


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(LIB_PATH)
        _lib.so_compare_plane_point.argtypes = [C.c_void_p, C.c_void_p]
        _lib.so_compare_plane_point.restype = C.c_int
        _lib.so_plane_line_intersection.argtypes = [C.c_void_p] * 4
        _lib.so_plane_from_points.argtypes = [C.c_void_p] * 4
        _lib.so_plane_from_point_normal.argtypes = [C.c_void_p] * 3
        _lib.so_kdop_calc.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32] + [C.c_void_p] * 3
        _lib.so_apply_fracture.argtypes = ([C.c_void_p] * 4 + [C.c_uint32, C.c_void_p, C.c_void_p, C.c_uint32,
                                                                C.POINTER(_SoOut)])
        _lib.so_apply_fracture.restype = C.c_int64
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def compare_plane_point(plane, p) -> int:
    plane = np.ascontiguousarray(plane, np.float32)
    p = np.ascontiguousarray(p, np.float32)
    return lib().so_compare_plane_point(_p(plane), _p(p))


def plane_line_intersection(a, b, plane):
    a, b, plane = (np.ascontiguousarray(x, np.float32) for x in (a, b, plane))
    out = np.zeros(3, np.float32)
    lib().so_plane_line_intersection(_p(a), _p(b), _p(plane), _p(out))
    return out


def plane_from_points(a, b, c):
    a, b, c = (np.ascontiguousarray(x, np.float32) for x in (a, b, c))
    out = np.zeros(4, np.float32)
    lib().so_plane_from_points(_p(a), _p(b), _p(c), _p(out))
    return out


def plane_from_point_normal(a, n):
    a, n = (np.ascontiguousarray(x, np.float32) for x in (a, n))
    out = np.zeros(4, np.float32)
    lib().so_plane_from_point_normal(_p(a), _p(n), _p(out))
    return out


def kdop_calc(verts4, normals):
    verts4 = np.ascontiguousarray(verts4, np.float32)
    normals = np.ascontiguousarray(normals, np.float32)
    k = len(normals)
    dist = np.zeros((k, 2), np.float32)
    arg = np.zeros((k, 2), np.int32)
    planes = np.zeros((k, 2, 4), np.float32)
    lib().so_kdop_calc(_p(verts4), len(verts4), _p(normals), k, _p(dist), _p(arg), _p(planes))
    return dist, arg, planes


def apply_fracture(pieces: PolySet, planes, plane_off, moments=True, inertia=False,
                   cap_frags=None, cap_verts=None) -> PolySet:
    """All cells x all pieces, cell-major / piece-minor (Surtr.cpp:2098-2149).  Returns a PolySet with
    cell/piece/nfaces/volume/centroid filled (plus .inertia when requested)."""
    planes = np.ascontiguousarray(planes, np.float32)
    plane_off = np.ascontiguousarray(plane_off, np.uint32)
    n_cells = len(plane_off) - 1
    cap_frags = cap_frags or max(1024, 8 * (pieces.n + n_cells))
    while True:
        cv = cap_verts or cap_frags * 48
        cr = cv * 4
        verts = np.zeros((cv, 4), np.float32)
        vert_off = np.zeros(cap_frags + 1, np.uint32)
        ring_off = np.zeros(cv + 1, np.uint32)
        ring = np.zeros(cr, np.uint16)
        rec = np.zeros((cap_frags, 4), np.uint32)
        vol = np.zeros(cap_frags, np.float64) if moments else None
        cen = np.zeros((cap_frags, 3), np.float32) if moments else None
        ine = np.zeros((cap_frags, 6), np.float64) if inertia else None
        o = _SoOut(_p(verts), cv, _p(vert_off), _p(ring_off), _p(ring), cr, _p(rec), cap_frags, _p(vol), _p(cen),
                   _p(ine))
        n = lib().so_apply_fracture(_p(pieces.verts), _p(pieces.vert_off), _p(pieces.ring_off), _p(pieces.ring),
                                    pieces.n, _p(planes), _p(plane_off), n_cells, C.byref(o))
        if n == -1:
            cap_frags *= 4
            cap_verts = None
            continue
        if n < 0:
            raise RuntimeError(f"so_apply_fracture failed: {n}")
        break
    n = int(n)
    nv = int(vert_off[n])
    ne = int(ring_off[nv])
    ps = PolySet(verts[:nv].copy(), vert_off[:n + 1].copy(), ring_off[:nv + 1].copy(), ring[:ne].copy(),
                 cell=rec[:n, 0].copy(), piece=rec[:n, 1].copy(), nfaces=rec[:n, 3].copy(),
                 volume=None if vol is None else vol[:n].copy(),
                 centroid=None if cen is None else cen[:n].copy())
    ps.inertia = None if ine is None else ine[:n].copy()
    return ps


def _alloc_out(cap_frags, cap_verts, moments=True, inertia=False):
    cr = cap_verts * 4
    bufs = dict(
        verts=np.zeros((cap_verts, 4), np.float32), vert_off=np.zeros(cap_frags + 1, np.uint32),
        ring_off=np.zeros(cap_verts + 1, np.uint32), ring=np.zeros(cr, np.uint16),
        rec=np.zeros((cap_frags, 4), np.uint32),
        vol=np.zeros(cap_frags, np.float64) if moments else None,
        cen=np.zeros((cap_frags, 3), np.float32) if moments else None,
        ine=np.zeros((cap_frags, 6), np.float64) if inertia else None)
    o = _SoOut(_p(bufs["verts"]), cap_verts, _p(bufs["vert_off"]), _p(bufs["ring_off"]), _p(bufs["ring"]), cr,
               _p(bufs["rec"]), cap_frags, _p(bufs["vol"]), _p(bufs["cen"]), _p(bufs["ine"]))
    return o, bufs


def _collect(n, b) -> PolySet:
    nv = int(b["vert_off"][n])
    ne = int(b["ring_off"][nv])
    ps = PolySet(b["verts"][:nv].copy(), b["vert_off"][:n + 1].copy(), b["ring_off"][:nv + 1].copy(),
                 b["ring"][:ne].copy(), cell=b["rec"][:n, 0].copy(), piece=b["rec"][:n, 1].copy(),
                 nfaces=b["rec"][:n, 3].copy(),
                 volume=None if b["vol"] is None else b["vol"][:n].copy(),
                 centroid=None if b["cen"] is None else b["cen"][:n].copy())
    ps.inertia = None if b["ine"] is None else b["ine"][:n].copy()
    return ps


def voronoi_cells(seeds, nb_off, nb_idx) -> PolySet:
    """Cells of `seeds` in the unit box from neighbour lists (the NEW derivation, see surtr_oracle.h)."""
    L = lib()
    L.so_voronoi_cells.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p, C.POINTER(_SoOut), C.c_void_p,
                                   C.c_void_p, C.c_uint64]
    L.so_voronoi_cells.restype = C.c_int64
    seeds = np.ascontiguousarray(seeds, np.float32)
    nb_off = np.ascontiguousarray(nb_off, np.uint32)
    nb_idx = np.ascontiguousarray(nb_idx, np.uint32)
    n = len(seeds)
    o, b = _alloc_out(n, n * 96)
    cap_pl = n * 96
    planes = np.zeros((cap_pl, 4), np.float32)
    plane_off = np.zeros(n + 1, np.uint32)
    r = L.so_voronoi_cells(_p(seeds), n, _p(nb_off), _p(nb_idx), C.byref(o), _p(plane_off), _p(planes), cap_pl)
    if r < 0:
        raise RuntimeError(f"so_voronoi_cells failed: {r}")
    ps = _collect(int(r), b)
    ps.planes = planes[:int(plane_off[n])].copy()
    ps.poly_face_off = plane_off
    return ps


def clip_each(ps: PolySet, planes, pl_off) -> PolySet:
    L = lib()
    L.so_clip_each.argtypes = [C.c_void_p] * 4 + [C.c_uint32, C.c_void_p, C.c_void_p, C.POINTER(_SoOut)]
    L.so_clip_each.restype = C.c_int64
    planes = np.ascontiguousarray(planes, np.float32)
    pl_off = np.ascontiguousarray(pl_off, np.uint32)
    nv_in = int(ps.vert_off[-1])
    o, b = _alloc_out(ps.n + 1, nv_in + 64 * (ps.n + 1) + int(len(planes)) * 4)
    r = L.so_clip_each(_p(ps.verts), _p(ps.vert_off), _p(ps.ring_off), _p(ps.ring), ps.n, _p(planes), _p(pl_off),
                       C.byref(o))
    if r < 0:
        raise RuntimeError(f"so_clip_each failed: {r}")
    return _collect(int(r), b)


def face_planes(ps: PolySet):
    """(planes4, plane_off) of every polyhedron, PolygonFace::AddVertex route (VMACH.cpp:289-310)."""
    L = lib()
    L.so_face_planes_set.argtypes = [C.c_void_p] * 4 + [C.c_uint32, C.c_void_p, C.c_void_p, C.c_uint64]
    L.so_face_planes_set.restype = C.c_int64
    cap = int(len(ps.ring)) + 16
    planes = np.zeros((cap, 4), np.float32)
    off = np.zeros(ps.n + 1, np.uint32)
    r = L.so_face_planes_set(_p(ps.verts), _p(ps.vert_off), _p(ps.ring_off), _p(ps.ring), ps.n, _p(off), _p(planes), cap)
    if r < 0:
        raise RuntimeError("so_face_planes_set overflow")
    return planes[:int(r)].copy(), off
