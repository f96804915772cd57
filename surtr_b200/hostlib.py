"""ctypes binding of the flat C entry points of libsurtr_host.so (surtr_b200/host/capi.cpp): the host-side steps of
pattern generation at batch scale -- Delaunay neighbour lists (host/DT3D.cpp) and face planes of cell polyhedra
(VMACH::PolygonFace semantics) -- for callers that are not C++.  Product code."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libsurtr_host.so")
_lib = None


def load():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(f"{LIB_PATH} is missing: build it with python -c 'import __graft_entry__ as g; g.build()'")
        from . import engine
        engine.load_library()      # libsurtr_host.so links against libsurtr_b200.so
        lib = C.CDLL(LIB_PATH)
        vp, u32, u64 = C.c_void_p, C.c_uint32, C.c_uint64
        lib.surtr_host_last_error.restype = C.c_char_p
        lib.surtr_host_dt3d_neighbors_batch.restype = u64
        lib.surtr_host_dt3d_neighbors_batch.argtypes = [vp, vp, u32, vp, vp, u64]
        lib.surtr_host_face_planes.restype = u64
        lib.surtr_host_face_planes.argtypes = [vp, vp, vp, vp, u32, vp, vp, u64]
        _lib = lib
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def dt3d_neighbors_batch(seeds: np.ndarray, set_off: np.ndarray):
    """Delaunay neighbour CSR (ascending, indices local to each seed set) of many seed sets at once."""
    lib = load()
    seeds = np.ascontiguousarray(seeds, np.float32).reshape(-1, 3)
    set_off = np.ascontiguousarray(set_off, np.uint32)
    n = int(set_off[-1])
    off = np.zeros(n + 1, np.uint32)
    idx = np.zeros(max(64, 18 * n), np.uint32)
    total = int(lib.surtr_host_dt3d_neighbors_batch(_p(seeds), _p(set_off), len(set_off) - 1, _p(off), _p(idx), len(idx)))
    if total > len(idx):
        idx = np.zeros(total, np.uint32)
        total = int(lib.surtr_host_dt3d_neighbors_batch(_p(seeds), _p(set_off), len(set_off) - 1, _p(off), _p(idx), len(idx)))
    if total == 0 and n:
        raise RuntimeError("surtr_host_dt3d_neighbors_batch: " + lib.surtr_host_last_error().decode())
    return off, idx[:total].copy()


def face_planes(verts4, vert_off, ring_off, ring):
    """(planes4 [NF,4], plane_off [n+1]) of flat polyhedra: faces in Poly::ExtractFaces order, planes as a
    VMACH::PolygonFace built by AddVertex holds them."""
    lib = load()
    verts4 = np.ascontiguousarray(verts4, np.float32)
    vert_off = np.ascontiguousarray(vert_off, np.uint32)
    ring_off = np.ascontiguousarray(ring_off, np.uint32)
    ring = np.ascontiguousarray(ring, np.uint16)
    n = len(vert_off) - 1
    plane_off = np.zeros(n + 1, np.uint32)
    cap = len(ring) // 2 + 4 * n + 16          # Euler: F = 2 - V + E per closed polyhedron
    planes = np.zeros((cap, 4), np.float32)
    total = int(lib.surtr_host_face_planes(_p(verts4), _p(vert_off), _p(ring_off), _p(ring), n, _p(planes), _p(plane_off), cap))
    if total > cap:
        planes = np.zeros((total, 4), np.float32)
        total = int(lib.surtr_host_face_planes(_p(verts4), _p(vert_off), _p(ring_off), _p(ring), n, _p(planes), _p(plane_off), total))
    return planes[:total].copy(), plane_off
