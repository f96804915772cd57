"""ctypes binding of surtr_b200/libsurtr_hosttest.so: flat entry points over the C++ host-side mirror classes
(Poly / Kdop / VMACH / DT3D / SurtrHost), used by the tests to drive the class-level API like a C++ caller."""
import ctypes as C
import os

import numpy as np

from oracle.refapi import PolySet

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB_PATH = os.path.join(ROOT, "surtr_b200", "libsurtr_hosttest.so")
_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(LIB_PATH)
        _lib.hosttest_error.restype = C.c_char_p
        _lib.hosttest_dt3d_neighbors.restype = C.c_uint64
        _lib.hosttest_dt3d_neighbors.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p, C.c_uint64]
        _lib.hosttest_kdop_ach.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32, C.c_double, C.c_float,
                                           C.c_void_p, C.c_void_p]
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _err():
    return lib().hosttest_error().decode()


def export() -> PolySet:
    L = lib()
    sizes = np.zeros(4, np.uint64)
    L.hosttest_sizes(_p(sizes))
    n, nv, ne, npl = (int(x) for x in sizes)
    ps = PolySet(np.zeros((nv, 4), np.float32), np.zeros(n + 1, np.uint32), np.zeros(nv + 1, np.uint32),
                 np.zeros(ne, np.uint16), cell=np.zeros(n, np.uint32), piece=np.zeros(n, np.uint32),
                 nfaces=np.zeros(n, np.uint32), volume=np.zeros(n, np.float64), centroid=np.zeros((n, 3), np.float32))
    ps.planes = np.zeros((npl, 4), np.float32)
    L.hosttest_export(_p(ps.verts), _p(ps.vert_off), _p(ps.ring_off), _p(ps.ring), _p(ps.cell), _p(ps.piece),
                      _p(ps.nfaces), _p(ps.volume), _p(ps.centroid), _p(ps.planes), None)
    return ps


def seeds(seed, n):
    out = np.zeros((n, 3), np.float32)
    lib().hosttest_seeds(seed, n, _p(out))
    return out


def dt3d_neighbors(s):
    s = np.ascontiguousarray(s, np.float32)
    off = np.zeros(len(s) + 1, np.uint32)
    idx = np.zeros(64 * len(s) + 64, np.uint32)
    n = lib().hosttest_dt3d_neighbors(_p(s), len(s), _p(off), _p(idx), len(idx))
    return off, idx[:int(n)].copy()


def dt3d_tets(s, use_grid=True):
    """Tets as point indices in construction order (grid-accelerated or plain-scan Bowyer-Watson)."""
    s = np.ascontiguousarray(s, np.float32)
    L = lib()
    L.hosttest_dt3d_tets.restype = C.c_uint64
    L.hosttest_dt3d_tets.argtypes = [C.c_void_p, C.c_uint32, C.c_int, C.c_void_p, C.c_uint64]
    out = np.zeros((16 * len(s) + 64, 4), np.int32)
    n = int(L.hosttest_dt3d_tets(_p(s), len(s), int(use_grid), _p(out), len(out)))
    if n > len(out):      # degenerate inputs (lattices, coplanar points) can leave many more tets: size and repeat
        out = np.zeros((n, 4), np.int32)
        n = int(L.hosttest_dt3d_tets(_p(s), len(s), int(use_grid), _p(out), len(out)))
    return out[:n].copy()


def dt3d_triangulate(s):
    s = np.ascontiguousarray(s, np.float32)
    nt, nf, ne = C.c_uint32(0), C.c_uint32(0), C.c_uint32(0)
    viol = lib().hosttest_dt3d_triangulate(_p(s), len(s), C.byref(nt), C.byref(nf), C.byref(ne))
    return nt.value, nf.value, ne.value, viol


def box_planes():
    out = np.zeros((6, 4), np.float32)
    lib().hosttest_box_planes(_p(out))
    return out


def extract_faces(verts4, ring_off, ring):
    verts4 = np.ascontiguousarray(verts4, np.float32)
    ro = np.ascontiguousarray(ring_off - ring_off[0], np.uint32)
    ring = np.ascontiguousarray(ring, np.uint16)
    fo = np.zeros(len(ring) + 2, np.uint32)
    fi = np.zeros(len(ring) + 2, np.uint16)
    n = lib().hosttest_extract_faces(_p(verts4), _p(ro), _p(ring), len(verts4), _p(fo), _p(fi))
    return [fi[fo[f]:fo[f + 1]].tolist() for f in range(n)]


def compare_plane_point(plane, p):
    plane, p = np.ascontiguousarray(plane, np.float32), np.ascontiguousarray(p, np.float32)
    return lib().hosttest_compare_plane_point(_p(plane), _p(p))


def plane_line_intersection(a, b, plane):
    a, b, plane = (np.ascontiguousarray(x, np.float32) for x in (a, b, plane))
    out = np.zeros(3, np.float32)
    lib().hosttest_plane_line_intersection(_p(a), _p(b), _p(plane), _p(out))
    return out


def voronoi_planes(s):
    s = np.ascontiguousarray(s, np.float32)
    if lib().hosttest_voronoi(_p(s), len(s)):
        raise RuntimeError(_err())
    sizes = np.zeros(4, np.uint64)
    lib().hosttest_sizes(_p(sizes))
    planes = np.zeros((int(sizes[3]), 4), np.float32)
    off = np.zeros(len(s) + 1, np.uint32)
    lib().hosttest_export(None, None, None, None, None, None, None, None, None, _p(planes), _p(off))
    return planes, off


def apply_fracture(pieces: PolySet, cells: PolySet) -> PolySet:
    rc = lib().hosttest_apply_fracture(_p(pieces.verts), _p(pieces.vert_off), _p(pieces.ring_off), _p(pieces.ring), pieces.n,
                                       _p(cells.planes), _p(cells.plane_off), _p(cells.verts), _p(cells.vert_off), cells.n)
    if rc:
        raise RuntimeError(_err())
    return export()


def export_mesh() -> PolySet:
    lib().hosttest_swap_sets()
    try:
        return export()
    finally:
        lib().hosttest_swap_sets()


def mesh_polyhedron(verts4, indices) -> PolySet:
    verts4 = np.ascontiguousarray(verts4, np.float32)
    indices = np.ascontiguousarray(indices, np.int32).reshape(-1)
    if lib().hosttest_mesh_polyhedron(_p(verts4), len(verts4), _p(indices), len(indices)):
        raise RuntimeError(_err())
    return export()


def apply_fracture_mesh(convex: PolySet, mesh: PolySet, planes, plane_off, cell_verts, cell_vert_off):
    planes, cell_verts = np.ascontiguousarray(planes, np.float32), np.ascontiguousarray(cell_verts, np.float32)
    plane_off, cell_vert_off = np.ascontiguousarray(plane_off, np.uint32), np.ascontiguousarray(cell_vert_off, np.uint32)
    rc = lib().hosttest_apply_fracture_mesh(_p(convex.verts), _p(convex.vert_off), _p(convex.ring_off), _p(convex.ring),
                                            _p(mesh.verts), _p(mesh.vert_off), _p(mesh.ring_off), _p(mesh.ring), convex.n,
                                            _p(planes), _p(plane_off), _p(cell_verts), _p(cell_vert_off), len(plane_off) - 1)
    if rc:
        raise RuntimeError(_err())
    return export(), export_mesh()


def do_fracture(convex: PolySet, mesh: PolySet, seeds, cloud, impact, impact_radius, max_axis_scale, partial):
    seeds, cloud, impact = (np.ascontiguousarray(x, np.float32) for x in (seeds, cloud, impact))
    ncomp = C.c_uint32(0)
    mass = np.zeros((4096, 10), np.float32)
    L = lib()
    L.hosttest_do_fracture.argtypes = ([C.c_void_p] * 8 + [C.c_uint32, C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32, C.c_void_p,
                                        C.c_float, C.c_float, C.c_int, C.c_void_p, C.c_void_p, C.c_uint32])
    rc = L.hosttest_do_fracture(_p(convex.verts), _p(convex.vert_off), _p(convex.ring_off), _p(convex.ring),
                                _p(mesh.verts), _p(mesh.vert_off), _p(mesh.ring_off), _p(mesh.ring), convex.n,
                                _p(seeds), len(seeds), _p(cloud), len(cloud), _p(impact), impact_radius, max_axis_scale,
                                int(partial), C.byref(ncomp), _p(mass), len(mass))
    if rc:
        raise RuntimeError(_err())
    return export(), export_mesh(), ncomp.value, mass[:ncomp.value].copy()


def last_do_fracture_ms() -> float:
    """Wall time of SurtrHost::DoFracture itself inside the last do_fracture call (no wrapper conversions)."""
    L = lib()
    L.hosttest_last_do_fracture_ms.restype = C.c_double
    return float(L.hosttest_last_do_fracture_ms())


def transform(verts4, matrix16):
    verts4 = np.ascontiguousarray(verts4, np.float32)
    m = np.ascontiguousarray(matrix16, np.float32).reshape(16)
    out = np.zeros_like(verts4)
    L = lib()
    L.hosttest_transform.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p]
    L.hosttest_transform(_p(verts4), len(verts4), _p(m), _p(out))
    return out


def combine_mass(volume, centroid, inertia, density=10.0):
    volume = np.ascontiguousarray(volume, np.float64)
    centroid, inertia = np.ascontiguousarray(centroid, np.float32), np.ascontiguousarray(inertia, np.float32)
    out = np.zeros(10, np.float32)
    L = lib()
    L.hosttest_combine_mass.argtypes = [C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_float, C.c_void_p]
    L.hosttest_combine_mass(len(volume), _p(volume), _p(centroid), _p(inertia), density, _p(out))
    return out


def config1_full(verts4, indices, seeds):
    verts4, seeds = np.ascontiguousarray(verts4, np.float32), np.ascontiguousarray(seeds, np.float32)
    indices = np.ascontiguousarray(indices, np.int32).reshape(-1)
    ach_nv = C.c_uint32(0)
    rc = lib().hosttest_config1_full(_p(verts4), len(verts4), _p(indices), len(indices), _p(seeds), len(seeds), C.byref(ach_nv))
    if rc:
        raise RuntimeError(_err())
    return export(), export_mesh(), ach_nv.value


def clip_and_moments(ps: PolySet, planes):
    planes = np.ascontiguousarray(planes, np.float32)
    vol = C.c_double(0)
    cen = np.zeros(3, np.float32)
    rc = lib().hosttest_clip_and_moments(_p(ps.verts), _p(ps.ring_off), _p(ps.ring), len(ps.verts), _p(planes), len(planes),
                                         C.byref(vol), _p(cen))
    if rc:
        raise RuntimeError(_err())
    return export(), vol.value, cen


def kdop_ach(verts4, normals, max_axis, gap_inv, boxverts4):
    verts4, normals, boxverts4 = (np.ascontiguousarray(x, np.float32) for x in (verts4, normals, boxverts4))
    planes = np.zeros((len(normals), 2, 4), np.float32)
    rc = lib().hosttest_kdop_ach(_p(verts4), len(verts4), _p(normals), len(normals), max_axis, gap_inv, _p(boxverts4), _p(planes))
    if rc:
        raise RuntimeError(_err())
    return planes, export()


def ich_normals(verts4, limit):
    verts4 = np.ascontiguousarray(verts4, np.float32)
    out = np.zeros((4096, 3), np.float32)
    L = lib()
    L.hosttest_ich_normals.restype = C.c_uint32
    L.hosttest_ich_normals.argtypes = [C.c_void_p, C.c_uint32, C.c_int, C.c_void_p, C.c_uint32]
    n = L.hosttest_ich_normals(_p(verts4), len(verts4), limit, _p(out), len(out))
    return out[:n].copy()


def refit(convex: PolySet, mesh_verts4, mesh_vert_off, limit=4) -> PolySet:
    mesh_verts4 = np.ascontiguousarray(mesh_verts4, np.float32)
    mesh_vert_off = np.ascontiguousarray(mesh_vert_off, np.uint32)
    L = lib()
    L.hosttest_refit.argtypes = [C.c_void_p] * 4 + [C.c_uint32, C.c_void_p, C.c_void_p, C.c_int]
    rc = L.hosttest_refit(_p(convex.verts), _p(convex.vert_off), _p(convex.ring_off), _p(convex.ring), convex.n,
                          _p(mesh_verts4), _p(mesh_vert_off), limit)
    if rc:
        raise RuntimeError(_err())
    return export()


def config1(verts4, seeds) -> PolySet:
    verts4, seeds = np.ascontiguousarray(verts4, np.float32), np.ascontiguousarray(seeds, np.float32)
    L = lib()
    L.hosttest_config1.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32, C.c_void_p]
    ach_nv = C.c_uint32(0)
    if L.hosttest_config1(_p(verts4), len(verts4), _p(seeds), len(seeds), C.byref(ach_nv)):
        raise RuntimeError(_err())
    ps = export()
    ps.ach_nv = ach_nv.value
    return ps
