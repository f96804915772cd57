"""ctypes binding of libsurtr_b200.so -- the C ABI declared in include/surtr_b200.h.

This is the thin Python face of the product path (used by bench.py, the tests and __graft_entry__).
It never computes anything itself and has no CPU fallback: if the CUDA library is missing or no sm_100
device is present, construction raises.
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libsurtr_b200.so")

SURTR_OK = 0
ERRORS = {1: "SURTR_ERR_CUDA", 2: "SURTR_ERR_INVALID", 3: "SURTR_ERR_NO_DEVICE", 4: "SURTR_ERR_OVERFLOW",
          5: "SURTR_ERR_NOMEM"}

# include/surtr_b200.h: struct surtr_fragment (64 bytes)
FRAGMENT_DTYPE = np.dtype([
    ("cell", np.uint32), ("piece", np.uint32), ("vert_off", np.uint32), ("n_verts", np.uint16),
    ("n_faces", np.uint16), ("volume", np.float64), ("centroid", np.float32, 3), ("inertia", np.float32, 6),
    ("n_ring", np.uint32)], align=True)
assert FRAGMENT_DTYPE.itemsize == 64

# every symbol include/surtr_b200.h declares
EXPORTS = [
    "surtr_ctx_create", "surtr_ctx_destroy", "surtr_last_error", "surtr_version", "surtr_set_kdop_directions", "surtr_set_clip_build",
    "surtr_upload_pieces", "surtr_upload_cells", "surtr_fragments_to_pieces", "surtr_fragments_to_pieces_per_event", "surtr_fracture_event",
    "surtr_event_counts", "surtr_download_fragments", "surtr_device_fragments", "surtr_kdop_calc",
    "surtr_last_event_ms", "surtr_last_event_launches", "surtr_set_profiling", "surtr_kdop_calc_batch",
    "surtr_upload_pattern", "surtr_place_pattern", "surtr_download_fragments_async", "surtr_sync",
    "surtr_transform_pieces", "surtr_download_pieces", "surtr_measure_fp32_peak",
    "surtr_last_event_phases", "surtr_failed_pairs", "surtr_input_blob_layout", "surtr_upload_blob", "surtr_download_blob_async", "surtr_upload_pieces3", "surtr_upload_cells3", "surtr_download_fragments_packed", "surtr_download_fragments_packed_async",
]


class SurtrError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"{ERRORS.get(code, code)}: {msg}")
        self.code = code


class Counts(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in ("n_pairs", "n_candidates", "n_fragments", "n_verts", "n_ring",
                                          "n_seq_cuts", "n_tier2", "n_tier3", "n_tier1b", "n_failed")]


class InLayout(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in ("verts3", "vert_off", "ring_base", "ring_len", "ring", "planes4", "plane_off", "cell_verts3",
                                          "cvert_off", "ev_piece_off", "ev_cell_off", "total")]


class OutLayout(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in ("fragments", "verts3", "ring_len", "ring", "total", "n_fragments", "n_verts", "n_ring",
                                          "ring_entry_bytes")]


class DeviceView(C.Structure):
    _fields_ = [("fragments", C.c_void_p), ("verts4", C.c_void_p), ("ring_off", C.c_void_p), ("ring", C.c_void_p)]


_lib = None


def load_library():
    """Load libsurtr_b200.so; raises if it has not been built (see __graft_entry__.build)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(f"{LIB_PATH} is missing: the CUDA extension has not been built "
                           "(python -c 'import __graft_entry__ as g; g.build()'); there is no CPU fallback")
    lib = C.CDLL(LIB_PATH)
    vp, u32, i32 = C.c_void_p, C.c_uint32, C.c_int
    lib.surtr_ctx_create.argtypes = [i32, vp, C.POINTER(vp)]
    lib.surtr_ctx_destroy.argtypes = [vp]
    lib.surtr_ctx_destroy.restype = None
    lib.surtr_last_error.argtypes = [vp]
    lib.surtr_last_error.restype = C.c_char_p
    lib.surtr_version.restype = C.c_char_p
    lib.surtr_set_kdop_directions.argtypes = [vp, i32]
    lib.surtr_set_clip_build.argtypes = [vp, i32]
    lib.surtr_upload_pieces.argtypes = [vp, vp, vp, vp, vp, u32, vp, u32]
    lib.surtr_upload_cells.argtypes = [vp, vp, vp, vp, vp, u32, vp, u32]
    lib.surtr_upload_pieces3.argtypes = [vp, vp, vp, vp, vp, u32, vp, u32]
    lib.surtr_upload_cells3.argtypes = [vp, vp, vp, vp, vp, u32, vp, u32]
    lib.surtr_download_fragments_packed.argtypes = [vp, vp, vp, vp, vp]
    lib.surtr_download_fragments_packed_async.argtypes = [vp, vp, vp, vp, vp]
    lib.surtr_fragments_to_pieces.argtypes = [vp, vp, u32]
    lib.surtr_fragments_to_pieces_per_event.argtypes = [vp]
    lib.surtr_upload_pattern.argtypes = [vp, vp, vp, u32, vp, u32]
    lib.surtr_place_pattern.argtypes = [vp, vp, vp, u32]
    lib.surtr_fracture_event.argtypes = [vp]
    lib.surtr_event_counts.argtypes = [vp, C.POINTER(Counts)]
    lib.surtr_download_fragments.argtypes = [vp, vp, vp, vp, vp]
    lib.surtr_download_fragments_async.argtypes = [vp, vp, vp, vp, vp]
    lib.surtr_sync.argtypes = [vp]
    lib.surtr_transform_pieces.argtypes = [vp, vp, vp, u32]
    lib.surtr_download_pieces.argtypes = [vp, vp]
    lib.surtr_measure_fp32_peak.argtypes = [vp, C.POINTER(C.c_float)]
    lib.surtr_device_fragments.argtypes = [vp, C.POINTER(DeviceView)]
    lib.surtr_kdop_calc.argtypes = [vp, vp, u32, vp, u32, vp, vp, vp]
    lib.surtr_last_event_ms.argtypes = [vp, C.POINTER(C.c_float), C.POINTER(C.c_float)]
    lib.surtr_last_event_launches.argtypes = [vp]
    lib.surtr_set_profiling.argtypes = [vp, i32]
    lib.surtr_last_event_phases.argtypes = [vp, vp]
    lib.surtr_failed_pairs.argtypes = [vp, vp, C.c_uint64, C.POINTER(C.c_uint64)]
    u64 = C.c_uint64
    lib.surtr_input_blob_layout.argtypes = [u32, u64, u64, u32, u64, u64, u32, u32, C.POINTER(InLayout)]
    lib.surtr_upload_blob.argtypes = [vp, vp, u32, u64, u64, u32, u64, u64, u32, u32]
    lib.surtr_download_blob_async.argtypes = [vp, vp, u64, C.POINTER(OutLayout)]
    lib.surtr_kdop_calc_batch.argtypes = [vp, vp, vp, u32, vp, vp, vp, vp, vp]
    _lib = lib
    return lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _arr(a, dt):
    return None if a is None else np.ascontiguousarray(a, dt)


@dataclass
class Fragments:
    """Host copy of one event's fragments (flat layout of include/surtr_b200.h)."""
    rec: np.ndarray        # FRAGMENT_DTYPE [n]
    verts: np.ndarray      # float32 [NV,4]
    ring_off: np.ndarray   # uint32 [NV+1]
    ring: np.ndarray       # uint16 [NE]

    @property
    def n(self) -> int:
        return len(self.rec)

    @property
    def vert_off(self) -> np.ndarray:
        out = np.zeros(self.n + 1, np.uint32)
        if self.n:
            out[:-1] = self.rec["vert_off"]
            out[-1] = len(self.verts)
        return out

    def poly(self, i):
        v0 = int(self.rec["vert_off"][i])
        v1 = v0 + int(self.rec["n_verts"][i])
        rings = [self.ring[self.ring_off[v]:self.ring_off[v + 1]].tolist() for v in range(v0, v1)]
        return self.verts[v0:v1, :3].copy(), rings


class FractureContext:
    """One context per GPU / stream.  Mirrors the call sequence of Surtr::ApplyFracture (Surtr.cpp:2098-2149):
    upload_pieces + upload_cells -> fracture_event -> counts / download."""

    def __init__(self, device: int = 0, stream: int | None = None):
        self._lib = load_library()
        h = C.c_void_p()
        rc = self._lib.surtr_ctx_create(device, C.c_void_p(stream) if stream else None, C.byref(h))
        if rc != SURTR_OK:
            raise SurtrError(rc, self._lib.surtr_last_error(None).decode())
        self._h = h
        self._keep = []

    def close(self):
        if getattr(self, "_h", None):
            self._lib.surtr_ctx_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc):
        if rc != SURTR_OK:
            raise SurtrError(rc, self._lib.surtr_last_error(self._h).decode())

    def set_kdop_directions(self, k: int):
        self._ck(self._lib.surtr_set_kdop_directions(self._h, k))

    def set_clip_build(self, mode: int):
        """0 = by the size of the previous event (default), 1 = throughput build of K3's small tier, 2 = latency build."""
        self._ck(self._lib.surtr_set_clip_build(self._h, mode))

    def upload_pieces(self, verts4, vert_off, ring_off, ring, ev_piece_off=None):
        verts4 = _arr(verts4, np.float32)
        vert_off = _arr(vert_off, np.uint32)
        ring_off = _arr(ring_off, np.uint32)
        ring = _arr(ring, np.uint16)
        ev = _arr(ev_piece_off, np.uint32)
        self._ck(self._lib.surtr_upload_pieces(self._h, _p(verts4), _p(vert_off), _p(ring_off), _p(ring),
                                               len(vert_off) - 1, _p(ev), 0 if ev is None else len(ev) - 1))

    def upload_cells(self, planes4, plane_off, cell_verts4=None, cvert_off=None, ev_cell_off=None):
        planes4 = _arr(planes4, np.float32)
        plane_off = _arr(plane_off, np.uint32)
        cell_verts4 = _arr(cell_verts4, np.float32)
        cvert_off = _arr(cvert_off, np.uint32)
        ev = _arr(ev_cell_off, np.uint32)
        self._ck(self._lib.surtr_upload_cells(self._h, _p(planes4), _p(plane_off), _p(cell_verts4), _p(cvert_off),
                                              len(plane_off) - 1, _p(ev), 0 if ev is None else len(ev) - 1))

    def transform_pieces(self, matrices, piece_matrix=None):
        """Poly::Transform on the resident pieces (row-major 4x4 world matrices, see include/surtr_b200.h)."""
        matrices = _arr(np.asarray(matrices, np.float32).reshape(-1, 16), np.float32)
        piece_matrix = _arr(piece_matrix, np.uint32)
        self._ck(self._lib.surtr_transform_pieces(self._h, _p(matrices), _p(piece_matrix), len(matrices)))

    def download_pieces(self, n_verts):
        out = np.zeros((n_verts, 4), np.float32)
        self._ck(self._lib.surtr_download_pieces(self._h, _p(out)))
        return out

    def measure_fp32_peak(self) -> float:
        """FP32 FMA throughput of the device in TFLOP/s (measurement helper)."""
        t = C.c_float(0)
        self._ck(self._lib.surtr_measure_fp32_peak(self._h, C.byref(t)))
        return float(t.value)

    def upload_pattern(self, face_verts4, face_vert_off, cell_face_off):
        """Resident fracture pattern: the VertexVec of every face of every cell (see include/surtr_b200.h)."""
        face_verts4 = _arr(face_verts4, np.float32)
        face_vert_off = _arr(face_vert_off, np.uint32)
        cell_face_off = _arr(cell_face_off, np.uint32)
        self._ck(self._lib.surtr_upload_pattern(self._h, _p(face_verts4), _p(face_vert_off), len(face_vert_off) - 1,
                                                _p(cell_face_off), len(cell_face_off) - 1))

    def place_pattern(self, scale3, translate3):
        """Polygon3D::Scale + Translate on the device, one placement (= one event) per row of scale3 / translate3."""
        scale3 = _arr(np.atleast_2d(scale3), np.float32)
        translate3 = _arr(np.atleast_2d(translate3), np.float32)
        assert scale3.shape == translate3.shape and scale3.shape[1] == 3
        self._ck(self._lib.surtr_place_pattern(self._h, _p(scale3), _p(translate3), len(scale3)))

    def fragments_to_pieces(self, ev_piece_off=None):
        ev = _arr(ev_piece_off, np.uint32)
        self._ck(self._lib.surtr_fragments_to_pieces(self._h, _p(ev), 0 if ev is None else len(ev) - 1))

    def fragments_to_pieces_per_event(self):
        """The fragments of event e become the pieces of event e (event boundaries found on the device)."""
        self._ck(self._lib.surtr_fragments_to_pieces_per_event(self._h))

    def fracture_event(self):
        self._ck(self._lib.surtr_fracture_event(self._h))

    def counts(self, allow_failed: bool = False) -> Counts:
        """Counters of the last event.  Pairs that cannot be cut (malformed rings ...) raise SurtrError(4) unless
        allow_failed: then the counters (n_failed > 0) are returned and failed_pairs() lists them."""
        c = Counts()
        rc = self._lib.surtr_event_counts(self._h, C.byref(c))
        if rc == 4 and allow_failed and c.n_failed:
            return c
        self._ck(rc)
        return c

    def failed_pairs(self) -> np.ndarray:
        """(piece, cell) of every pair of the last event that could not be cut."""
        n = C.c_uint64(0)
        self._ck(self._lib.surtr_failed_pairs(self._h, None, 0, C.byref(n)))
        out = np.zeros((int(n.value), 2), np.uint32)
        if n.value:
            self._ck(self._lib.surtr_failed_pairs(self._h, _p(out), int(n.value), C.byref(n)))
        return out

    def download(self, geometry: bool = True, allow_failed: bool = False) -> Fragments:
        c = self.counts(allow_failed)
        rec = np.zeros(c.n_fragments, FRAGMENT_DTYPE)
        if geometry:
            verts = np.zeros((c.n_verts, 4), np.float32)
            ring_off = np.zeros(c.n_verts + 1, np.uint32)
            ring = np.zeros(c.n_ring, np.uint16)
        else:
            verts = ring_off = ring = None
        self._ck(self._lib.surtr_download_fragments(self._h, _p(rec), _p(verts), _p(ring_off), _p(ring)))
        if not geometry:
            verts = np.zeros((0, 4), np.float32)
            ring_off = np.zeros(1, np.uint32)
            ring = np.zeros(0, np.uint16)
        return Fragments(rec, verts, ring_off, ring)

    def download_into(self, rec, verts, ring_off, ring):
        """D2H into caller-owned (e.g. pinned) buffers; sizes from counts()."""
        self._ck(self._lib.surtr_download_fragments(self._h, C.c_void_p(rec), C.c_void_p(verts), C.c_void_p(ring_off),
                                                    C.c_void_p(ring)))

    def download_into_async(self, rec, verts, ring_off, ring):
        """Enqueue the D2H copies only; the buffers are complete after sync() (or the next counts())."""
        self._ck(self._lib.surtr_download_fragments_async(self._h, C.c_void_p(rec), C.c_void_p(verts), C.c_void_p(ring_off),
                                                          C.c_void_p(ring)))

    def sync(self):
        self._ck(self._lib.surtr_sync(self._h))

    def upload_pieces_ptr(self, verts4, vert_off, ring_off, ring, n_pieces, ev=None, n_events=0):
        self._ck(self._lib.surtr_upload_pieces(self._h, C.c_void_p(verts4), C.c_void_p(vert_off), C.c_void_p(ring_off),
                                               C.c_void_p(ring), n_pieces, C.c_void_p(ev) if ev else None, n_events))

    def upload_cells_ptr(self, planes4, plane_off, cell_verts4, cvert_off, n_cells, ev=None, n_events=0):
        self._ck(self._lib.surtr_upload_cells(self._h, C.c_void_p(planes4), C.c_void_p(plane_off),
                                              C.c_void_p(cell_verts4) if cell_verts4 else None,
                                              C.c_void_p(cvert_off) if cvert_off else None, n_cells,
                                              C.c_void_p(ev) if ev else None, n_events))

    # ---- PCIe wire format: float3 vertex streams up, float3 + one byte of ring length per vertex down ----
    def upload_pieces3_ptr(self, verts3, vert_off, ring_off, ring, n_pieces, ev=None, n_events=0):
        self._ck(self._lib.surtr_upload_pieces3(self._h, C.c_void_p(verts3), C.c_void_p(vert_off), C.c_void_p(ring_off),
                                                C.c_void_p(ring), n_pieces, C.c_void_p(ev) if ev else None, n_events))

    def upload_cells3_ptr(self, planes4, plane_off, cell_verts3, cvert_off, n_cells, ev=None, n_events=0):
        self._ck(self._lib.surtr_upload_cells3(self._h, C.c_void_p(planes4), C.c_void_p(plane_off),
                                               C.c_void_p(cell_verts3) if cell_verts3 else None,
                                               C.c_void_p(cvert_off) if cvert_off else None, n_cells,
                                               C.c_void_p(ev) if ev else None, n_events))

    def upload_pieces3(self, verts3, vert_off, ring_off, ring, ev_piece_off=None):
        verts3 = _arr(np.asarray(verts3, np.float32)[:, :3], np.float32)
        vert_off, ring_off, ring = _arr(vert_off, np.uint32), _arr(ring_off, np.uint32), _arr(ring, np.uint16)
        ev = _arr(ev_piece_off, np.uint32)
        self._ck(self._lib.surtr_upload_pieces3(self._h, _p(verts3), _p(vert_off), _p(ring_off), _p(ring),
                                                len(vert_off) - 1, _p(ev), 0 if ev is None else len(ev) - 1))

    def upload_cells3(self, planes4, plane_off, cell_verts3=None, cvert_off=None, ev_cell_off=None):
        planes4, plane_off = _arr(planes4, np.float32), _arr(plane_off, np.uint32)
        cell_verts3 = None if cell_verts3 is None else _arr(np.asarray(cell_verts3, np.float32)[:, :3], np.float32)
        cvert_off = _arr(cvert_off, np.uint32)
        ev = _arr(ev_cell_off, np.uint32)
        self._ck(self._lib.surtr_upload_cells3(self._h, _p(planes4), _p(plane_off), _p(cell_verts3), _p(cvert_off),
                                               len(plane_off) - 1, _p(ev), 0 if ev is None else len(ev) - 1))

    # ---- one-copy transfers: one blob per direction (include/surtr_b200.h) ----
    @staticmethod
    def input_blob_layout(n_pieces, n_pverts, n_pring, n_cells, n_planes, n_cverts, n_events, ring_entry_bytes) -> InLayout:
        L = InLayout()
        rc = load_library().surtr_input_blob_layout(n_pieces, n_pverts, n_pring, n_cells, n_planes, n_cverts, n_events, ring_entry_bytes, C.byref(L))
        if rc:
            raise SurtrError(rc, "surtr_input_blob_layout: ring_entry_bytes must be 1 or 2")
        return L

    @staticmethod
    def fill_input_blob(buf: np.ndarray, pieces, cells, ev_piece_off=None, ev_cell_off=None, bounded=True):
        """Lays the arrays of one batch out in `buf` (uint8, e.g. a view of pinned memory; None = only size it) in the
        compact wire format of surtr_upload_blob: float3 positions, one ring-length byte per vertex, the first ring entry
        of every piece, ring entries as bytes when no piece has more than 256 vertices.
        pieces / cells: objects with verts, vert_off, ring_off, ring / planes, plane_off, verts, vert_off.
        Returns (sizes tuple for upload_blob, total bytes)."""
        n_ev = 0 if ev_piece_off is None else len(ev_piece_off) - 1
        n_cv = len(cells.verts) if bounded else 0
        vo = np.asarray(pieces.vert_off, np.int64)
        rb = 1 if len(vo) < 2 or int(np.diff(vo).max()) <= 256 else 2
        sizes = (len(pieces.vert_off) - 1, len(pieces.verts), len(pieces.ring), len(cells.plane_off) - 1, len(cells.planes), n_cv, n_ev, rb)
        L = FractureContext.input_blob_layout(*sizes)
        if buf is None:
            return sizes, int(L.total)

        def put(off, a, dt):
            a = np.ascontiguousarray(a, dt).reshape(-1).view(np.uint8)
            buf[off:off + a.size] = a

        ro = np.asarray(pieces.ring_off, np.int64)
        put(L.verts3, np.asarray(pieces.verts)[:, :3], np.float32)
        put(L.vert_off, pieces.vert_off, np.uint32)
        put(L.ring_base, ro[vo], np.uint32)
        put(L.ring_len, np.diff(ro), np.uint8)
        put(L.ring, pieces.ring, np.uint8 if rb == 1 else np.uint16)
        put(L.planes4, cells.planes, np.float32)
        put(L.plane_off, cells.plane_off, np.uint32)
        if n_cv:
            put(L.cell_verts3, np.asarray(cells.verts)[:, :3], np.float32)
            put(L.cvert_off, cells.vert_off, np.uint32)
        if n_ev:
            put(L.ev_piece_off, ev_piece_off, np.uint32)
            put(L.ev_cell_off, ev_cell_off, np.uint32)
        return sizes, int(L.total)

    def upload_blob_ptr(self, blob_ptr, sizes):
        self._ck(self._lib.surtr_upload_blob(self._h, C.c_void_p(blob_ptr), *sizes))

    def download_blob_into_async(self, blob_ptr, capacity) -> OutLayout:
        L = OutLayout()
        self._ck(self._lib.surtr_download_blob_async(self._h, C.c_void_p(blob_ptr), capacity, C.byref(L)))
        return L

    @staticmethod
    def unpack_output_blob(buf: np.ndarray, L: OutLayout) -> Fragments:
        """Host-side view of an output blob as the usual arrays (float4 with w = 0, ring_off as prefix sum)."""
        nf, nv, nr = int(L.n_fragments), int(L.n_verts), int(L.n_ring)
        rec = np.frombuffer(buf[L.fragments:L.fragments + 64 * nf].tobytes(), dtype=FRAGMENT_DTYPE)
        v3 = np.frombuffer(buf[L.verts3:L.verts3 + 12 * nv].tobytes(), dtype=np.float32).reshape(nv, 3)
        rl = np.frombuffer(buf[L.ring_len:L.ring_len + nv].tobytes(), dtype=np.uint8)
        if int(L.ring_entry_bytes) == 1:
            ring = np.frombuffer(buf[L.ring:L.ring + nr].tobytes(), dtype=np.uint8).astype(np.uint16)
        else:
            ring = np.frombuffer(buf[L.ring:L.ring + 2 * nr].tobytes(), dtype=np.uint16)
        verts = np.zeros((nv, 4), np.float32)
        verts[:, :3] = v3
        ring_off = np.concatenate([[0], np.cumsum(rl, dtype=np.uint64)]).astype(np.uint32)
        return Fragments(rec, verts, ring_off, ring)

    def download_packed(self) -> Fragments:
        """Fragments through the packed wire format, unpacked to the usual arrays (float4 with w = 0, ring_off as the
        prefix sum of the ring lengths)."""
        c = self.counts()
        rec = np.zeros(c.n_fragments, FRAGMENT_DTYPE)
        v3 = np.zeros((c.n_verts, 3), np.float32)
        rl = np.zeros(c.n_verts, np.uint8)
        ring = np.zeros(c.n_ring, np.uint16)
        self._ck(self._lib.surtr_download_fragments_packed(self._h, _p(rec), _p(v3), _p(rl), _p(ring)))
        verts = np.zeros((c.n_verts, 4), np.float32)
        verts[:, :3] = v3
        ring_off = np.concatenate([[0], np.cumsum(rl, dtype=np.uint64)]).astype(np.uint32)
        return Fragments(rec, verts, ring_off, ring)

    def download_packed_into(self, rec, verts3, ring_len, ring):
        self._ck(self._lib.surtr_download_fragments_packed(self._h, C.c_void_p(rec), C.c_void_p(verts3), C.c_void_p(ring_len),
                                                           C.c_void_p(ring)))

    def download_packed_into_async(self, rec, verts3, ring_len, ring):
        self._ck(self._lib.surtr_download_fragments_packed_async(self._h, C.c_void_p(rec), C.c_void_p(verts3),
                                                                 C.c_void_p(ring_len), C.c_void_p(ring)))

    def device_view(self) -> DeviceView:
        v = DeviceView()
        self._ck(self._lib.surtr_device_fragments(self._h, C.byref(v)))
        return v

    def kdop_calc(self, verts4, normals):
        verts4 = _arr(verts4, np.float32)
        normals = _arr(normals, np.float32)
        k = len(normals)
        dist = np.zeros((k, 2), np.float32)
        arg = np.zeros((k, 2), np.int32)
        planes = np.zeros((k, 2, 4), np.float32)
        self._ck(self._lib.surtr_kdop_calc(self._h, _p(verts4), len(verts4), _p(normals), k, _p(dist), _p(arg), _p(planes)))
        return dist, arg, planes

    def kdop_calc_batch(self, verts4, vert_off, normals, normal_off):
        verts4, normals = _arr(verts4, np.float32), _arr(normals, np.float32)
        vert_off, normal_off = _arr(vert_off, np.uint32), _arr(normal_off, np.uint32)
        k = len(normals)
        dist = np.zeros((k, 2), np.float32)
        arg = np.zeros((k, 2), np.int32)
        planes = np.zeros((k, 2, 4), np.float32)
        self._ck(self._lib.surtr_kdop_calc_batch(self._h, _p(verts4), _p(vert_off), len(vert_off) - 1, _p(normals), _p(normal_off),
                                                 _p(dist), _p(arg), _p(planes)))
        return dist, arg, planes

    def last_event_ms(self):
        t, c = C.c_float(0), C.c_float(0)
        self._ck(self._lib.surtr_last_event_ms(self._h, C.byref(t), C.byref(c)))
        return t.value, c.value

    PHASES = ("k1_extents", "k2a_masks", "k2b_compact", "k3_clip_small", "k3_clip_large", "k4_scan", "k4_gather", "finish")

    def last_event_phases(self) -> dict:
        """Per-kernel milliseconds of the last event (needs set_profiling(True) before it was launched)."""
        ms = (C.c_float * 8)()
        self._ck(self._lib.surtr_last_event_phases(self._h, ms))
        return dict(zip(self.PHASES, (float(x) for x in ms)))

    def set_profiling(self, on: bool):
        self._ck(self._lib.surtr_set_profiling(self._h, int(on)))

    def last_event_launches(self) -> int:
        return self._lib.surtr_last_event_launches(self._h)
