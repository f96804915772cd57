"""CPU tests of the drop-in boundary: the C-ABI library loads, exports every symbol include/surtr_b200.h declares,
and fails loudly (no CPU fallback) when there is no B200."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    text = open(os.path.join(ROOT, "include", "surtr_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(surtr_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from surtr_b200 import engine
    lib = engine.load_library()
    syms = header_symbols()
    assert len(syms) >= 15
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/surtr_b200.h but not exported"
    assert sorted(engine.EXPORTS) == syms
    nm = subprocess.run(["nm", "-D", "--defined-only", engine.LIB_PATH], capture_output=True, text=True).stdout
    for s in syms:
        assert re.search(rf"\bT {s}\b", nm), s
    assert b"sm_100a" in lib.surtr_version()


def test_fragment_record_layout_matches_header():
    from surtr_b200 import FRAGMENT_DTYPE
    assert FRAGMENT_DTYPE.itemsize == 64
    assert [FRAGMENT_DTYPE.fields[k][1] for k in ("cell", "piece", "vert_off", "n_verts", "n_faces", "volume",
                                                  "centroid", "inertia", "n_ring")] == [0, 4, 8, 12, 14, 16, 24, 36, 60]


def test_no_device_fails_loudly():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from surtr_b200 import FractureContext, SurtrError
    with pytest.raises(SurtrError) as e:
        FractureContext(0)
    assert e.value.code == 3 and "no CPU fallback" in str(e.value)


def test_plain_c_caller_compiles_links_and_fails_loudly_without_a_gpu(tmp_path):
    """The boundary is a C ABI: a C99 translation unit (what a cgo / JNI / ctypes stub sees) must compile against
    include/surtr_b200.h with -pedantic, link libsurtr_b200.so, and -- on a machine without a GPU -- get
    SURTR_ERR_NO_DEVICE and an error text from surtr_ctx_create instead of any CPU fallback."""
    src = tmp_path / "caller.c"
    src.write_text('''
#include <stdio.h>
#include <string.h>
#include "surtr_b200.h"
int main(void)
{
    surtr_ctx* ctx = NULL;
    surtr_counts counts;
    surtr_fragment frag;
    int rc;
    memset(&counts, 0, sizeof counts);
    memset(&frag, 0, sizeof frag);
    if (sizeof(surtr_fragment) != 64) return 10;
    if (!strstr(surtr_version(), "sm_100a")) return 11;
    rc = surtr_ctx_create(0, NULL, &ctx);
    if (rc == SURTR_OK) { surtr_ctx_destroy(ctx); puts("device"); return 0; }
    if (rc != SURTR_ERR_NO_DEVICE || ctx != NULL) return 12;
    if (!strstr(surtr_last_error(NULL), "no CPU fallback")) return 13;
    /* every entry point refuses a NULL context instead of touching memory */
    if (surtr_fracture_event(NULL) == SURTR_OK || surtr_upload_pieces3(NULL, NULL, NULL, NULL, NULL, 0, NULL, 0) == SURTR_OK ||
        surtr_download_fragments_packed(NULL, NULL, NULL, NULL, NULL) == SURTR_OK || surtr_event_counts(NULL, &counts) == SURTR_OK)
        return 14;
    puts("no device");
    return 0;
}
''')
    exe = tmp_path / "caller"
    libdir = os.path.join(ROOT, "surtr_b200")
    cc = subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-I", os.path.join(ROOT, "include"), str(src),
                         "-o", str(exe), "-L", libdir, "-lsurtr_b200", f"-Wl,-rpath,{libdir}"], capture_output=True, text=True)
    assert cc.returncode == 0, cc.stderr
    run = subprocess.run([str(exe)], capture_output=True, text=True)
    assert run.returncode == 0, (run.returncode, run.stdout, run.stderr)
    import torch
    assert run.stdout.strip() == ("device" if torch.cuda.is_available() else "no device")


def test_product_never_imports_oracle():
    """The oracle is test infrastructure: nothing under surtr_b200/ or include/ may reference it."""
    bad = []
    for base in ("surtr_b200", "include"):
        for dp, _, files in os.walk(os.path.join(ROOT, base)):
            for f in files:
                if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp", ".cpp")):
                    txt = open(os.path.join(dp, f), errors="ignore").read()
                    if re.search(r"(from|import)\s+oracle|oracle/|surtr_oracle|libsurtr_ref", txt):
                        if "Nothing here imports oracle" in txt and txt.count("oracle") == 1:
                            continue
                        bad.append(os.path.join(dp, f))
    assert not bad, bad


@pytest.mark.gpu
def test_plain_c_caller_runs_an_event_and_matches_the_reference_fixture(tmp_path):
    """A C99 program -- nothing but include/surtr_b200.h and libc -- runs the committed reference event (unit cube x 64
    Voronoi cells, tests/golden/cube_x64.npz, written out as raw arrays) through upload -> fracture_event -> counts ->
    download, and again through the one-blob calls, and compares every output array with the reference build's bytes."""
    import sys
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from test_oracle_port import load_polyset
    d = np.load(os.path.join(ROOT, "tests", "golden", "cube_x64.npz"))
    cells, pieces, want = load_polyset(d, "cells_"), load_polyset(d, "pieces_"), load_polyset(d, "frag_")
    raw = tmp_path / "event.bin"
    arrays = [np.ascontiguousarray(pieces.verts, np.float32), np.ascontiguousarray(pieces.vert_off, np.uint32),
              np.ascontiguousarray(pieces.ring_off, np.uint32), np.ascontiguousarray(pieces.ring, np.uint16),
              np.ascontiguousarray(cells.planes, np.float32), np.ascontiguousarray(cells.plane_off, np.uint32),
              np.ascontiguousarray(cells.verts, np.float32), np.ascontiguousarray(cells.vert_off, np.uint32),
              # expected
              np.ascontiguousarray(want.cell, np.uint32), np.ascontiguousarray(want.piece, np.uint32),
              np.ascontiguousarray(want.nverts, np.uint32), np.ascontiguousarray(want.nfaces, np.uint32),
              np.ascontiguousarray(want.volume, np.float64), np.ascontiguousarray(want.centroid, np.float32),
              np.ascontiguousarray(want.verts, np.float32), np.ascontiguousarray(want.ring_off, np.uint32),
              np.ascontiguousarray(want.ring, np.uint16)]
    with open(raw, "wb") as f:
        f.write(np.array([a.nbytes for a in arrays], np.uint64).tobytes())
        for a in arrays:
            f.write(a.tobytes())
    src = tmp_path / "event.c"
    src.write_text(r'''
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "surtr_b200.h"
#define NA 17
static void* arr[NA];
static unsigned long long nbytes[NA];
#define CK(x) do { int rc_ = (x); if (rc_ != SURTR_OK) { printf("%s -> %d: %s\n", #x, rc_, surtr_last_error(ctx)); return 20; } } while (0)
int main(int argc, char** argv)
{
    FILE* f = fopen(argv[1], "rb");
    int i;
    surtr_ctx* ctx = NULL;
    surtr_counts c;
    surtr_fragment* rec;
    float* verts;
    unsigned* ring_off;
    unsigned short* ring;
    unsigned n_pieces, n_cells, n;
    (void)argc;
    if (!f || fread(nbytes, 8, NA, f) != NA) return 1;
    for (i = 0; i < NA; i++) { arr[i] = malloc(nbytes[i] + 16); if (fread(arr[i], 1, nbytes[i], f) != nbytes[i]) return 2; }
    fclose(f);
    n_pieces = (unsigned)(nbytes[1] / 4 - 1);
    n_cells = (unsigned)(nbytes[5] / 4 - 1);
    CK(surtr_ctx_create(0, NULL, &ctx));
    CK(surtr_upload_pieces(ctx, arr[0], arr[1], arr[2], arr[3], n_pieces, NULL, 0));
    CK(surtr_upload_cells(ctx, arr[4], arr[5], arr[6], arr[7], n_cells, NULL, 0));
    CK(surtr_fracture_event(ctx));
    CK(surtr_event_counts(ctx, &c));
    n = (unsigned)c.n_fragments;
    if (n != nbytes[8] / 4 || c.n_verts * 16 != nbytes[14] || c.n_ring * 2 != nbytes[16]) { puts("counts differ"); return 3; }
    rec = malloc(sizeof(surtr_fragment) * n);
    verts = malloc(16 * c.n_verts);
    ring_off = malloc(4 * (c.n_verts + 1));
    ring = malloc(2 * c.n_ring + 2);
    CK(surtr_download_fragments(ctx, rec, verts, ring_off, ring));
    for (i = 0; i < (int)n; i++)
    {
        if (rec[i].cell != ((unsigned*)arr[8])[i] || rec[i].piece != ((unsigned*)arr[9])[i] || rec[i].n_verts != ((unsigned*)arr[10])[i] ||
            rec[i].n_faces != ((unsigned*)arr[11])[i] || memcmp(&rec[i].volume, (double*)arr[12] + i, 8) ||
            memcmp(rec[i].centroid, (float*)arr[13] + 3 * i, 12)) { printf("fragment %d differs\n", i); return 4; }
    }
    if (memcmp(verts, arr[14], nbytes[14]) || memcmp(ring_off, arr[15], nbytes[15]) || memcmp(ring, arr[16], nbytes[16])) { puts("geometry differs"); return 5; }
    /* the one-blob calls: same fragments */
    {
        surtr_in_layout L;
        surtr_out_layout O;
        unsigned char *in, *out;
        unsigned long long npv = nbytes[0] / 16, npr = nbytes[3] / 2, npl = nbytes[4] / 16, ncv = nbytes[6] / 16, k;
        CK(surtr_input_blob_layout(n_pieces, npv, npr, n_cells, npl, ncv, 0, 1, &L));   /* pieces of <= 256 vertices: ring entries as bytes */
        in = calloc(1, L.total);
        for (k = 0; k < npv; k++) memcpy(in + L.verts3 + 12 * k, (float*)arr[0] + 4 * k, 12);
        memcpy(in + L.vert_off, arr[1], nbytes[1]);
        for (k = 0; k <= n_pieces; k++) ((unsigned*)(in + L.ring_base))[k] = ((unsigned*)arr[2])[((unsigned*)arr[1])[k]];
        for (k = 0; k < npv; k++) in[L.ring_len + k] = (unsigned char)(((unsigned*)arr[2])[k + 1] - ((unsigned*)arr[2])[k]);
        for (k = 0; k < npr; k++) in[L.ring + k] = (unsigned char)((unsigned short*)arr[3])[k];
        memcpy(in + L.planes4, arr[4], nbytes[4]); memcpy(in + L.plane_off, arr[5], nbytes[5]);
        for (k = 0; k < ncv; k++) memcpy(in + L.cell_verts3 + 12 * k, (float*)arr[6] + 4 * k, 12);
        memcpy(in + L.cvert_off, arr[7], nbytes[7]);
        CK(surtr_upload_blob(ctx, in, n_pieces, npv, npr, n_cells, npl, ncv, 0, 1));
        CK(surtr_fracture_event(ctx));
        out = malloc(64 * (size_t)n + 13 * c.n_verts + 2 * c.n_ring + 1024);
        CK(surtr_download_blob_async(ctx, out, 64 * (unsigned long long)n + 13 * c.n_verts + 2 * c.n_ring + 1024, &O));
        CK(surtr_sync(ctx));
        if (O.n_fragments != n || O.ring_entry_bytes != 1 || memcmp(out + O.fragments, rec, sizeof(surtr_fragment) * n)) { puts("blob differs"); return 6; }
        for (k = 0; k < c.n_ring; k++) if (out[O.ring + k] != ring[k]) { puts("blob rings differ"); return 6; }
        for (k = 0; k < c.n_verts; k++)
            if (memcmp(out + O.verts3 + 12 * k, verts + 4 * k, 12) || out[O.ring_len + k] != ring_off[k + 1] - ring_off[k]) { puts("blob geometry differs"); return 7; }
    }
    surtr_ctx_destroy(ctx);
    printf("ok %u fragments\n", n);
    return 0;
}
''')
    exe = tmp_path / "event"
    libdir = os.path.join(ROOT, "surtr_b200")
    cc = subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-I", os.path.join(ROOT, "include"), str(src),
                         "-o", str(exe), "-L", libdir, "-lsurtr_b200", f"-Wl,-rpath,{libdir}"], capture_output=True, text=True)
    assert cc.returncode == 0, cc.stderr
    run = subprocess.run([str(exe), str(raw)], capture_output=True, text=True)
    assert run.returncode == 0 and run.stdout.startswith("ok 64"), (run.returncode, run.stdout, run.stderr)


def test_input_blob_layout_and_packing_round_trip():
    """The compact input blob of surtr_upload_blob (host side only, no GPU): sections are 256-byte aligned, disjoint and
    in the documented order for both ring entry widths; fill_input_blob writes what the header says -- float3 positions,
    one ring-length byte per vertex, the first ring entry of every piece, one- or two-byte ring entries -- and the arrays
    can be read back from it."""
    import common
    from surtr_b200 import FractureContext
    pieces, cells = common.voronoi(1234, 30), common.voronoi(46354, 6)
    for rb in (1, 2):
        L = FractureContext.input_blob_layout(pieces.n, len(pieces.verts), len(pieces.ring), cells.n, len(cells.planes), len(cells.verts), 0, rb)
        offs = [L.verts3, L.vert_off, L.ring_base, L.ring_len, L.ring, L.planes4, L.plane_off, L.cell_verts3, L.cvert_off, L.ev_piece_off, L.ev_cell_off, L.total]
        sizes = [12 * len(pieces.verts), 4 * (pieces.n + 1), 4 * (pieces.n + 1), len(pieces.verts), rb * len(pieces.ring), 16 * len(cells.planes),
                 4 * (cells.n + 1), 12 * len(cells.verts), 4 * (cells.n + 1), 4, 4]
        assert all(o % 256 == 0 for o in offs)
        assert all(offs[i] + sizes[i] <= offs[i + 1] for i in range(len(sizes)))
    with pytest.raises(Exception):
        FractureContext.input_blob_layout(1, 1, 1, 1, 1, 1, 0, 3)
    sizes, total = FractureContext.fill_input_blob(None, pieces, cells)
    assert sizes[-1] == 1                                     # Voronoi cells: far fewer than 256 vertices each
    buf = np.zeros(total, np.uint8)
    FractureContext.fill_input_blob(buf, pieces, cells)
    L = FractureContext.input_blob_layout(*sizes)
    nv, ne = len(pieces.verts), len(pieces.ring)
    assert np.array_equal(buf[L.verts3:L.verts3 + 12 * nv].view(np.float32).reshape(nv, 3), pieces.verts[:, :3])
    ring_len = buf[L.ring_len:L.ring_len + nv]
    ring_base = buf[L.ring_base:L.ring_base + 4 * (pieces.n + 1)].view(np.uint32)
    assert np.array_equal(ring_len, np.diff(pieces.ring_off)) and ring_base[-1] == ne
    # ring offsets as expand_blob_kernel rebuilds them: the piece's first entry + the running sum of its lengths
    rebuilt = np.zeros(nv + 1, np.uint32)
    for p in range(pieces.n):
        v0, v1 = int(pieces.vert_off[p]), int(pieces.vert_off[p + 1])
        rebuilt[v0:v1] = ring_base[p] + np.concatenate([[0], np.cumsum(ring_len[v0:v1 - 1], dtype=np.uint32)]) if v1 > v0 else 0
    rebuilt[nv] = ring_base[-1]
    assert np.array_equal(rebuilt, pieces.ring_off)
    assert np.array_equal(buf[L.ring:L.ring + ne], pieces.ring.astype(np.uint8))
