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
