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
