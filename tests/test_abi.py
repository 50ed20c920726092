"""The C-ABI library loads and exports every symbol include/frankenz_b200.h declares (no GPU needed)."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from frankenz_b200 import build, _lib
    build.build_library()
    return _lib.load()


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "frankenz_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(fzb_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_exported(lib):
    from frankenz_b200 import _lib
    names = declared_symbols()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), "libfzb200.so does not export %s" % n
        assert n in _lib.SIGNATURES, "ctypes binding misses %s" % n
    assert set(_lib.SIGNATURES) == set(names)


def test_struct_layouts(lib):
    from frankenz_b200 import _lib
    assert C.sizeof(_lib.FzbConfig) == 56
    assert C.sizeof(_lib.FzbFitOut) == 7 * 8
    assert C.sizeof(_lib.FzbStats) == 18 * 8
    assert lib.fzb_version() >= 100


def test_no_cpu_fallback(lib):
    """Without a CUDA device the library refuses to create a handle instead of computing on the CPU."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    h = C.c_void_p()
    rc = lib.fzb_create(0, C.byref(h))
    assert rc != 0 and b"no CPU fallback" in lib.fzb_last_error()
    import numpy as np
    import frankenz_b200
    bf = frankenz_b200.BruteForce(np.ones((4, 5)), np.ones((4, 5)), np.ones((4, 5)))
    with pytest.raises(RuntimeError):
        bf.fit(np.ones((2, 5)), np.ones((2, 5)), np.ones((2, 5)), verbose=False)


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "frankenz_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in txt.replace("the oracle", ""), "%s references oracle/" % f
    # tools/ are measurement and debugging aids of the product: they may not execute the oracle either (scripts that time
    # the oracle beside the GPU live under tests/scripts/)
    for f in os.listdir(os.path.join(ROOT, "tools")):
        if f.endswith((".py", ".sh")):
            txt = open(os.path.join(ROOT, "tools", f)).read()
            assert "import fz_oracle" not in txt and "from oracle" not in txt, "tools/%s imports the oracle" % f
