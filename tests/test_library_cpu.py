"""CPU-side checks of the boundary: the C-ABI library loads, exports every symbol the header
declares, and refuses to run without a GPU (no silent fallback).  No compute calls here."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    text = open(os.path.join(ROOT, "include", "pykriging_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(pk_[a-zA-Z0-9_]+)\s*\(", text)))


def test_header_declares_the_documented_entry_points():
    syms = _header_symbols()
    for must in ("pk_create", "pk_destroy", "pk_set_data", "pk_nll_batch", "pk_fit", "pk_predict_batch",
                 "pk_get_psi", "pk_get_U", "pk_psi_batch", "pk_last_error"):
        assert must in syms


def test_library_exports_every_header_symbol(lib):
    from pykriging_b200 import _lib

    syms = _header_symbols()
    assert sorted(_lib.EXPORTS) == syms, "keep _lib.EXPORTS in sync with include/pykriging_b200.h"
    for s in syms:
        assert hasattr(lib, s), "libpykriging_b200.so does not export %s" % s
    assert lib.pk_version() == 100


def test_header_cites_reference_interfaces():
    text = open(os.path.join(ROOT, "include", "pykriging_b200.h")).read()
    for cite in ("matrixops.py:18-22", "matrixops.py:24-42", "matrixops.py:45-66", "matrixops.py:68-77",
                 "matrixops.py:79-95", "krige.py:429-452", "regressionkrige.py:428-452", "krige.py:201-234"):
        assert cite in text


def test_no_cpu_fallback_without_gpu(lib):
    """Without a CUDA device pk_create must fail loudly (on a GPU box it succeeds)."""
    h = ctypes.c_void_p()
    rc = lib.pk_create(ctypes.byref(h), 0)
    try:
        import torch

        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = rc == 0
    if has_gpu:
        assert rc == 0
        lib.pk_destroy(h)
    else:
        assert rc != 0 and not h.value
        assert b"no CPU fallback" in lib.pk_last_error(None)
        from pykriging_b200._lib import Handle, PkError

        with pytest.raises(PkError):
            Handle(0)


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "pykriging_b200")
    for fn in os.listdir(pkg):
        if fn.endswith(".py"):
            src = open(os.path.join(pkg, fn)).read()
            assert "oracle" not in re.sub(r'""".*?"""', "", src, flags=re.S).replace("oracle/", ""), fn
