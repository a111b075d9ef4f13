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


@pytest.mark.parametrize("nt", [1, 2, 3, 5, 16, 17, 64, 129])
@pytest.mark.parametrize("single_below,blocks_per_task", [(0, 1), (4, 2), (0, 2), (0, 30), (3, 5), (100, 7)])
def test_task_schedule_is_a_valid_dependency_order(lib, nt, single_below, blocks_per_task):
    """Task mode of the persistent Cholesky (csrc/pk_chol_cta.cu) hands out the tile tasks of a candidate in the order
    of this list, one ticket after the other; a task only ever waits for tasks of the same candidate that come EARLIER,
    which is what makes the device-side spin-waits deadlock free.  Re-checked here, independently of the library's own
    validator: every tile is produced exactly once, a diagonal tile follows the panel task that produces column j-1 of
    its row block, a panel task follows its diagonal tile and column j-1 of all its row blocks, and no task is longer
    than the 30 row blocks a warp can watch."""
    from pykriging_b200 import _lib

    tasks = _lib.task_schedule(nt, single_below, blocks_per_task)
    ddone = 0
    rdone = [0] * nt           # finished block columns per row block
    produced = set()
    for typ, j, rb, nb in tasks.tolist():
        if typ == 0:
            assert j == ddone, "diagonal tiles in order"
            if j > 0:
                assert rdone[j] == j, "row block %d is not complete up to column %d" % (j, j - 1)
            ddone = j + 1
            produced.add(("D", j))
        else:
            assert 1 <= nb <= 30 and j < rb and rb + nb <= nt
            assert ddone >= j + 1, "panel task of column %d before its diagonal tile" % j
            for b in range(rb, rb + nb):
                assert rdone[b] == j, "row block %d: column %d before column %d" % (b, j, rdone[b])
                rdone[b] = j + 1
                assert ("P", j, b) not in produced
                produced.add(("P", j, b))
    assert ddone == nt
    assert len(produced) == nt + nt * (nt - 1) // 2
    # look-ahead: the diagonal tile of column j+1 is handed out right after the FIRST panel task of column j
    for i, (typ, j, rb, nb) in enumerate(tasks.tolist()):
        if typ == 0 and j >= 1:
            prev = tasks[i - 1].tolist()
            assert prev[0] == 1 and prev[1] == j - 1 and prev[2] == j
