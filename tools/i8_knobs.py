"""Knobs of the INT8 Cholesky path measured on one GPU (C3 shape): L2 policy of the digit tiles (PK_I8_L2) and the
ragged-first pipeline (PK_RAGGED_FIRST) over population sizes.  One JSON line per setting; results must agree bit for
bit between the settings (a candidate's bits depend on nothing but the candidate)."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import oracle_np as onp  # noqa: E402
from pykriging_b200 import _lib  # noqa: E402

n, k = 1024, 10
X = onp.rlh(n, k, 3)
y = onp.f_stybtang_norm(X)
Xn, yn, _, _ = onp.normalize_data(X, y)


def cands_for(C):
    c = np.concatenate([onp.candidates_uniform(C // 2, k, 10), onp.candidates_loguniform(C - C // 2, k, 20, lo=-2)])
    return np.ascontiguousarray(c[:, :k]), np.ascontiguousarray(c[:, k:])


def measure(env, C, reps=5):
    for kk, v in env.items():
        os.environ[kk] = v
    h = _lib.Handle()
    for kk in env:
        del os.environ[kk]
    h.set_data(Xn, yn)
    th, p = cands_for(C)
    out = h.nll_batch(th, p)
    h.nll_batch(th, p)
    t0 = time.perf_counter()
    for _ in range(reps):
        h.nll_batch(th, p)
    wall = (time.perf_counter() - t0) / reps * 1e3
    h.set_profiling(True)
    h.phase_times()
    for _ in range(reps):
        h.nll_batch(th, p)
    ph = h.phase_times()
    h.set_profiling(False)
    h.close()
    return out, {"env": env, "C": C, "wall_ms": round(wall, 3), "psi_ms": round(ph["psi"] / reps, 3),
                 "chol_ms": round(ph["cholesky"] / reps, 3)}


base = {}
for C in (1024, 300):
    for l2 in ("0", "1", "2", "3"):
        out, line = measure({"PK_I8_L2": l2}, C)
        if l2 == "0":
            base[C] = out
        line["same_bits"] = all(np.array_equal(out[nm], base[C][nm], equal_nan=True) for nm in ("nll", "mu", "sigma2", "lndet", "info"))
        print(json.dumps(line), flush=True)
for C in (152, 300, 310, 333, 600, 1024):
    ref = None
    for rf in ("0", "1"):
        out, line = measure({"PK_RAGGED_FIRST": rf}, C)
        if ref is None:
            ref = out
        line["same_bits"] = all(np.array_equal(out[nm], ref[nm], equal_nan=True) for nm in ("nll", "mu", "sigma2", "lndet", "info"))
        print(json.dumps(line), flush=True)
