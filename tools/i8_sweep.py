"""Cholesky time of the INT8 kernel (variant 5) against task mode (variant 4) over population sizes and orders: the
evidence behind launch_cholesky's rule (INT8 from 64 candidates on).  One JSON line per shape."""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import oracle_np as onp  # noqa: E402
from pykriging_b200 import _lib  # noqa: E402

h = _lib.Handle()
for n, k in ((512, 6), (1024, 10), (2048, 6)):
    X = onp.rlh(n, k, 3)
    y = onp.f_stybtang_norm(X)
    Xn, yn, _, _ = onp.normalize_data(X, y)
    h.set_data(Xn, yn)
    for C in (8, 21, 32, 48, 64, 96, 128, 148, 200, 296, 300, 444, 1024):
        if n == 2048 and C > 300:
            continue
        cands = np.concatenate([onp.candidates_uniform(C // 2, k, 10), onp.candidates_loguniform(C - C // 2, k, 20, lo=-2)])
        th, p = np.ascontiguousarray(cands[:, :k]), np.ascontiguousarray(cands[:, k:])
        line = {"n": n, "k": k, "C": C}
        for v in (4, 5):
            h.set_cholesky_variant(v)
            h.nll_batch(th, p)
            h.set_profiling(True)
            h.phase_times()
            for _ in range(3):
                h.nll_batch(th, p)
            ph = h.phase_times()
            h.set_profiling(False)
            line["chol_ms_v%d" % v] = round(ph["cholesky"] / 3, 3)
        h.set_cholesky_variant(0)
        line["int8_over_task"] = round(line["chol_ms_v5"] / line["chol_ms_v4"], 3)
        print(json.dumps(line), flush=True)
