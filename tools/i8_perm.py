"""INT8 Cholesky with the first and with the permuted digit-plane layout (PK_I8_PERM: shuffle-free triangular solve): time per
population and bit equality (the two layouts perform the same arithmetic on every element)."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import oracle_np as onp  # noqa: E402
from pykriging_b200 import _lib  # noqa: E402

shapes = [(1024, 10, 1024), (1024, 10, 128), (2048, 6, 148), (770, 3, 300), (4096, 6, 148)]
if len(sys.argv) > 1:
    shapes = shapes[: int(sys.argv[1])]
for n, k, C in shapes:
    X = onp.rlh(n, k, 3)
    y = onp.f_stybtang_norm(X)
    Xn, yn, _, _ = onp.normalize_data(X, y)
    c = np.concatenate([onp.candidates_uniform(C // 2, k, 10), onp.candidates_loguniform(C - C // 2, k, 20, lo=-2)])
    th, p = np.ascontiguousarray(c[:, :k]), np.ascontiguousarray(c[:, k:])
    th[3] = 1e-7          # nearly singular: fails, or succeeds with a wild likelihood -- both layouts must agree
    p[3] = 2.0
    ref = None
    for perm in ("0", "1"):
        os.environ["PK_I8_PERM"] = perm
        h = _lib.Handle()
        del os.environ["PK_I8_PERM"]
        h.set_data(Xn, yn)
        h.set_cholesky_variant(5)
        out = h.nll_batch(th, p)
        t0 = time.perf_counter()
        for _ in range(4):
            h.nll_batch(th, p)
        wall = (time.perf_counter() - t0) / 4 * 1e3
        h.set_profiling(True)
        h.phase_times()
        for _ in range(4):
            h.nll_batch(th, p)
        ph = h.phase_times()
        h.set_profiling(False)
        h.close()
        if ref is None:
            ref = out
        same = all(np.array_equal(out[nm], ref[nm], equal_nan=True) for nm in ("nll", "mu", "sigma2", "lndet", "info"))
        print(json.dumps({"n": n, "k": k, "C": C, "perm": int(perm), "wall_ms": round(wall, 3), "chol_ms": round(ph["cholesky"] / 4, 3),
                          "failed": int((out["info"] != 0).sum()), "same_bits": same}), flush=True)
