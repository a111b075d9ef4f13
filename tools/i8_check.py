"""INT8-tensor-core Cholesky (variant 5) against the DMMA kernel (variant 2): values and timing.

    python tools/i8_check.py [n k C] ...
"""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import oracle_np as onp  # noqa: E402
from pykriging_b200 import _lib  # noqa: E402


def run(h, n, k, C, seed=3, reps=3):
    X = onp.rlh(n, k, seed)
    y = onp.f_stybtang_norm(X)
    Xn, yn, _, _ = onp.normalize_data(X, y)
    h.set_data(Xn, yn)
    cands = np.concatenate([onp.candidates_uniform(C // 2, k, seed + 10), onp.candidates_loguniform(C - C // 2, k, seed + 20, lo=-3)])
    if C >= 8:
        cands[5, :k] = 1e-5
        cands[5, k:] = 2.0
    res = {}
    for v in (2, 5):
        h.set_cholesky_variant(v)
        out = h.nll_batch(cands[:, :k], cands[:, k:])
        h.synchronize()
        t0 = time.perf_counter()
        for _ in range(reps):
            out = h.nll_batch(cands[:, :k], cands[:, k:])
        h.synchronize()
        res[v] = (out, (time.perf_counter() - t0) / reps * 1e3)
    h.set_cholesky_variant(0)
    a, b = res[2][0], res[5][0]
    ok = (a["info"] == 0) & (b["info"] == 0)
    rel = lambda x, y: float(np.max(np.abs(x[ok] - y[ok]) / np.maximum(1.0, np.abs(x[ok])))) if ok.any() else 0.0
    line = {"n": n, "k": k, "C": C, "ms_v2": round(res[2][1], 3), "ms_v5": round(res[5][1], 3),
            "info_equal": bool(np.array_equal(a["info"] != 0, b["info"] != 0)), "fails_v2": int((a["info"] != 0).sum()),
            "fails_v5": int((b["info"] != 0).sum()),
            "nll_rel": rel(a["nll"], b["nll"]), "mu_rel": rel(a["mu"], b["mu"]), "s2_rel": rel(a["sigma2"], b["sigma2"]),
            "lndet_rel": rel(a["lndet"], b["lndet"])}
    # a few candidates against the oracle
    worst = 0.0
    for i in range(min(C, 6)):
        Psi = onp.psi_fast(Xn, cands[i, :k], cands[i, k:])
        nll, mu, s2, lndet, info = onp.neglikelihood_fast(Psi, yn)
        if info == 0 and b["info"][i] == 0:
            worst = max(worst, abs(b["nll"][i] - nll) / max(1.0, abs(nll)) / max(1.0, np.linalg.cond(Psi) * 2.2e-16 * 1e3))
    line["oracle_nll_err_over_tol"] = worst
    print(json.dumps(line), flush=True)


if __name__ == "__main__":
    h = _lib.Handle()
    shapes = [(200, 3, 16), (333, 7, 12), (640, 10, 40), (1024, 10, 128), (1024, 10, 1024)]
    if len(sys.argv) > 1:
        a = [int(x) for x in sys.argv[1:]]
        shapes = [tuple(a[i:i + 3]) for i in range(0, len(a), 3)]
    for n, k, C in shapes:
        run(h, n, k, C)
