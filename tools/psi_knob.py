"""A knob of the population Psi kernel (default: PK_PSI_CONST, (theta, p) through constant memory) on / off over a few
shapes: phase times and bits (results must agree bit for bit).

    python tools/psi_knob.py [ENV_NAME]
"""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import oracle_np as onp  # noqa: E402
from pykriging_b200 import _lib  # noqa: E402

KNOB = sys.argv[1] if len(sys.argv) > 1 else "PK_PSI_CONST"


def run(env, n, k, C, reps=5):
    os.environ.update(env)
    h = _lib.Handle()
    for kk in env:
        del os.environ[kk]
    X = onp.rlh(n, k, 3)
    y = onp.f_stybtang_norm(X) if k > 3 else onp.f_squared(X)
    Xn, yn, _, _ = onp.normalize_data(X, y)
    h.set_data(Xn, yn)
    c = np.concatenate([onp.candidates_uniform(C // 2, k, 10), onp.candidates_loguniform(C - C // 2, k, 20, lo=-2)])
    th, p = np.ascontiguousarray(c[:, :k]), np.ascontiguousarray(c[:, k:])
    if C > 8:
        p[3, :] = 2.0   # exact-exponent path
        p[5, 0] = 1.0
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
    return out, {"env": env, "n": n, "k": k, "C": C, "wall_ms": round(wall, 4), "psi_ms": round(ph["psi"] / reps, 4),
                 "chol_ms": round(ph["cholesky"] / reps, 4)}


SHAPES = [tuple(int(v) for v in a.split(",")) for a in sys.argv[2:]] or [(1024, 10, 1024), (1024, 10, 128), (1024, 10, 300), (770, 7, 200), (1024, 8, 300), (200, 9, 130), (4096, 6, 16)]
for n, k, C in SHAPES:
    base = None
    for env in (({KNOB: "4"}, {KNOB: "1"}) if KNOB == "PK_PSI_PPT_SMALL" else ({KNOB: "99"}, {KNOB: "1"}) if KNOB == "PK_PSI_CONST_MINK" else ({KNOB: "0"}, {})):
        out, line = run(env, n, k, C)
        if base is None:
            base = out
        line["same_bits"] = all(np.array_equal(out[nm], base[nm], equal_nan=True) for nm in ("nll", "mu", "sigma2", "lndet", "info"))
        line["failed"] = int((out["info"] != 0).sum())
        print(json.dumps(line), flush=True)
