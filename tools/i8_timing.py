"""Per-phase cycle counts of the INT8 Cholesky (PK_CTA_TIMING build of chol_i8_kernel) on the C3 shape: the FP64 warps'
phases (warp 3) and warp 0's view of the 64 x 64 tile factorisation.  The counts go to stderr.

    python tools/i8_timing.py [C ...]
"""
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
for C in [int(a) for a in sys.argv[1:]] or [1024]:
    c = np.concatenate([onp.candidates_uniform(C // 2, k, 10), onp.candidates_loguniform(C - C // 2, k, 20, lo=-2)])
    th, p = np.ascontiguousarray(c[:, :k]), np.ascontiguousarray(c[:, k:])
    for env in ({}, {"PK_CTA_TIMING": "1"}):
        os.environ.update(env)
        h = _lib.Handle()
        for kk in env:
            del os.environ[kk]
        h.set_data(Xn, yn)
        h.set_cholesky_variant(5)
        h.nll_batch(th, p)
        t0 = time.perf_counter()
        out = h.nll_batch(th, p)
        wall = (time.perf_counter() - t0) * 1e3
        h.set_profiling(True)
        h.phase_times()
        for _ in range(3):
            h.nll_batch(th, p)
        ph = h.phase_times()
        h.set_profiling(False)
        sig = float(np.nansum(out["nll"]))
        print("C", C, env, "ms per call %.3f" % wall, "psi %.3f chol %.3f" % (ph["psi"] / 3, ph["cholesky"] / 3), "failed",
              int((out["info"] != 0).sum()), "sum nll %r" % sig, flush=True)
        h.close()
