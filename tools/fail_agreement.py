"""Diagnostic: how often does the CUDA path's 'pivot <= 0' outcome agree with LAPACK dpotrf on
numerically singular candidates (the knife-edge set of SURVEY App. A.4)?  Run on a GPU box."""
import json
import sys

import numpy as np
from scipy.linalg.lapack import dpotrf

sys.path.insert(0, ".")
from oracle import oracle_np as onp  # noqa: E402
from pykriging_b200._lib import Handle  # noqa: E402

h = Handle(0)
res = []
for (n, k, seed) in [(20, 2, 1), (40, 3, 2), (100, 2, 3), (300, 4, 4)]:
    X = onp.rlh(n, k, seed)
    y = onp.f_squared(X)
    Xn, yn, _, _ = onp.normalize_data(X, y)
    h.set_data(Xn, yn)
    rng = np.random.default_rng(seed)
    C = 600 if n <= 100 else 200
    th = 10 ** rng.uniform(-5, -0.5, (C, k))
    p = np.where(rng.uniform(size=(C, k)) < 0.5, 2.0, rng.uniform(1.5, 2, (C, k)))
    lap = np.array([dpotrf(onp.psi_fast(Xn, th[i], p[i]), lower=1)[1] for i in range(C)])
    lap_fail = lap > 0
    conds = np.array([np.linalg.cond(onp.psi_fast(Xn, th[i], p[i])) for i in range(C)])
    for ulps in (0, 1, 2, 4, 8, 16, 32, 64):
        h.set_pivot_tolerance(ulps * 2.220446049250313e-16)
        out = h.nll_batch(th, p)
        gpu_fail = out["info"] > 0
        mism = gpu_fail != lap_fail
        res.append({"n": n, "k": k, "C": C, "tol_ulps": ulps, "lapack_fail": int(lap_fail.sum()), "gpu_fail": int(gpu_fail.sum()),
                    "agree": int((~mism).sum()), "mismatch_min_cond": float(conds[mism].min()) if mism.any() else None,
                    "gpu_only_fail": int((gpu_fail & ~lap_fail).sum()), "lapack_only_fail": int((lap_fail & ~gpu_fail).sum()),
                    "gpu_only_fail_min_cond": float(conds[gpu_fail & ~lap_fail].min()) if (gpu_fail & ~lap_fail).any() else None})
        print(json.dumps(res[-1]))
