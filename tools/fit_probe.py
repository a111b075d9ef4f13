"""pk_fit latency (Psi + factorisation with L^-T rows + finalisation) per n, gang kernel vs per-block-column kernels."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pykriging_b200._lib import Handle  # noqa: E402

h = Handle(0)
k = 6
for n in (150, 200, 300, 512, 1024, 2048, 4096):
    rng = np.random.default_rng(n)
    X = np.zeros((n, k))
    for d in range(k):
        X[:, d] = (rng.permutation(n) + 0.5) / n
    y = np.sum((X - 0.3) ** 2, axis=1) + rng.normal(0, 0.05, n)
    y = (y - y.min()) / (y.max() - y.min())
    h.set_data(X, y)
    th, p = rng.uniform(0.1, 10, k), rng.uniform(1.82, 2.0, k)
    res = {}
    for variant in (1, 0):
        h.set_cholesky_variant(variant)
        out4, info = h.fit(th, p, 0.01, 1)
        h.synchronize()
        t0 = time.perf_counter()
        for _ in range(3):
            out4, info = h.fit(th, p, 0.01, 1)
        h.synchronize()
        res[variant] = ((time.perf_counter() - t0) / 3 * 1e3, out4.copy(), info)
    a, b = res[1], res[0]
    print("n %5d  fit ms: per-block-column %.3f  gang %.3f   rel diff nll %.2e mu %.2e s2 %.2e info %d %d" % (
        n, a[0], b[0], abs(a[1][0] - b[1][0]) / abs(a[1][0]), abs(a[1][1] - b[1][1]) / abs(a[1][1]),
        abs(a[1][2] - b[1][2]) / abs(a[1][2]), a[2], b[2]), flush=True)
