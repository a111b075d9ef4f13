"""Gang-mode Cholesky probe: time and info for a few (n, C) with variant 3."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pykriging_b200._lib import Handle  # noqa: E402

h = Handle(0)
dev = torch.device("cuda", 0)
k = 6
n = int(sys.argv[1])
rng = np.random.default_rng(n)
X = np.zeros((n, k))
for d in range(k):
    X[:, d] = (rng.permutation(n) + 0.5) / n
y = np.sum((X - 0.3) ** 2, axis=1) + rng.normal(0, 0.05, n)
y = (y - y.min()) / (y.max() - y.min())
h.set_data(X, y)
for C in [int(c) for c in sys.argv[2].split(",")]:
    th = torch.from_numpy(rng.uniform(1e-4, 100, (C, k))).to(dev)
    p = torch.from_numpy(rng.uniform(1.81231, 2.0, (C, k))).to(dev)
    lam = torch.from_numpy(rng.uniform(0, 1, C)).to(dev)
    out = torch.empty(4, C, dtype=torch.float64, device=dev)
    info = torch.empty(C, dtype=torch.int32, device=dev)
    h.set_cholesky_variant(3)
    h.nll_batch_raw(C, th, p, lam, 1, out[0], out[1], out[2], out[3], info)  # warm-up (allocations, attributes)
    h.synchronize()
    h.set_profiling(True)
    h.phase_times()
    h.nll_batch_raw(C, th, p, lam, 1, out[0], out[1], out[2], out[3], info)
    h.synchronize()
    print(n, C, "chol ms", round(h.phase_times()["cholesky"], 3), "info", info.cpu().numpy().tolist()[:8], flush=True)
