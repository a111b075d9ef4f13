"""pk_nll_batch on one shape, a few repetitions (target of `ncu -k regex:chol_task` / `regex:psi_pop` captures).
    python tools/task_probe.py N K C VARIANT [reps]"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pykriging_b200._lib import Handle  # noqa: E402

n, k, C, variant = (int(v) for v in sys.argv[1:5])
reps = int(sys.argv[5]) if len(sys.argv) > 5 else 4
h = Handle(0)
dev = torch.device("cuda", 0)
rng = np.random.default_rng(n)
X = np.zeros((n, k))
for d in range(k):
    X[:, d] = (rng.permutation(n) + 0.5) / n
y = np.sum((X - 0.3) ** 2, axis=1)
y = (y - y.min()) / (y.max() - y.min())
h.set_data(X, y)
th = torch.from_numpy(rng.uniform(1e-4, 100, (C, k))).to(dev)
p = torch.from_numpy(rng.uniform(1.0, 2.0, (C, k))).to(dev)
out = torch.empty(4, C, dtype=torch.float64, device=dev)
info = torch.empty(C, dtype=torch.int32, device=dev)
h.set_cholesky_variant(variant)
h.set_profiling(True)
for rep in range(reps):
    if rep == 1:
        h.synchronize()
        h.phase_times()
    h.nll_batch_raw(C, th, p, None, 0, out[0], out[1], out[2], out[3], info)
h.synchronize()
ph = h.phase_times()
print({kk: round(v / (reps - 1), 4) for kk, v in ph.items()}, "fails", int((info != 0).sum().item()))
