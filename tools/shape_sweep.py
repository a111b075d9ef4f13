"""Throughput of pk_nll_batch over problem shapes (evals/s and Cholesky TFLOP/s), device-resident inputs."""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pykriging_b200._lib import Handle  # noqa: E402

h = Handle(0)
dev = torch.device("cuda", 0)
shapes = [(256, 4, 2048, 0), (512, 6, 2048, 0), (1024, 10, 1024, 0), (2048, 8, 256, 0), (4096, 6, 16, 1), (4096, 6, 64, 1),
          (4096, 6, 160, 1)]
for n, k, C, mode in shapes:
    rng = np.random.default_rng(n)
    X = np.zeros((n, k))
    for d in range(k):
        X[:, d] = (rng.permutation(n) + 0.5) / n
    y = np.sum((X - 0.3) ** 2, axis=1) + (rng.normal(0, 0.05, n) if mode else 0.0)
    y = (y - y.min()) / (y.max() - y.min())
    h.set_data(X, y)
    th = torch.from_numpy(rng.uniform(1e-4, 100, (C, k))).to(dev)
    p = torch.from_numpy(rng.uniform(1.81231 if mode else 1.0, 2.0, (C, k))).to(dev)
    lam = torch.from_numpy(rng.uniform(0, 1, C)).to(dev) if mode else None
    out = torch.empty(4, C, dtype=torch.float64, device=dev)
    info = torch.empty(C, dtype=torch.int32, device=dev)
    for variant in ((0,) if n < 4096 else (1, 2)):
        h.set_cholesky_variant(variant)
        h.set_profiling(True)
        for rep in range(3):
            if rep == 1:
                h.synchronize()
                h.phase_times()
                t0 = time.perf_counter()
            h.nll_batch_raw(C, th, p, lam, mode, out[0], out[1], out[2], out[3], info)
        h.synchronize()
        dt = (time.perf_counter() - t0) / 2
        ph = h.phase_times()
        print(json.dumps({"n": n, "k": k, "C": C, "mode": mode, "variant": variant, "evals_per_s": C / dt,
                          "ms": dt * 1e3, "psi_ms": ph["psi"] / 2, "chol_ms": ph["cholesky"] / 2,
                          "chol_tflops": C * n ** 3 / 3 / (ph["cholesky"] / 2 * 1e-3) * 1e-12,
                          "fails": int((info != 0).sum().item())}), flush=True)
