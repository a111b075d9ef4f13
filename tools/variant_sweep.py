"""Cholesky variant 1 (per-block-column launches) vs 2 (persistent CTA per candidate) over (n, C): the data behind
launch_cholesky's automatic choice."""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pykriging_b200._lib import Handle  # noqa: E402

h = Handle(0)
dev = torch.device("cuda", 0)
k = 6
for n in (256, 512, 1024, 2048, 4096):
    rng = np.random.default_rng(n)
    X = np.zeros((n, k))
    for d in range(k):
        X[:, d] = (rng.permutation(n) + 0.5) / n
    y = np.sum((X - 0.3) ** 2, axis=1) + rng.normal(0, 0.05, n)
    y = (y - y.min()) / (y.max() - y.min())
    h.set_data(X, y)
    for C in ((1, 2, 4, 8, 16, 21, 32, 48, 64, 96, 128, 148) if os.environ.get("PK_SWEEP_SMALL") else (8, 16, 32, 64, 96, 128, 160, 224, 296, 320, 444, 592)):
        if n >= 4096 and C > 320:
            continue
        th = torch.from_numpy(rng.uniform(1e-4, 100, (C, k))).to(dev)
        p = torch.from_numpy(rng.uniform(1.81231, 2.0, (C, k))).to(dev)
        lam = torch.from_numpy(rng.uniform(0, 1, C)).to(dev)
        out = torch.empty(4, C, dtype=torch.float64, device=dev)
        info = torch.empty(C, dtype=torch.int32, device=dev)
        res = {}
        for variant in (1, 2, 3):
            h.set_cholesky_variant(variant)
            h.set_profiling(True)
            for rep in range(3):
                if rep == 1:
                    h.synchronize()
                    h.phase_times()
                h.nll_batch_raw(C, th, p, lam, 1, out[0], out[1], out[2], out[3], info)
            h.synchronize()
            res[variant] = h.phase_times()["cholesky"] / 2
        print(json.dumps({"n": n, "C": C, "v1_ms": round(res[1], 3), "v2_ms": round(res[2], 3),
                          "v1_tf": round(C * n ** 3 / 3 / (res[1] * 1e-3) * 1e-12, 2),
                          "v2_tf": round(C * n ** 3 / 3 / (res[2] * 1e-3) * 1e-12, 2), "v3_ms": round(res[3], 3),
                          "v3_tf": round(C * n ** 3 / 3 / (res[3] * 1e-3) * 1e-12, 2)}), flush=True)
