"""Cholesky of a population: one CTA per candidate (variant 2) vs task mode (variant 4, with its knobs) vs the automatic
choice, over the shapes the verdict names: n = 1024 with 128 (a 1024-population sharded over 8 GPUs), 148, 296, 300
(reference default) and 1024 candidates; n = 4096 regression with 8 / 64 / 150 candidates."""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pykriging_b200._lib import Handle  # noqa: E402

h = Handle(0)
dev = torch.device("cuda", 0)
shapes = [(1024, 10, 0, (1, 21, 64, 128, 148, 200, 300, 444, 1024)), (4096, 6, 1, (8, 150)),
          (2048, 8, 0, (128, 300)), (512, 6, 0, (300,))]
if os.environ.get("PK_SWEEP_QUICK"):
    shapes = [(1024, 10, 0, (128, 300, 1024))]
# (name, cholesky variant, single_below, blocks_per_task, skew)
configs = [("v2", 2, -1, 0, 0), ("auto", 0, -1, 0, 0), ("t_0_1", 4, 0, 1, 1), ("t_4_2", 4, 4, 2, 1), ("t_0_2", 4, 0, 2, 1),
           ("t_0_4", 4, 0, 4, 1), ("t_0_30", 4, 0, 30, 1)]
for n, k, mode, Cs in shapes:
    rng = np.random.default_rng(n)
    X = np.zeros((n, k))
    for d in range(k):
        X[:, d] = (rng.permutation(n) + 0.5) / n
    y = np.sum((X - 0.3) ** 2, axis=1) + (rng.normal(0, 0.05, n) if mode else 0.0)
    y = (y - y.min()) / (y.max() - y.min())
    h.set_data(X, y)
    for C in Cs:
        th = torch.from_numpy(rng.uniform(1e-4, 100, (C, k))).to(dev)
        p = torch.from_numpy(rng.uniform(1.81231 if mode else 1.0, 2.0, (C, k))).to(dev)
        lam = torch.from_numpy(rng.uniform(0, 1, C)).to(dev) if mode else None
        out = torch.empty(4, C, dtype=torch.float64, device=dev)
        info = torch.empty(C, dtype=torch.int32, device=dev)
        row = {"n": n, "C": C}
        base = None
        for name, variant, sb, bpt, skew in configs:
            if name == "gang" and C > 98:
                continue
            if name == "v2" and n >= 4096 and C < 64:
                continue
            h.set_cholesky_variant(variant)
            h.set_task_options(sb, bpt, 0, skew)
            h.set_profiling(True)
            reps = 3 if n < 4096 else 2
            for rep in range(reps + 1):
                if rep == 1:
                    h.synchronize()
                    h.phase_times()
                h.nll_batch_raw(C, th, p, lam, mode, out[0], out[1], out[2], out[3], info)
            h.synchronize()
            ms = h.phase_times()["cholesky"] / reps
            h.set_profiling(False)
            res = out.clone()
            if base is None and name in ("v2", "t_0_1"):
                base = res
            same = base is not None and bool(torch.equal(torch.nan_to_num(res, nan=-1.0), torch.nan_to_num(base, nan=-1.0)))
            row[name] = {"ms": round(ms, 3), "tf": round(C * n ** 3 / 3 / (ms * 1e-3) * 1e-12, 2), "same_bits_as_v2": same}
        print(json.dumps(row), flush=True)
h.set_cholesky_variant(0)
h.set_task_options()
