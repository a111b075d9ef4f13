"""Small predict-only workload for ncu: n = 512, k = 2, 2 Mi query points on the device."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import make_branin_samples  # noqa: E402
from pykriging_b200.krige import kriging  # noqa: E402

X, y = make_branin_samples(512)
model = kriging(X, y)
model.update([5.0, 5.0, 1.8, 1.8])
model.neglikelihood()
h = model._pk_handle()
if len(sys.argv) > 1:
    h.set_predict_variant(int(sys.argv[1]))
m = 1 << 21
Xq = torch.rand(m, 2, dtype=torch.float64, device="cuda")
out = torch.empty(3, m, dtype=torch.float64, device="cuda")
for _ in range(2):
    h.predict_batch_raw(m, Xq, np.array([5.0, 5.0]), np.array([1.8, 1.8]), float(model.mu), float(model.SigmaSqr), 0.0, -1.0,
                        0.0, out[0], out[1], out[2])
h.synchronize()
if os.environ.get("PK_PROBE_TIME"):
    # timing: 16 Mi-point grid (config C5) on the handle's stream
    g = torch.linspace(0, 1, 4096, dtype=torch.float64, device="cuda")
    grid = torch.stack(torch.meshgrid(g, g, indexing="ij"), dim=-1).reshape(-1, 2).contiguous()
    mg = grid.shape[0]
    o = torch.empty(3, mg, dtype=torch.float64, device="cuda")
    args = (mg, grid, np.array([5.0, 5.0]), np.array([1.8, 1.8]), float(model.mu), float(model.SigmaSqr), 0.0, -1.0, 0.0,
            o[0], o[1], o[2])
    h.predict_batch_raw(*args)
    h.synchronize()
    import time
    best = 1e9
    for _ in range(int(os.environ.get("PK_PROBE_REPS", "3"))):
        t0 = time.perf_counter()
        h.predict_batch_raw(*args)
        h.synchronize()
        best = min(best, time.perf_counter() - t0)
    print("predictions/s %.4g  (%.2f ms per 16 Mi points)" % (mg / best, best * 1e3))
print("ok", float(out[2].max()))
