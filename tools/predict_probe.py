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
print("ok", float(out[2].max()))
