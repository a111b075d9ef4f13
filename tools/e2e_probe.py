import sys, time
sys.path.insert(0, ".")
import numpy as np, torch
from bench import make_samples, make_candidates
from pykriging_b200.krige import kriging
X, y = make_samples(1024, 10)
m = kriging(X, y)
h = m._pk_handle()
C, k = 1024, 10
th, p = make_candidates(C, k, 5)
tht = torch.from_numpy(th).pin_memory(); pt = torch.from_numpy(p).pin_memory()
out = torch.empty(4, C, dtype=torch.float64).pin_memory(); info = torch.empty(C, dtype=torch.int32).pin_memory()
thd, pd = tht.cuda(), pt.cuda(); outd = out.cuda(); infod = info.cuda()
torch.cuda.synchronize()
for rep in range(3):
    h.nll_batch_raw(C, thd, pd, None, 0, outd[0], outd[1], outd[2], outd[3], infod); h.synchronize()
h.set_profiling(True)
for mode in ("device", "host", "device", "host"):
    ts = []
    for rep in range(5):
        h.phase_times()
        t0 = time.perf_counter()
        if mode == "device":
            h.nll_batch_raw(C, thd, pd, None, 0, outd[0], outd[1], outd[2], outd[3], infod)
            t1 = time.perf_counter()
            h.synchronize()
        else:
            h.nll_batch_raw(C, tht.numpy(), pt.numpy(), None, 0, out[0].numpy(), out[1].numpy(), out[2].numpy(), out[3].numpy(), info.numpy())
            t1 = time.perf_counter()
        t2 = time.perf_counter()
        ph = h.phase_times()
        ts.append((round((t1 - t0) * 1e3, 2), round((t2 - t0) * 1e3, 2), {a: round(b, 2) for a, b in ph.items()}))
    print(mode, ts)
