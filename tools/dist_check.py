"""Run under torchrun (one rank per GPU): the sharded population evaluation and the sharded grid
prediction must reproduce the single-GPU result on every rank.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/dist_check.py
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pykriging_b200.krige import kriging  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
rng = np.random.default_rng(3)
n, k = 300, 4
X = rng.uniform(0, 1, (n, k))
y = np.sum((X - 0.4) ** 2, axis=1)
model = kriging(X, y, device=local)
crng = np.random.default_rng(4)
cands = np.concatenate([crng.uniform(1e-5, 100, (257, k)), crng.uniform(1, 2, (257, k))], axis=1)
cands[5, :k] = 1e-5
cands[5, k:] = 2.0  # fails
single = model.fittingObjective(cands.tolist(), {})
model.enable_sharding(min_points=1000)
sharded = model.fittingObjective(cands.tolist(), {})
assert len(single) == len(sharded) == 257
assert all((a == b) or (a != a and b != b) for a, b in zip(single, sharded)), "sharded fitness differs"
assert sharded[5] == 10000
model.update(cands[0])
model.neglikelihood()
Xq = np.random.default_rng(5).uniform(0, 1, (50001, k))
model.disable_sharding()
a = model.predict_all_batch(Xq)
model.enable_sharding(min_points=1000)
b = model.predict_all_batch(Xq)
for nm in ("yhat", "s", "ei"):
    assert np.array_equal(a[nm], b[nm]), nm
# GA under a fixed seed with a sharded population selects the same best theta/p on every rank
model.train("ga", seed=7, pop_size=64, max_evaluations=640)
best = np.concatenate([model.theta, model.pl])
t = torch.from_numpy(best).cuda()
ref = t.clone()
dist.broadcast(ref, src=0)
assert torch.equal(t, ref), "ranks diverged"
dist.barrier()
if rank == 0:
    print("dist_check ok: world", world, "best", best)
dist.destroy_process_group()
