"""Wall-clock breakdown of kriging.train() (GA + SLSQP polish) on a C3-shaped model: where the time goes between the
device (pk_nll_batch / pk_fit) and the host glue (inspyred-compatible GA in Python, SciPy SLSQP)."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import make_samples  # noqa: E402
from pykriging_b200.krige import kriging  # noqa: E402

n, k, pop = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
X, y = make_samples(n, k)
model = kriging(X, y)
stats = {"pop_calls": 0, "pop_s": 0.0, "pop_cands": 0, "loc_calls": 0, "loc_s": 0.0}
fo, fl = model.fittingObjective, model.fittingObjective_local


def fo_timed(candidates, args):
    cands = candidates
    t0 = time.perf_counter()
    r = fo(cands, args)
    stats["pop_s"] += time.perf_counter() - t0
    stats["pop_calls"] += 1
    stats["pop_cands"] += len(cands)
    return r


def fl_timed(entry):
    t0 = time.perf_counter()
    r = fl(entry)
    stats["loc_s"] += time.perf_counter() - t0
    stats["loc_calls"] += 1
    return r


model.fittingObjective = fo_timed
model.fittingObjective_local = fl_timed
t0 = time.perf_counter()
model.train("ga", seed=3, pop_size=pop, max_evaluations=int(sys.argv[4]) if len(sys.argv) > 4 else 30000)
total = time.perf_counter() - t0
print("n %d k %d pop %d: train %.3f s; population calls %d (%d candidates) %.3f s; single evaluations %d %.3f s; "
      "host glue (GA operators, SLSQP, wrappers) %.3f s" % (n, k, pop, total, stats["pop_calls"], stats["pop_cands"],
                                                              stats["pop_s"], stats["loc_calls"], stats["loc_s"],
                                                              total - stats["pop_s"] - stats["loc_s"]))
