"""World-size-2 (and 3) gloo tests of the multi-GPU host logic: candidate / query-grid sharding,
the scalar all_gather, the global arg-max and the elite broadcast.  The per-shard evaluator is a
NumPy stand-in (the CUDA library cannot run here); what is under test is that every rank ends up
with exactly the single-process result."""
import os
import socket
import sys

import numpy as np
import pytest
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _fake_nll(theta, p):
    # deterministic stand-in with the output layout of Handle.nll_batch
    nll = np.sum(theta * np.sin(p), axis=1)
    info = (theta[:, 0] < 0.05).astype(np.float64) * 3
    return np.stack([nll, nll * 0.5, nll * nll, np.cos(nll), info], axis=1)


def _worker(rank, world, port, count, q):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist

    from pykriging_b200 import sharding

    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    try:
        rng = np.random.default_rng(5)
        theta, p = rng.uniform(0, 1, (count, 4)), rng.uniform(1, 2, (count, 4))
        calls = []

        def eval_slice(lo, hi):
            calls.append((lo, hi))
            return _fake_nll(theta[lo:hi], p[lo:hi])

        # count every collective the sharded evaluation issues: the contract is ONE all_gather per generation
        ncoll = {"n": 0}
        for name in ("all_gather_into_tensor", "all_reduce", "all_gather", "broadcast", "barrier"):
            orig = getattr(dist, name)

            def counted(*a, _orig=orig, **kw):
                ncoll["n"] += 1
                return _orig(*a, **kw)

            setattr(dist, name, counted)
        rows = sharding.sharded_rows(count, eval_slice, ncols=5)
        assert ncoll["n"] == 1, ncoll
        # a failure on ONE rank reaches every rank through the same single collective, and the group stays usable
        def eval_fail(lo, hi):
            if rank == world - 1:
                raise RuntimeError("boom on rank %d" % rank)
            return _fake_nll(theta[lo:hi], p[lo:hi])

        if count >= world:  # every rank owns a slice
            ncoll["n"] = 0
            try:
                sharding.sharded_rows(count, eval_fail, ncols=5)
                raise AssertionError("no exception on rank %d" % rank)
            except RuntimeError as e:  # ShardError is a RuntimeError
                assert ("boom" in str(e)) == (rank == world - 1), str(e)
            assert ncoll["n"] == 1, ncoll
            again = sharding.sharded_rows(count, eval_slice, ncols=5)
            assert np.array_equal(again, rows)
            calls[:] = calls[: len(calls) // 2]
        lo, hi = sharding.shard_bounds(count, rank, world)
        if count >= 64:
            # a population handed over as a list of lists: with sharding enabled only this rank's slice is converted
            # (matrixops._pk_candidate_rows) -- the rows of the other ranks are poisoned here and must not be touched
            from pykriging_b200.matrixops import matrixops

            m = object.__new__(matrixops)
            m.__dict__["_pk_shard"] = (None, 1)
            full = np.hstack([theta, p]).tolist()
            poisoned = [row if lo <= i < hi else ["poison"] * 8 for i, row in enumerate(full)]
            arr, get = m._pk_candidate_rows(poisoned)
            assert arr.shape == (count, 8) and np.array_equal(arr[lo:hi], np.hstack([theta, p])[lo:hi])
            assert np.array_equal(get(lo), arr[lo])
            arr2, get2 = m._pk_candidate_rows(full)
            other = 0 if lo > 0 else count - 1
            assert np.array_equal(get2(other), np.asarray(full[other])) and np.array_equal(get2(hi - 1), arr2[hi - 1])
            m.__dict__.pop("_pk_shard")
            arr3, get3 = m._pk_candidate_rows(full)   # not sharded: the whole list
            assert np.array_equal(arr3, np.hstack([theta, p])) and np.array_equal(get3(other), arr3[other])
        vals = rows[lo:hi, 0]
        best = sharding.global_argmax(vals, lo)
        elite = sharding.broadcast_candidate(np.arange(9.0) + 100 * rank, src=world - 1)
        q.put((rank, rows, calls, best, elite))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,count", [(2, 300), (2, 7), (3, 10), (2, 1)])
def test_population_sharding_matches_single_process(world, count):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, count, q)) for r in range(world)]
    for pr in procs:
        pr.start()
    results = [q.get(timeout=120) for _ in range(world)]
    for pr in procs:
        pr.join(timeout=60)
        assert pr.exitcode == 0
    rng = np.random.default_rng(5)
    theta, p = rng.uniform(0, 1, (count, 4)), rng.uniform(1, 2, (count, 4))
    want = _fake_nll(theta, p)
    covered = []
    for rank, rows, calls, best, elite in results:
        assert rows.shape == want.shape
        assert np.array_equal(rows, want)  # bit-identical on every rank
        covered += calls
        assert best[1] == int(np.argmax(want[:, 0])) and best[0] == want[:, 0].max()
        assert np.array_equal(elite, np.arange(9.0) + 100 * (world - 1))
    # the slices partition [0, count) exactly once
    covered = sorted(c for c in covered if c[1] > c[0])
    assert covered[0][0] == 0 and covered[-1][1] == count
    for a, b in zip(covered, covered[1:]):
        assert a[1] == b[0]


def test_shard_bounds_partition():
    from pykriging_b200.sharding import shard_bounds

    for count in (0, 1, 7, 8, 1024, 16777216):
        for world in (1, 2, 3, 8):
            spans = [shard_bounds(count, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == count
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            assert max(h - l for l, h in spans) - min(h - l for l, h in spans) <= -(-count // world)


def test_single_process_passthrough():
    from pykriging_b200 import sharding

    rows = sharding.sharded_rows(5, lambda lo, hi: np.arange(lo, hi, dtype=float)[:, None] * np.ones((1, 3)))
    assert rows.shape == (5, 3) and rows[4, 2] == 4.0
    assert sharding.global_argmax(np.array([1.0, 5.0, 5.0]), 10) == (5.0, 11)


def test_sharded_paths_bracket_their_slice_with_the_shard_hint():
    """The sharded glue of matrixops tells the handle the size of the WHOLE population / query set before it
    evaluates its slice and clears the hint afterwards (kernel selection then follows the whole: bit-identical
    shards), also when the evaluation raises.  Fake handle, single process (no process group: one slice)."""
    from pykriging_b200.matrixops import matrixops

    class FakeHandle(object):
        def __init__(self):
            self.log = []

        def set_shard_hint(self, population=0, query_points=0):
            self.log.append(("hint", int(population), int(query_points)))

        def set_shard_offset(self, first_candidate=0):
            self.log.append(("offset", int(first_candidate)))

        def nll_batch(self, theta, p, lam, mode):
            self.log.append(("nll", theta.shape[0]))
            z = np.zeros(theta.shape[0])
            return {"nll": z, "mu": z, "sigma2": z, "lndet": z, "info": np.zeros(theta.shape[0], dtype=np.int32)}

        def predict_batch(self, Xq, theta, pl, mu, s2, y_min=0.0, ei_weight=-1.0, var_extra=0.0, want=("yhat", "s", "ei")):
            self.log.append(("predict", Xq.shape[0]))
            if Xq.shape[0] == 7:
                raise RuntimeError("boom")
            return {nm: (np.zeros(Xq.shape[0]) if nm in want else None) for nm in ("yhat", "s", "ei")}

    m = matrixops.__new__(matrixops)
    fake = FakeHandle()
    m.__dict__["_pk_handle"] = lambda: fake
    m.__dict__["_pk_shard"] = (None, 1)
    m.__dict__.update(_has_U=True, _pending_factor=None, mu=0.5, SigmaSqr=1.0, theta=np.ones(2), pl=np.ones(2) * 2, k=2)
    out = m._pk_nll_population(np.ones((13, 2)), np.ones((13, 2)), None, 0)
    assert out["nll"].shape == (13,)
    assert fake.log == [("hint", 13, 0), ("offset", 0), ("nll", 13), ("hint", 0, 0)]
    fake.log.clear()
    batched = m._pk_predict
    batched(np.zeros((11, 2)), ("yhat",))
    assert fake.log == [("hint", 0, 11), ("predict", 11), ("hint", 0, 0)]
    fake.log.clear()
    with pytest.raises(RuntimeError):
        batched(np.zeros((7, 2)), ("yhat",))
    assert fake.log[0] == ("hint", 0, 7) and fake.log[-1] == ("hint", 0, 0)
