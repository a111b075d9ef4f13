"""Host logic on CPU: the batched finite-difference stencil of the SLSQP polish (SciPy ``workers`` hook,
pykriging_b200/krige.py::_pk_slsqp) must walk exactly the iterates of SciPy's own sequential differences
(what the reference runs, krige.py:413).  The device objective is replaced by a smooth CPU function."""
import numpy as np

from pykriging_b200.krige import kriging


class _Host(kriging):
    """kriging with the device objective swapped for a closed-form one (no handle, no GPU)."""

    def __init__(self, k):
        self.k = k
        self.batches = []

    def fittingObjective(self, candidates, args):
        self.batches.append(len(candidates))
        out = []
        for c in candidates:
            c = np.asarray(c, dtype=float)
            if c[0] < 0.05:           # "Cholesky failed": the reference's penalty (krige.py:447-450)
                out.append(10000)
            else:
                out.append(np.float64(np.sum(100.0 * (c[1:] - c[:-1] ** 2) ** 2 + (1 - c[:-1]) ** 2)))
        return out


def test_batched_stencil_walks_the_sequential_iterates():
    bounds = [[0.0, 2.0]] * 4
    start = [1.7, 0.3, 1.9, 0.2]
    seq, bat = _Host(2), _Host(2)
    seq.slsqp_batched = False
    r0 = seq._pk_slsqp(start, bounds)
    r1 = bat._pk_slsqp(start, bounds)
    assert np.array_equal(r0.x, r1.x) and r0.fun == r1.fun
    assert r0.nit == r1.nit and r0.nfev == r1.nfev
    assert set(seq.batches) == {1}                       # one candidate per device call
    assert bat.batches.count(4) == bat._pk_stencil_calls  # one 2k-point population per gradient
    assert bat._pk_stencil_calls == r1.njev


def test_stencil_respects_bounds_and_penalty():
    """Start on the upper bound (SciPy flips the stencil inwards) next to the penalty region."""
    bounds = [[0.0, 2.0]] * 4
    start = [0.06, 2.0, 2.0, 0.0]
    seq, bat = _Host(2), _Host(2)
    seq.slsqp_batched = False
    r0 = seq._pk_slsqp(start, bounds)
    r1 = bat._pk_slsqp(start, bounds)
    assert np.array_equal(r0.x, r1.x) and r0.fun == r1.fun and r0.nit == r1.nit
