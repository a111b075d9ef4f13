"""Host logic on CPU: the product GA/PSO drivers against the literal inspyred restatement
(oracle/inspyred_port.py, 'parity unpinned') under fixed seeds -- identical candidate sequences,
identical fitness traces, identical final populations."""
from random import Random

import numpy as np
import pytest

from oracle import inspyred_port as ip
from pykriging_b200 import optimizers as po


def _rosen(c):
    c = np.asarray(c, dtype=float)
    return float(np.sum(100.0 * (c[1:] - c[:-1] ** 2) ** 2 + (1 - c[:-1]) ** 2))


class _Problem(object):
    """Evaluator/generator/terminator with the signatures pyKriging hands to inspyred."""

    def __init__(self, lo, hi, penalise=False):
        self.lo, self.hi = lo, hi
        self.calls = []
        self.penalise = penalise

    def generate(self, random, args):
        b = args["_ec"].bounder
        return [random.uniform(lo, hi) for lo, hi in zip(b.lower_bound, b.upper_bound)]

    def evaluate(self, candidates, args):
        cands = [list(map(float, c)) for c in candidates]
        self.calls.append(cands)
        out = []
        for c in cands:
            f = _rosen(c)
            if self.penalise and c[0] > 1.5:
                f = 10000          # the int penalty of krige.py:447-450, produces ties
            out.append(f)
        return out

    def terminate(self, population, num_generations, num_evaluations, args):
        # krige.py:325-353
        max_generations = args.setdefault('max_generations', 10)
        previous_best = args.setdefault('previous_best', None)
        max_evaluations = args.setdefault('max_evaluations', 30000)
        current_best = np.around(max(population).fitness, decimals=4)
        if previous_best is None or previous_best != current_best:
            args['previous_best'] = current_best
            args['generation_count'] = 0
            return False or (num_evaluations >= max_evaluations)
        if args['generation_count'] >= max_generations:
            return True
        args['generation_count'] += 1
        return False or (num_evaluations >= max_evaluations)


def _run(kind, module, seed, dims, pop, penalise, **kw):
    lo, hi = [-2.0] * dims, [2.0] * dims
    prob = _Problem(lo, hi, penalise)
    ea = getattr(module, kind)(Random(seed))
    ea.terminator = prob.terminate
    final = ea.evolve(generator=prob.generate, evaluator=prob.evaluate, pop_size=pop, maximize=False,
                      bounder=module.Bounder(lo, hi), **kw)
    return prob.calls, [(list(map(float, i.candidate)), float(i.fitness)) for i in final], ea


@pytest.mark.parametrize("seed,dims,pop,penalise", [(1, 4, 30, False), (2, 6, 51, True), (3, 20, 64, True)])
def test_ga_matches_inspyred_port(seed, dims, pop, penalise):
    kw = dict(max_evaluations=3000, num_elites=5, mutation_rate=.05)
    a_calls, a_final, a = _run("GA", ip, seed, dims, pop, penalise, **kw)
    b_calls, b_final, b = _run("GA", po, seed, dims, pop, penalise, **kw)
    assert len(a_calls) == len(b_calls) and len(a_calls) > 3
    for ca, cb in zip(a_calls, b_calls):
        assert ca == cb
    assert a_final == b_final
    assert a.num_generations == b.num_generations and a.num_evaluations == b.num_evaluations


@pytest.mark.parametrize("seed,dims,pop,hood,penalise", [(4, 4, 30, 6, False), (5, 7, 155, 30, True), (6, 20, 40, 20, True)])
def test_pso_matches_inspyred_port(seed, dims, pop, hood, penalise):
    kw = dict(max_evaluations=4000, neighborhood_size=hood, num_inputs=dims)
    a_calls, a_final, a = _run("PSO", ip, seed, dims, pop, penalise, **kw)
    b_calls, b_final, b = _run("PSO", po, seed, dims, pop, penalise, **kw)
    assert len(a_calls) == len(b_calls) and len(a_calls) > 3
    for ca, cb in zip(a_calls, b_calls):
        assert ca == cb
    assert a_final == b_final
    assert a.num_generations == b.num_generations


def test_ga_only_recombines_initial_genes():
    """bit_flip_mutation is a no-op on real-valued genes -> every gene value ever
    evaluated already exists in the initial population at the same coordinate."""
    calls, _, _ = _run("GA", po, 7, 5, 20, False, max_evaluations=600, num_elites=2, mutation_rate=.05)
    init = np.array(calls[0])
    for gen in calls[1:]:
        g = np.array(gen)
        for d in range(g.shape[1]):
            assert set(g[:, d]).issubset(set(init[:, d]))


# ---- device-resident population (optimizers.DeviceGA / DevicePSO) on a NumPy stand-in of the C ABI ---------------------
class _FakeOptHandle(object):
    """pk_opt_set / pk_opt_get / pk_ga_step / pk_pso_step restated in NumPy (what csrc/pk_opt.cu does on the GPU), so
    that the HOST side of the device-resident optimisers -- draw order, vectorised sort / selection / topology, index
    bookkeeping -- can be checked on CPU against the host-array classes."""
    OPT_POP, OPT_CHILD, OPT_PREV, OPT_ARCHIVE = 0, 1, 2, 3

    def __init__(self):
        self.buf = {}
        self.started = False

    def opt_set(self, which, rows):
        rows = np.array(rows, dtype=float)
        if which == 0:
            self.n, self.dims = rows.shape
            self.buf = {b: np.zeros_like(rows) for b in range(4)}
            self.started = False
        self.buf[which] = rows.copy()

    def opt_get(self, which, row0, nrows, dims):
        return self.buf[which][row0:row0 + nrows].copy()

    def ga_step(self, nc, keep, parents, cut):
        if keep is not None:
            pool = np.concatenate([self.buf[1][:nc], self.buf[0]])
            self.buf[0] = pool[np.asarray(keep)].copy()
        for i in range(len(parents) // 2):
            mom, dad = self.buf[0][parents[2 * i]], self.buf[0][parents[2 * i + 1]]
            c = int(cut[i])
            if c == 0:
                self.buf[1][2 * i], self.buf[1][2 * i + 1] = mom, dad
            else:
                self.buf[1][2 * i] = np.concatenate([dad[:c], mom[c:]])
                self.buf[1][2 * i + 1] = np.concatenate([mom[:c], dad[c:]])

    def pso_step(self, r, nbest, improved, inertia, cognitive, social, lo, hi):
        x, arch = self.buf[0], self.buf[3]
        if improved is not None:
            m = np.asarray(improved) != 0
            arch[m] = x[m]
        prev = x if not self.started else self.buf[2]
        r = np.asarray(r).reshape(self.n, self.dims, 2)
        v = x + inertia * (x - prev) + cognitive * r[:, :, 0] * (arch - x) + social * r[:, :, 1] * (arch[np.asarray(nbest)] - x)
        self.buf[2] = x.copy()
        self.buf[0] = np.maximum(np.minimum(v, hi), lo)
        self.started = True


def _run_device(kind, seed, dims, pop, penalise, **kw):
    lo, hi = [-2.0] * dims, [2.0] * dims
    prob = _Problem(lo, hi, penalise)
    h = _FakeOptHandle()

    def evaluate(which, count):
        return prob.evaluate(h.buf[which][:count].tolist(), {})

    ea = getattr(po, kind)(Random(seed), h, evaluate)
    ea.terminator = prob.terminate
    final = ea.evolve(generator=prob.generate, pop_size=pop, maximize=False, bounder=po.Bounder(lo, hi), **kw)
    return prob.calls, [(list(map(float, i.candidate)), float(i.fitness)) for i in final], ea


@pytest.mark.parametrize("seed,dims,pop,penalise", [(1, 4, 30, False), (2, 6, 51, True), (3, 20, 64, True), (8, 21, 300, True)])
def test_device_resident_ga_follows_the_host_ga(seed, dims, pop, penalise):
    kw = dict(max_evaluations=3000 if pop < 100 else 9000, num_elites=5 if pop < 100 else 10, mutation_rate=.05)
    a_calls, a_final, a = _run("GA", po, seed, dims, pop, penalise, **kw)
    b_calls, b_final, b = _run_device("DeviceGA", seed, dims, pop, penalise, **kw)
    assert len(a_calls) == len(b_calls) and len(a_calls) > 3
    for g, (ca, cb) in enumerate(zip(a_calls, b_calls)):
        assert ca == cb, "generation %d" % g
    assert a_final == b_final
    assert a.num_generations == b.num_generations and a.num_evaluations == b.num_evaluations


@pytest.mark.parametrize("seed,dims,pop,hood,penalise", [(4, 4, 30, 6, False), (5, 7, 155, 30, True), (6, 20, 40, 20, True),
                                                         (9, 21, 150, 20, True)])
def test_device_resident_pso_follows_the_host_pso(seed, dims, pop, hood, penalise):
    kw = dict(max_evaluations=4000, neighborhood_size=hood, num_inputs=dims)
    a_calls, a_final, a = _run("PSO", po, seed, dims, pop, penalise, **kw)
    b_calls, b_final, b = _run_device("DevicePSO", seed, dims, pop, penalise, **kw)
    assert len(a_calls) == len(b_calls) and len(a_calls) > 3
    for g, (ca, cb) in enumerate(zip(a_calls, b_calls)):
        assert ca == cb, "generation %d" % g
    assert a_final == b_final
    assert a.num_generations == b.num_generations


def test_vectorised_draws_equal_the_generator_loop():
    """draw_uniforms hands the Mersenne-Twister state to NumPy and back: same doubles, same state afterwards."""
    for seed in (0, 7, 2024):
        a, b = Random(seed), Random(seed)
        for m in (1, 63, 64, 1000, 40960):
            va = po.draw_uniforms(a, m)
            vb = np.array([b.random() for _ in range(m)])
            assert np.array_equal(va, vb)
            assert a.sample(range(1, 20), 1) == b.sample(range(1, 20), 1) and a.uniform(1e-5, 100) == b.uniform(1e-5, 100)
        assert a.getstate() == b.getstate()
