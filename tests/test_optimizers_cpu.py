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
