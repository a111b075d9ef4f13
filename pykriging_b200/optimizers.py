"""GA / PSO drivers behind ``train()`` and ``infill()``.

The reference delegates these to the un-vendored, un-pinned third-party package ``inspyred``
(pyKriging/krige.py:276-291, 373-399; regressionkrige.py:272-287, 370-396).  This module supplies
the two algorithms with inspyred 1.0.x semantics -- same operator order, same
consumption of ``random.Random`` -- but with the population held as arrays and the whole population
evaluated by ONE batched device call per generation.  ``oracle/inspyred_port.py`` is the literal,
object-per-individual CPU restatement this module is checked against under fixed seeds.

Ordering convention (minimisation): ``better(a, b)`` <=> fitness a < fitness b.  inspyred wraps
candidates in ``Individual(maximize=False)`` whose ``<`` is "is worse than"; sorts and ``max`` below are
written to produce the identical permutations, including tie handling (Python's stable sort).
"""
from random import Random

import numpy as np


class Bounder(object):
    """inspyred.ec.Bounder: clip to [lower, upper] coordinate-wise."""

    def __init__(self, lower_bound=None, upper_bound=None):
        self.lower_bound = lower_bound
        self.upper_bound = upper_bound

    def __call__(self, candidate, args=None):
        if self.lower_bound is None or self.upper_bound is None:
            return candidate
        lo = np.asarray(self.lower_bound, dtype=float)
        hi = np.asarray(self.upper_bound, dtype=float)
        return np.maximum(np.minimum(np.asarray(candidate, dtype=float), hi), lo)


class _Worse(object):
    """Sort key reproducing ``Individual.__lt__`` for maximize=False: a < b  <=>  a is worse."""
    __slots__ = ("f",)

    def __init__(self, f):
        self.f = f

    def __lt__(self, other):
        return self.f > other.f


def _order_worst_first(fitness):
    """population.sort(): ascending under Individual.__lt__ (worst first), stable."""
    return sorted(range(len(fitness)), key=lambda i: _Worse(fitness[i]))


def _order_best_first(fitness):
    """population.sort(reverse=True): best first; Python keeps the original order of equals."""
    return sorted(range(len(fitness)), key=lambda i: _Worse(fitness[i]), reverse=True)


def _first_best(fitness, indices):
    """max(hood) over Individuals: first element not beaten by a later one (ties -> earliest)."""
    best = indices[0]
    fb = fitness[best]
    for i in indices[1:]:
        if fb > fitness[i]:  # (hood[i] > best)  <=>  best < hood[i]  <=>  best is worse
            best, fb = i, fitness[i]
    return best


class Individual(object):
    """Minimal stand-in for inspyred.ec.Individual (what train() iterates over)."""

    def __init__(self, candidate, fitness, maximize=False):
        self.candidate = candidate
        self.fitness = fitness
        self.maximize = maximize

    def __lt__(self, other):
        if self.maximize:
            return self.fitness < other.fitness
        return self.fitness > other.fitness

    def __gt__(self, other):
        return other < self

    def __repr__(self):
        return "{0} : {1}".format(self.candidate, self.fitness)


class _EC(object):
    def __init__(self, random=None):
        self._random = random if random is not None else Random()
        self.terminator = None
        self.bounder = None
        self.num_evaluations = 0
        self.num_generations = 0
        self.population = []
        self.history = []  # best fitness per generation (diagnostic)

    def _generate(self, generator, pop_size, args):
        return [list(generator(random=self._random, args=args)) for _ in range(pop_size)]

    @staticmethod
    def _as_individuals(cands, fits):
        return [Individual(list(map(float, c)), f) for c, f in zip(cands, fits)]

    def _terminate(self, cands, fits, args):
        pop = self._as_individuals(cands, fits)
        return bool(self.terminator(population=pop, num_generations=self.num_generations,
                                    num_evaluations=self.num_evaluations, args=args))


class GA(_EC):
    """inspyred.ec.GA: rank selection, 1-point crossover (rate 1.0), bit-flip mutation (a no-op on
    real-valued genes), generational replacement with elitism."""

    def evolve(self, generator, evaluator, pop_size=100, maximize=False, bounder=None, **args):
        assert not maximize
        args = dict(args)
        args["_ec"] = self
        args.setdefault("num_selected", pop_size)
        self.bounder = bounder if bounder is not None else Bounder()
        rnd = self._random
        cands = self._generate(generator, pop_size, args)
        fits = list(evaluator(candidates=cands, args=args))
        self.num_evaluations = len(fits)
        self.num_generations = 0
        while not self._terminate(cands, fits, args):
            # ---- rank_selection -------------------------------------------------------------
            num_selected = args["num_selected"]
            order = _order_worst_first(fits)
            cands = [cands[i] for i in order]
            fits = [fits[i] for i in order]
            n = len(cands)
            den = (n * (n + 1)) / 2.0
            psum = [(i + 1) / den for i in range(n)]
            for i in range(1, n):
                psum[i] += psum[i - 1]
            parents = []
            for _ in range(num_selected):
                cutoff = rnd.random()
                lower, upper = 0, n - 1
                while upper >= lower:
                    mid = (lower + upper) // 2
                    if psum[mid] > cutoff:
                        upper = mid - 1
                    else:
                        lower = mid + 1
                lower = max(0, min(n - 1, lower))
                parents.append(lower)
            off = [list(cands[i]) for i in parents]
            # ---- n_point_crossover (num_crossover_points = 1, crossover_rate = 1.0) ------------
            crossover_rate = args.setdefault("crossover_rate", 1.0)
            num_points = args.setdefault("num_crossover_points", 1)
            if len(off) % 2 == 1:
                off = off[:-1]
            children = []
            for mom, dad in zip(off[::2], off[1::2]):
                if rnd.random() < crossover_rate:
                    num_cuts = min(len(mom) - 1, num_points)
                    cut_points = rnd.sample(range(1, len(mom)), num_cuts)
                    cut_points.sort()
                    bro, sis = list(dad), list(mom)
                    normal = True
                    for i, (m, d) in enumerate(zip(mom, dad)):
                        if i in cut_points:
                            normal = not normal
                        if not normal:
                            bro[i] = m
                            sis[i] = d
                    children.append(bro)
                    children.append(sis)
                else:
                    children.append(mom)
                    children.append(dad)
            # ---- bit_flip_mutation: acts only on all-0/1 candidates -----------------------------
            rate = args.setdefault("mutation_rate", 0.1)
            for c in children:
                if all((x in (0, 1)) for x in c):
                    for i, m in enumerate(c):
                        if rnd.random() < rate:
                            c[i] = (m + 1) % 2
            off_fits = list(evaluator(candidates=children, args=args))
            self.num_evaluations += len(off_fits)
            # ---- generational_replacement ---------------------------------------------------------
            num_elites = args.setdefault("num_elites", 0)
            order = _order_best_first(fits)
            elite = order[:num_elites]
            pool_c = children + [cands[i] for i in elite]
            pool_f = off_fits + [fits[i] for i in elite]
            order = _order_best_first(pool_f)[:n]
            cands = [pool_c[i] for i in order]
            fits = [pool_f[i] for i in order]
            self.num_generations += 1
            self.history.append(fits[0])
        self.population = self._as_individuals(cands, fits)
        return self.population


class PSO(_EC):
    """inspyred.swarm.PSO with ring topology: inertia 0.5, cognitive = social = 2.1."""

    def __init__(self, random=None):
        _EC.__init__(self, random)
        self.topology = "ring"

    def evolve(self, generator, evaluator, pop_size=100, maximize=False, bounder=None, **args):
        assert not maximize
        args = dict(args)
        args["_ec"] = self
        self.bounder = bounder if bounder is not None else Bounder()
        rnd = self._random
        cands = self._generate(generator, pop_size, args)
        dims = len(cands[0])
        x = np.array(cands, dtype=float)
        fits = list(evaluator(candidates=[list(c) for c in x], args=args))
        self.num_evaluations = len(fits)
        self.num_generations = 0
        arch_x, arch_f = x.copy(), list(fits)   # personal bests (inspyred's archive)
        prev_x = None
        inertia = args.setdefault("inertia", 0.5)
        cognitive = args.setdefault("cognitive_rate", 2.1)
        social = args.setdefault("social_rate", 2.1)
        hood_size = args.setdefault("neighborhood_size", 3)
        half = hood_size // 2
        while not self._terminate(x, fits, args):
            if prev_x is None:
                prev_x = x.copy()
            n = len(x)
            nbest = np.empty(n, dtype=int)
            for i in range(n):
                start = (n - half + i) if i < half else (i - half)
                hood = [(start + j) % n for j in range(hood_size)]
                nbest[i] = _first_best(arch_f, hood)
            r = np.array([rnd.random() for _ in range(n * dims * 2)], dtype=float).reshape(n, dims, 2)
            r1, r2 = r[:, :, 0], r[:, :, 1]
            nb = arch_x[nbest]
            value = x + inertia * (x - prev_x) + cognitive * r1 * (arch_x - x) + social * r2 * (nb - x)
            new_x = self.bounder(value, args)
            new_f = list(evaluator(candidates=[list(c) for c in new_x], args=args))
            self.num_evaluations += len(new_f)
            prev_x, x, fits = x, np.asarray(new_x, dtype=float), new_f
            # _swarm_archiver: keep the archived particle only when the new one is strictly worse
            for i in range(n):
                if not (fits[i] > arch_f[i]):
                    arch_x[i] = x[i]
                    arch_f[i] = fits[i]
            self.num_generations += 1
            self.history.append(min(fits))
        self.population = self._as_individuals(x, fits)
        return self.population
