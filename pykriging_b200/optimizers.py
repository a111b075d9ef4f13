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


# --------------------------------------------------------------------------------------------------------------------
# Device-resident population: the same two algorithms, with the candidate vectors kept in
# GPU buffers between generations (pk_opt_set / pk_opt_evaluate / pk_ga_step / pk_pso_step).  What fixes the trajectory
# under a seed stays on the host, bit for bit: every draw of the `random.Random` generator, in the order the classes
# above make them, and every comparison of fitness values (sort permutations, ring-topology bests) -- all on the
# per-candidate scalars the evaluator returns.  What goes up per generation are index arrays (GA: survivors, parents,
# cut points) or uniforms (PSO); the recombination / velocity update itself runs on the device.
# --------------------------------------------------------------------------------------------------------------------
def draw_uniforms(rnd, m):
    """``np.array([rnd.random() for _ in range(m)])`` without the Python loop: the generator's Mersenne-Twister state
    is handed to NumPy's MT19937 (same `genrand_res53` doubles), advanced there and handed back."""
    m = int(m)
    try:
        state = rnd.getstate()
        ok = type(rnd) is Random and state[0] == 3 and len(state[1]) == 625
    except Exception:  # noqa: BLE001 -- SystemRandom and friends have no state
        ok = False
    if not ok or m < 64:
        return np.array([rnd.random() for _ in range(m)], dtype=float)
    rs = np.random.RandomState()
    rs.set_state(("MT19937", np.array(state[1][:624], dtype=np.uint32), int(state[1][624])))
    vals = rs.random_sample(m)
    ns = rs.get_state()
    rnd.setstate((3, tuple(int(v) for v in ns[1]) + (int(ns[2]),), state[2]))
    return vals


def _argsort_worst_first(fits):
    """Permutation of ``population.sort()`` under ``Individual.__lt__`` (maximize=False): descending fitness, stable."""
    f = np.asarray(fits, dtype=float)
    if np.isnan(f).any():
        return np.array(_order_worst_first(list(fits)), dtype=np.int64)  # NaN compares false both ways: keep Python's result
    return np.argsort(-f, kind="stable")


def _argsort_best_first(fits):
    """``population.sort(reverse=True)``: ascending fitness; Python's reversed sort keeps equal elements in their
    original order, and so does a stable argsort of the values themselves."""
    f = np.asarray(fits, dtype=float)
    if np.isnan(f).any():
        return np.array(_order_best_first(list(fits)), dtype=np.int64)
    return np.argsort(f, kind="stable")


class _LazyIndividual(Individual):
    """Individual of a device-resident population: the fitness is on the host, the candidate vector is read back from
    the GPU only if somebody asks for it (a terminator / observer looking at `candidate`, the final population)."""

    def __init__(self, owner, which, row, fitness):
        self._owner, self._which, self._row = owner, which, row
        self._cand = None
        self.fitness = fitness
        self.maximize = False

    @property
    def candidate(self):
        if self._cand is None:
            self._cand = [float(v) for v in self._owner._rows(self._which)[self._row]]
        return self._cand

    @candidate.setter
    def candidate(self, value):
        self._cand = value


class _DeviceEC(_EC):
    """Shared plumbing: `evaluate(which, count) -> list of fitness` is the model's device objective
    (kriging._pk_device_objective), `handle` its _lib.Handle."""

    def __init__(self, random, handle, evaluate):
        _EC.__init__(self, random)
        self._handle = handle
        self._evaluate = evaluate
        self._row_cache = {}
        self._dims = 0
        self.host_seconds = 0.0  # time spent outside pk_* calls of evolve's loop (diagnostic)

    def _rows(self, which):
        rows = self._row_cache.get(which)
        if rows is None:
            rows = self._handle.opt_get(which, 0, self._n, self._dims)
            self._row_cache[which] = rows
        return rows

    def _individuals(self, which, fits):
        return [_LazyIndividual(self, which, i, f) for i, f in enumerate(fits)]

    def _terminate_lazy(self, which, fits, args):
        self._row_cache = {}
        pop = self._individuals(which, fits)
        return bool(self.terminator(population=pop, num_generations=self.num_generations,
                                    num_evaluations=self.num_evaluations, args=args))


class DeviceGA(_DeviceEC):
    """`GA` above with the population resident on the GPU: one pk_ga_step + one pk_opt_evaluate per generation."""

    def evolve(self, generator, evaluator=None, pop_size=100, maximize=False, bounder=None, **args):
        assert not maximize
        args = dict(args)
        args["_ec"] = self
        args.setdefault("num_selected", pop_size)
        self.bounder = bounder if bounder is not None else Bounder()
        rnd, h = self._random, self._handle
        cands = np.array(self._generate(generator, pop_size, args), dtype=float)
        n, dims = cands.shape
        self._n, self._dims = n, dims
        h.opt_set(h.OPT_POP, cands)
        fits = list(self._evaluate(h.OPT_POP, n))
        self.num_evaluations = len(fits)
        self.num_generations = 0
        while not self._terminate_lazy(h.OPT_POP, fits, args):
            # ---- rank_selection: Python's sort permutation and the binary search, vectorised ------------------
            # (population.sort() reorders the population itself, worst first: the replacement below sees that order,
            # which decides between candidates of equal fitness -- the rows stay where they are, `order` maps to them)
            num_selected = args["num_selected"]
            order = _argsort_worst_first(fits)
            fits = [fits[i] for i in order]
            psum = np.cumsum(np.arange(1, n + 1, dtype=float) / ((n * (n + 1)) / 2.0))
            cutoffs = draw_uniforms(rnd, num_selected)
            sel = np.clip(np.searchsorted(psum, cutoffs, side="right"), 0, n - 1)
            parents = order[sel]
            if len(parents) % 2 == 1:
                parents = parents[:-1]
            # ---- n_point_crossover (one cut point): the draws, in inspyred's order --------------------------
            crossover_rate = args.setdefault("crossover_rate", 1.0)
            num_points = args.setdefault("num_crossover_points", 1)
            npairs = len(parents) // 2
            cut = np.zeros(npairs, dtype=np.int32)
            num_cuts = min(dims - 1, num_points)
            if num_cuts > 1:
                raise NotImplementedError("device-resident GA implements the reference's single cut point")
            for i in range(npairs):
                if rnd.random() < crossover_rate:
                    pts = rnd.sample(range(1, dims), num_cuts)
                    if pts:
                        cut[i] = pts[0]
                    else:  # a one-gene candidate has no cut point: inspyred returns copies of (dad, mom)
                        parents[2 * i], parents[2 * i + 1] = parents[2 * i + 1], parents[2 * i]
            # bit_flip_mutation acts only on candidates made of 0s and 1s: a no-op on [theta, p] vectors
            args.setdefault("mutation_rate", 0.1)
            h.ga_step(0, None, parents, cut)
            nchild = 2 * npairs
            off_fits = list(self._evaluate(h.OPT_CHILD, nchild))
            self.num_evaluations += nchild
            # ---- generational_replacement with elitism: survivors of [offspring | elites], best first ---------
            num_elites = args.setdefault("num_elites", 0)
            elite = _argsort_best_first(fits)[:num_elites]       # positions in the SORTED population
            pool_f = off_fits + [fits[i] for i in elite]
            keep_pool = _argsort_best_first(pool_f)[:n]
            if num_elites:
                elite_rows = order[elite]                          # the device rows they live in
                keep = np.where(keep_pool < nchild, keep_pool, nchild + elite_rows[np.maximum(keep_pool - nchild, 0)])
            else:
                keep = keep_pool
            h.ga_step(nchild, keep.astype(np.int32), np.zeros(0, dtype=np.int32), np.zeros(0, dtype=np.int32))
            fits = [pool_f[i] for i in keep_pool]
            self.num_generations += 1
            self.history.append(fits[0])
        self._row_cache = {}
        self.population = self._individuals(h.OPT_POP, fits)
        for ind in self.population:
            ind.candidate  # noqa: B018 -- materialise: the buffers will be reused
        return self.population


class DevicePSO(_DeviceEC):
    """`PSO` above with positions, previous positions and personal bests resident on the GPU."""

    def __init__(self, random, handle, evaluate):
        _DeviceEC.__init__(self, random, handle, evaluate)
        self.topology = "ring"

    def evolve(self, generator, evaluator=None, pop_size=100, maximize=False, bounder=None, **args):
        assert not maximize
        args = dict(args)
        args["_ec"] = self
        self.bounder = bounder if bounder is not None else Bounder()
        rnd, h = self._random, self._handle
        x0 = np.array(self._generate(generator, pop_size, args), dtype=float)
        n, dims = x0.shape
        self._n, self._dims = n, dims
        if self.bounder.lower_bound is None or self.bounder.upper_bound is None:
            lo, hi = np.full(dims, -np.inf), np.full(dims, np.inf)
        else:
            lo = np.asarray(self.bounder.lower_bound, dtype=float)
            hi = np.asarray(self.bounder.upper_bound, dtype=float)
        h.opt_set(h.OPT_POP, x0)
        h.opt_set(h.OPT_ARCHIVE, x0)
        fits = list(self._evaluate(h.OPT_POP, n))
        self.num_evaluations = n
        self.num_generations = 0
        arch_f = np.array(fits, dtype=float)
        improved = None
        inertia = args.setdefault("inertia", 0.5)
        cognitive = args.setdefault("cognitive_rate", 2.1)
        social = args.setdefault("social_rate", 2.1)
        hood_size = args.setdefault("neighborhood_size", 3)
        half = hood_size // 2
        idx = np.arange(n)
        start = np.where(idx < half, n - half + idx, idx - half)
        hood = (start[:, None] + np.arange(hood_size)[None, :]) % n
        while not self._terminate_lazy(h.OPT_POP, fits, args):
            # ring topology: the first neighbour not beaten by a later one (ties -> earliest) = first arg-min in hood order
            if np.isnan(arch_f).any():
                nbest = np.array([_first_best(list(arch_f), list(hd)) for hd in hood])
            else:
                nbest = hood[idx, np.argmin(arch_f[hood], axis=1)]
            r = draw_uniforms(rnd, n * dims * 2)
            h.pso_step(r, nbest, improved, inertia, cognitive, social, lo, hi)
            fits = list(self._evaluate(h.OPT_POP, n))
            self.num_evaluations += n
            new_f = np.array(fits, dtype=float)
            # _swarm_archiver: the archived particle stays only when the new one is strictly worse
            better = ~(new_f > arch_f)
            arch_f = np.where(better, new_f, arch_f)
            improved = better.astype(np.int32)
            self.num_generations += 1
            self.history.append(min(fits))
        self._row_cache = {}
        self.population = self._individuals(h.OPT_POP, fits)
        for ind in self.population:
            ind.candidate  # noqa: B018
        return self.population
