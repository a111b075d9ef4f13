"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the inspyred pieces pyKriging drives
(``inspyred.ec.GA``, ``inspyred.swarm.PSO`` with ``ring_topology``, ``inspyred.ec.Bounder``;
call sites pyKriging/krige.py:276-291, 373-399 and regressionkrige.py:272-287, 370-396).

PARITY UNPINNED: inspyred is a third-party dependency that is neither vendored in /root/reference
nor version-pinned (setup.py:25 lists it bare) and is not installable here (no network).  This file
restates the published algorithm of inspyred 1.0.x from SURVEY.md App. B: one ``Individual`` object
per candidate, the framework's evolve loop (select -> copy -> variate -> evaluate -> replace ->
archive -> terminate), ``rank_selection``, ``n_point_crossover``, ``bit_flip_mutation``,
``generational_replacement``, the swarm variator / archiver / replacer and ``ring_topology``.
No reference test or golden vector exists at this boundary, so the product optimiser
(``pykriging_b200/optimizers.py``) is checked against THIS restatement under fixed seeds, and the
claim "same best theta/p under a fixed optimiser seed" means: same as this restatement driving the
CPU oracle likelihood.
"""
import copy
import functools


class Individual(object):
    def __init__(self, candidate=None, maximize=True):
        self.candidate = candidate
        self.fitness = None
        self.maximize = maximize

    def __lt__(self, other):
        if self.fitness is not None and other.fitness is not None:
            if self.maximize:
                return self.fitness < other.fitness
            return self.fitness > other.fitness
        raise ValueError("fitness is not defined")

    def __le__(self, other):
        return self < other or not other < self

    def __gt__(self, other):
        return other < self

    def __ge__(self, other):
        return other < self or not self < other


class Bounder(object):
    def __init__(self, lower_bound=None, upper_bound=None):
        self.lower_bound = lower_bound
        self.upper_bound = upper_bound

    def __call__(self, candidate, args):
        if self.lower_bound is None or self.upper_bound is None:
            return candidate
        bounded = candidate
        for i, (c, lo, hi) in enumerate(zip(candidate, self.lower_bound, self.upper_bound)):
            bounded[i] = max(min(c, hi), lo)
        return bounded


# ---- selectors / variators / replacers ---------------------------------------------------------------
def default_selection(random, population, args):
    return population


def rank_selection(random, population, args):
    num_selected = args.setdefault('num_selected', 1)
    len_pop = len(population)
    population.sort()
    psum = list(range(len_pop))
    den = (len_pop * (len_pop + 1)) / 2.0
    for i in range(len_pop):
        psum[i] = (i + 1) / den
    for i in range(1, len_pop):
        psum[i] += psum[i - 1]
    selected = []
    for _ in range(num_selected):
        cutoff = random.random()
        lower = 0
        upper = len_pop - 1
        while upper >= lower:
            mid = (lower + upper) // 2
            if psum[mid] > cutoff:
                upper = mid - 1
            else:
                lower = mid + 1
        lower = max(0, min(len_pop - 1, lower))
        selected.append(population[lower])
    return selected


def crossover(cross):
    @functools.wraps(cross)
    def inspyred_crossover(random, candidates, args):
        if len(candidates) % 2 == 1:
            candidates = candidates[:-1]
        moms = candidates[::2]
        dads = candidates[1::2]
        children = []
        for i, (mom, dad) in enumerate(zip(moms, dads)):
            cross.index = i
            offspring = cross(random, mom, dad, args)
            for o in offspring:
                children.append(o)
        return children
    return inspyred_crossover


@crossover
def n_point_crossover(random, mom, dad, args):
    crossover_rate = args.setdefault('crossover_rate', 1.0)
    num_crossover_points = args.setdefault('num_crossover_points', 1)
    children = []
    if random.random() < crossover_rate:
        num_cuts = min(len(mom) - 1, num_crossover_points)
        cut_points = random.sample(range(1, len(mom)), num_cuts)
        cut_points.sort()
        bro = copy.copy(dad)
        sis = copy.copy(mom)
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
    return children


def mutator(mutate):
    @functools.wraps(mutate)
    def inspyred_mutator(random, candidates, args):
        mutants = []
        for i, cs in enumerate(candidates):
            mutants.append(mutate(random, cs, args))
        return mutants
    return inspyred_mutator


@mutator
def bit_flip_mutation(random, candidate, args):
    rate = args.setdefault('mutation_rate', 0.1)
    mutant = copy.copy(candidate)
    if len(mutant) == len([x for x in mutant if x in [0, 1]]):
        for i, m in enumerate(mutant):
            if random.random() < rate:
                mutant[i] = (m + 1) % 2
    return mutant


def generational_replacement(random, population, parents, offspring, args):
    num_elites = args.setdefault('num_elites', 0)
    population.sort(reverse=True)
    offspring.extend(population[:num_elites])
    offspring.sort(reverse=True)
    survivors = offspring[:len(population)]
    return survivors


def default_archiver(random, population, archive, args):
    return archive


def ring_topology(random, population, args):
    neighborhood_size = args.setdefault('neighborhood_size', 3)
    half_hood = neighborhood_size // 2
    neighbor_index_start = []
    for index in range(len(population)):
        if index < half_hood:
            neighbor_index_start.append(len(population) - half_hood + index)
        else:
            neighbor_index_start.append(index - half_hood)
    for start in neighbor_index_start:
        n = []
        for i in range(neighborhood_size):
            n.append(population[(start + i) % len(population)])
        yield n


# ---- framework -----------------------------------------------------------------------------------------
class EvolutionaryComputation(object):
    def __init__(self, random):
        self.selector = default_selection
        self.variator = None
        self.replacer = None
        self.archiver = default_archiver
        self.terminator = None
        self._random = random
        self.trace = []  # (generation, [candidates], [fitness]) after every evaluation -- for tests

    def _should_terminate(self, pop, ng, ne):
        return bool(self.terminator(population=pop, num_generations=ng, num_evaluations=ne, args=self._kwargs))

    def evolve(self, generator, evaluator, pop_size=100, seeds=None, maximize=True, bounder=None, **args):
        self._kwargs = args
        self._kwargs['_ec'] = self
        if bounder is None:
            bounder = Bounder()
        self.bounder = bounder
        self.maximize = maximize
        self.population = []
        self.archive = []
        initial_cs = []
        for _ in range(pop_size):
            initial_cs.append(generator(random=self._random, args=self._kwargs))
        initial_fit = evaluator(candidates=initial_cs, args=self._kwargs)
        for cs, fit in zip(initial_cs, initial_fit):
            ind = Individual(cs, maximize=maximize)
            ind.fitness = fit
            self.population.append(ind)
        self.trace.append((0, copy.deepcopy(initial_cs), list(initial_fit)))
        self.num_evaluations = len(initial_fit)
        self.num_generations = 0
        self.archive = self.archiver(random=self._random, population=list(self.population),
                                     archive=list(self.archive), args=self._kwargs)
        while not self._should_terminate(list(self.population), self.num_generations, self.num_evaluations):
            parents = self.selector(random=self._random, population=list(self.population), args=self._kwargs)
            offspring_cs = [copy.deepcopy(i.candidate) for i in parents]
            variators = self.variator if isinstance(self.variator, (list, tuple)) else [self.variator]
            for op in variators:
                offspring_cs = op(random=self._random, candidates=offspring_cs, args=self._kwargs)
            offspring_fit = evaluator(candidates=offspring_cs, args=self._kwargs)
            offspring = []
            for cs, fit in zip(offspring_cs, offspring_fit):
                off = Individual(cs, maximize=maximize)
                off.fitness = fit
                offspring.append(off)
            self.num_evaluations += len(offspring_fit)
            self.trace.append((self.num_generations + 1, copy.deepcopy(offspring_cs), list(offspring_fit)))
            self.population = self.replacer(random=self._random, population=self.population, parents=parents,
                                            offspring=offspring, args=self._kwargs)
            self.archive = self.archiver(random=self._random, archive=self.archive,
                                         population=list(self.population), args=self._kwargs)
            self.num_generations += 1
        return self.population


class GA(EvolutionaryComputation):
    def __init__(self, random):
        EvolutionaryComputation.__init__(self, random)
        self.selector = rank_selection
        self.variator = [n_point_crossover, bit_flip_mutation]
        self.replacer = generational_replacement

    def evolve(self, generator, evaluator, pop_size=100, seeds=None, maximize=True, bounder=None, **args):
        args.setdefault('num_selected', pop_size)
        return EvolutionaryComputation.evolve(self, generator, evaluator, pop_size, seeds, maximize, bounder, **args)


class PSO(EvolutionaryComputation):
    def __init__(self, random):
        EvolutionaryComputation.__init__(self, random)
        self.topology = ring_topology
        self._previous_population = []
        self.selector = self._swarm_selector
        self.replacer = self._swarm_replacer
        self.variator = self._swarm_variator
        self.archiver = self._swarm_archiver

    def _swarm_archiver(self, random, population, archive, args):
        if len(archive) == 0:
            return population[:]
        new_archive = []
        for p, a in zip(population[:], archive[:]):
            if p < a:
                new_archive.append(a)
            else:
                new_archive.append(p)
        return new_archive

    def _swarm_variator(self, random, candidates, args):
        inertia = args.setdefault('inertia', 0.5)
        cognitive_rate = args.setdefault('cognitive_rate', 2.1)
        social_rate = args.setdefault('social_rate', 2.1)
        if len(self.archive) == 0:
            self.archive = self.population[:]
        if len(self._previous_population) == 0:
            self._previous_population = self.population[:]
        neighbors = self.topology(self._random, self.archive, args)
        offspring = []
        for x, xprev, pbest, hood in zip(self.population, self._previous_population, self.archive, neighbors):
            nbest = max(hood)
            particle = []
            for xi, xpi, pbi, nbi in zip(x.candidate, xprev.candidate, pbest.candidate, nbest.candidate):
                value = (xi + inertia * (xi - xpi) +
                         cognitive_rate * random.random() * (pbi - xi) +
                         social_rate * random.random() * (nbi - xi))
                particle.append(value)
            particle = self.bounder(particle, args)
            offspring.append(particle)
        return offspring

    def _swarm_selector(self, random, population, args):
        return population

    def _swarm_replacer(self, random, population, parents, offspring, args):
        self._previous_population = population[:]
        return offspring
