"""``kriging`` -- the reference's interpolating model class (pyKriging/krige.py:22-728) on top of the
B200 ``matrixops`` mixin.

API kept: ``kriging(X, y, testfunction=None, name='', testPoints=None)``, ``train(optimizer)``,
``predict``, ``predict_var``, ``expimp``, ``weightedexpimp``, ``infill``, ``addPoint``, ``update``,
``updateModel``, ``fittingObjective``, ``fittingObjective_local``, ``infill_objective_mse/ei``,
``generate_population``, ``no_improvement_termination``, ``snapshot``, ``calcuatemeanMSE``,
``normX/inversenormX/normy/inversenormy/normalizeData``, attribute dict ``history``.

What changed relative to the reference:
  * the per-candidate Python loop of ``fittingObjective`` (krige.py:436-451) and the per-point loops of
    ``infill_objective_*`` (krige.py:243-257) are single batched device calls;
  * GA / PSO come from ``pykriging_b200.optimizers`` (inspyred semantics) and accept ``seed=``;
  * ``plot`` / ``saveFigure`` (matplotlib / mayavi GUI code, krige.py:474-674) are out of scope.
"""
import copy
from random import Random

import numpy as np
from scipy.optimize import minimize

from . import _lib
from .matrixops import matrixops
from .optimizers import GA, PSO, Bounder, DeviceGA, DevicePSO
from .samplingplan import samplingplan


def _scipy_has_workers():
    """SLSQP's ``workers`` map hook exists from SciPy 1.16 on; older versions would silently ignore it (an
    OptimizeWarning) and run the stencil sequentially -- same numbers, just slower, so fall back without it."""
    import scipy

    try:
        major, minor = (int(v) for v in scipy.__version__.split(".")[:2])
    except ValueError:
        return False
    return (major, minor) >= (1, 16)


class kriging(matrixops):
    _pk_mode = _lib.MODE_INTERP

    def __init__(self, X, y, testfunction=None, name='', testPoints=None, device=None, **kwargs):
        if device is not None:
            self._pk_device = int(device)
        self.X = copy.deepcopy(np.asarray(X, dtype=float))
        self.y = copy.deepcopy(np.asarray(y, dtype=float))
        self.testfunction = testfunction
        self.name = name
        self.n = self.X.shape[0]
        self.k = self.X.shape[1]
        self.theta = np.ones(self.k)
        self.pl = np.ones(self.k) * 2.
        self.sigma = 0
        self.normRange = []
        self.ynormRange = []
        self.normalizeData()
        self.sp = samplingplan(self.k)
        self._init_bounds()
        self._init_history(testPoints)
        matrixops.__init__(self)

    def _init_bounds(self):
        # krige.py:40-43
        self.thetamin = 1e-5
        self.thetamax = 100
        self.pmin = 1
        self.pmax = 2

    def _init_history(self, testPoints):
        # krige.py:45-73
        self.history = {}
        self.history['points'] = []
        self.history['neglnlike'] = []
        self.history['theta'] = []
        self.history['p'] = []
        self.history['rsquared'] = [0]
        self.history['adjrsquared'] = [0]
        self.history['chisquared'] = [1000]
        self.history['lastPredictedPoints'] = []
        self.history['avgMSE'] = []
        self.testPoints = None
        if testPoints:
            self.history['pointData'] = []
            self.testPoints = self.sp.rlh(testPoints)
            for point in self.testPoints:
                testPrimitive = {}
                testPrimitive['point'] = point
                if self.testfunction:
                    testPrimitive['actual'] = self.testfunction(point)[0]
                else:
                    testPrimitive['actual'] = None
                testPrimitive['predicted'] = []
                testPrimitive['mse'] = []
                testPrimitive['gradient'] = []
                self.history['pointData'].append(testPrimitive)
        else:
            self.history['pointData'] = None

    # -- normalisation: krige.py:78-133 (host glue; defines the model-unit inputs of the kernels) ------
    def normX(self, X):
        X = copy.deepcopy(X)
        if type(X) is np.float64:
            return np.array((X - self.normRange[0][0]) / float(self.normRange[0][1] - self.normRange[0][0]))
        else:
            for i in range(self.k):
                X[i] = (X[i] - self.normRange[i][0]) / float(self.normRange[i][1] - self.normRange[i][0])
        return X

    def inversenormX(self, X):
        X = copy.deepcopy(X)
        for i in range(self.k):
            X[i] = (X[i] * float(self.normRange[i][1] - self.normRange[i][0])) + self.normRange[i][0]
        return X

    def normy(self, y):
        return (y - self.ynormRange[0]) / (self.ynormRange[1] - self.ynormRange[0])

    def inversenormy(self, y):
        return (y * (self.ynormRange[1] - self.ynormRange[0])) + self.ynormRange[0]

    def normalizeData(self):
        for i in range(self.k):
            self.normRange.append([min(self.X[:, i]), max(self.X[:, i])])
        for i in range(self.n):
            self.X[i] = self.normX(self.X[i])
        self.ynormRange.append(min(self.y))
        self.ynormRange.append(max(self.y))
        for i in range(self.n):
            self.y[i] = self.normy(self.y[i])

    def _normX_batch(self, X):
        """normX for an (m, k) array: the same per-coordinate expression, vectorised over points."""
        X = np.array(X, dtype=float, copy=True)
        for i in range(self.k):
            X[:, i] = (X[:, i] - self.normRange[i][0]) / float(self.normRange[i][1] - self.normRange[i][0])
        return X

    # -- krige.py:135-178 -------------------------------------------------------------------------------
    incremental_addpoint = True  # bordered O(n^2) factor update when possible; False = always rebuild

    def addPoint(self, newX, newy, norm=True):
        if norm:
            newX = self.normX(newX)
            newy = self.normy(newy)
        self.X = np.append(self.X, [newX], axis=0)
        self.y = np.append(self.y, newy)
        self.n = self.X.shape[0]
        if self.incremental_addpoint and self._pk_try_append(newX, newy, self._pk_mode):
            return
        self.updateData()
        while True:
            try:
                self.updateModel()
            except Exception:
                self.train()
            else:
                break

    def _set_hyper(self, values):
        for i in range(self.k):
            self.theta[i] = values[i]
        for i in range(self.k):
            self.pl[i] = values[i + self.k]

    def update(self, values):
        self._set_hyper(values)
        self.updateModel()

    def updateModel(self):
        try:
            self.updatePsi()
        except Exception:
            raise Exception("bad params")

    # -- krige.py:180-258 -------------------------------------------------------------------------------
    def predict(self, X):
        X = copy.deepcopy(X)
        X = self.normX(X)
        return self.inversenormy(self.predict_normalized(X))

    def predict_var(self, X):
        X = copy.deepcopy(X)
        X = self.normX(X)
        return self.predicterr_normalized(X)

    def expimp(self, x):
        return np.float64(self.expimp_batch(np.atleast_2d(np.asarray(x, dtype=float)))[0])

    def weightedexpimp(self, x, w):
        return np.float64(self.expimp_batch(np.atleast_2d(np.asarray(x, dtype=float)), w=w)[0])

    def predict_batch(self, X):
        """predict() for an (m, k) array of real-world points (one device call)."""
        return self.inversenormy(self.predict_normalized_batch(self._normX_batch(X)))

    def predict_var_batch(self, X):
        return self.predicterr_normalized_batch(self._normX_batch(X))

    def infill_objective_mse(self, candidates, args):
        s = self.predicterr_normalized_batch(np.asarray(candidates, dtype=float))
        return [-1 * v for v in s]

    def infill_objective_ei(self, candidates, args):
        ei = self.expimp_batch(np.asarray(candidates, dtype=float))
        return [-1 * v for v in ei]

    # -- krige.py:260-309 -------------------------------------------------------------------------------
    def infill(self, points, method='error', addPoint=True, seed=None):
        initX = np.copy(self.X)
        inity = np.copy(self.y)
        returnValues = np.zeros([points, self.k], dtype=float)
        for i in range(points):
            ea = PSO(Random(seed) if seed is not None else Random())
            ea.terminator = self.no_improvement_termination
            if method == 'ei':
                evaluator = self.infill_objective_ei
            else:
                evaluator = self.infill_objective_mse
            final_pop = ea.evolve(generator=self.generate_population,
                                  evaluator=evaluator,
                                  pop_size=155,
                                  maximize=False,
                                  bounder=Bounder([0] * self.k, [1] * self.k),
                                  max_evaluations=20000,
                                  neighborhood_size=30,
                                  num_inputs=self.k)
            final_pop.sort(reverse=True)
            newpoint = final_pop[0].candidate
            returnValues[i][:] = self._infill_result(newpoint)
            if addPoint:
                self.addPoint(returnValues[i], self.predict(returnValues[i]), norm=True)
            if seed is not None:
                seed += 1
        self.X = np.copy(initX)
        self.y = np.copy(inity)
        self.n = len(self.X)
        self.updateData()
        while True:
            try:
                self.updateModel()
            except Exception:
                self.train()
            else:
                break
        return returnValues

    def _infill_result(self, newpoint):
        return self.inversenormX(np.array(newpoint, dtype=float))  # krige.py:294

    # -- krige.py:311-353 -------------------------------------------------------------------------------
    def generate_population(self, random, args):
        bounder = args["_ec"].bounder
        chromosome = []
        for lo, hi in zip(bounder.lower_bound, bounder.upper_bound):
            chromosome.append(random.uniform(lo, hi))
        return chromosome

    def no_improvement_termination(self, population, num_generations, num_evaluations, args):
        max_generations = args.setdefault('max_generations', 10)
        previous_best = args.setdefault('previous_best', None)
        max_evaluations = args.setdefault('max_evaluations', 30000)
        current_best = np.around(max(population).fitness, decimals=4)
        if previous_best is None or previous_best != current_best:
            args['previous_best'] = current_best
            args['generation_count'] = 0
            return False or (num_evaluations >= max_evaluations)
        else:
            if args['generation_count'] >= max_generations:
                return True
            else:
                args['generation_count'] += 1
                return False or (num_evaluations >= max_evaluations)

    # -- krige.py:355-427 -------------------------------------------------------------------------------
    _train_pop = {'pso': 300, 'ga': 300}
    device_optimizer = True  # GA / PSO population resident on the GPU (False: host arrays, one upload per generation)

    def _bounds(self):
        lowerBound = [self.thetamin] * self.k + [self.pmin] * self.k
        upperBound = [self.thetamax] * self.k + [self.pmax] * self.k
        return lowerBound, upperBound

    def train(self, optimizer='pso', seed=None, pop_size=None, max_evaluations=30000):
        """Global GA/PSO search + SLSQP polish, as krige.py:355-427.  ``seed`` makes the run
        reproducible (the reference always uses an unseeded ``Random()``)."""
        self.updateData()
        lowerBound, upperBound = self._bounds()
        rnd = Random(seed) if seed is not None else Random()
        # population resident on the GPU (pk_ga_step / pk_pso_step) unless switched off or the population is sharded
        # over ranks (every rank then evolves the replicated host population and evaluates its slice)
        on_device = self.device_optimizer and self.__dict__.get("_pk_shard") is None
        if optimizer == 'pso':
            ea = DevicePSO(rnd, self._pk_handle(), self._pk_device_objective) if on_device else PSO(rnd)
            ea.terminator = self.no_improvement_termination
            final_pop = ea.evolve(generator=self.generate_population,
                                  evaluator=self.fittingObjective,
                                  pop_size=pop_size or self._train_pop['pso'],
                                  maximize=False,
                                  bounder=Bounder(lowerBound, upperBound),
                                  max_evaluations=max_evaluations,
                                  neighborhood_size=20,
                                  num_inputs=self.k)
            final_pop.sort(reverse=True)
        elif optimizer == 'ga':
            ea = DeviceGA(rnd, self._pk_handle(), self._pk_device_objective) if on_device else GA(rnd)
            ea.terminator = self.no_improvement_termination
            final_pop = ea.evolve(generator=self.generate_population,
                                  evaluator=self.fittingObjective,
                                  pop_size=pop_size or self._train_pop['ga'],
                                  maximize=False,
                                  bounder=Bounder(lowerBound, upperBound),
                                  max_evaluations=max_evaluations,
                                  num_elites=10,
                                  mutation_rate=.05)
        else:
            raise UnboundLocalError("local variable 'final_pop' referenced before assignment")
        self._last_ea = ea
        locOP_bounds = [[lo, hi] for lo, hi in zip(lowerBound, upperBound)]
        for entry in final_pop:
            newValues = entry.candidate
            lopResults = self._pk_slsqp(newValues, locOP_bounds)
            newValues = lopResults['x']
            self._set_hyper_all(newValues)
            try:
                self.updateModel()
            except Exception:
                pass
            else:
                break

    # -- SLSQP polish (krige.py:402-427): the finite-difference stencil as ONE population ----------------
    slsqp_batched = True  # class-level switch; False = SciPy's own sequential differences (same numbers)

    def _pk_slsqp(self, start, bounds):
        """``minimize(self.fittingObjective_local, start, method='SLSQP', bounds=...)`` exactly as the
        reference calls it (krige.py:413).  SciPy approximates the gradient by forward differences: 2k[+1]
        single-candidate likelihoods per iterate, evaluated one after the other in the reference.  Here the
        SAME stencil points SciPy generates are handed to the device as one ``fittingObjective`` population
        through SciPy's ``workers`` map hook, so the iterates are those of the sequential run."""
        options = {'disp': False}
        if self.slsqp_batched and _scipy_has_workers():
            options['workers'] = self._pk_stencil_map
        return minimize(self.fittingObjective_local, start, method='SLSQP', bounds=bounds, options=options)

    def _pk_stencil_map(self, fun, points):
        """Map-like for SciPy's numerical differentiation (``workers(fun, iterable)``): every stencil point
        of one gradient in a single batched device call instead of ``[fun(x) for x in points]``."""
        pts = [np.asarray(x, dtype=float).tolist() for x in points]
        if not pts:
            return []
        self._pk_stencil_calls = getattr(self, "_pk_stencil_calls", 0) + 1
        return [np.atleast_1d(np.float64(v)) for v in self.fittingObjective(pts, {})]

    def _set_hyper_all(self, values):
        self._set_hyper(values)

    # -- krige.py:429-472: the batched entry point ---------------------------------------------------------
    def fittingObjective(self, candidates, args):
        """NegLnLike of every candidate [theta_1..k, p_1..k]: ONE batched device call.
        Cholesky failure -> the reference's penalty, the Python int 10000 (krige.py:447-450)."""
        if len(candidates) == 0:
            return []
        cands, get_cand = self._pk_candidate_rows(candidates)
        k = self.k
        out = self._pk_nll_population(np.ascontiguousarray(cands[:, :k]), np.ascontiguousarray(cands[:, k:2 * k]),
                                      None, _lib.MODE_INTERP)
        return self._pk_fitness_poststate(out, get_cand)

    def _pk_fitness_poststate(self, out, get_cand):
        """Fitness list + post-state of the reference's candidate loop from the per-candidate scalars of one batched
        evaluation.  ``get_cand(i)`` returns candidate i (a host row, or a row read back from the device-resident
        population): theta/pl = LAST candidate; mu/SigmaSqr/NegLnLike/U = last successfully evaluated one."""
        k = self.k
        info = out["info"]
        fitness = list(out["nll"])               # np.float64 scalars, as self.NegLnLike is in the reference
        bad = np.nonzero(info != 0)[0]
        for i in bad:
            fitness[i] = 10000                   # krige.py:447-450: the Python int
        ok = np.nonzero(info == 0)[0]
        last = get_cand(len(fitness) - 1)
        self._set_hyper(last)
        if len(ok):
            j = int(ok[-1])
            c = last if j == len(fitness) - 1 else get_cand(j)
            vals = (out["nll"][j], out["mu"][j], out["sigma2"][j], out["lndet"][j])
            self._pk_record_batch_state(vals, (np.array(c[:k], dtype=float), np.array(c[k:2 * k], dtype=float), 0.0),
                                        _lib.MODE_INTERP)
        else:
            self._pk_record_batch_state(None, None, _lib.MODE_INTERP)
        return fitness

    def _pk_device_objective(self, which, count):
        """The evaluator of a device-resident GA / PSO generation (optimizers.DeviceGA / DevicePSO): the candidates are
        rows of a GPU buffer, only the per-candidate scalars come back; same fitness list and post-state as
        ``fittingObjective``."""
        h = self._pk_handle()
        out = h.opt_evaluate(which, count, self._pk_mode)
        dims = 2 * self.k + (1 if self._pk_mode == _lib.MODE_REGRESSION else 0)
        return self._pk_fitness_poststate(out, lambda i: h.opt_get(which, i, 1, dims)[0])

    def fittingObjective_local(self, entry):
        """krige.py:454-472.  A population of one through the same device path as the batched objective, so
        that SLSQP's f(x) and its (batched) stencil values f(x + h e_i) come from the same kernels."""
        return self.fittingObjective([np.asarray(entry, dtype=float).tolist()], {})[0]

    # -- krige.py:676-728 (bookkeeping; consumers of the batched predictors) -------------------------------
    def calcuatemeanMSE(self, p2s=200, points=None):
        if points is None:
            points = self.sp.rlh(p2s)
        values = self.predict_var_batch(np.asarray(points, dtype=float))
        return np.mean(values), np.std(values)

    def snapshot(self):
        self.history['points'].append(self.n)
        self.history['neglnlike'].append(self.NegLnLike)
        self.history['theta'].append(copy.deepcopy(self.theta))
        self.history['p'].append(copy.deepcopy(self.pl))
        self.history['avgMSE'].append(self.calcuatemeanMSE(points=self.testPoints)[0])
        currentPredictions = []
        if self.history['pointData'] is not None:
            pts = np.array([pp['point'] for pp in self.history['pointData']], dtype=float)
            pred = self.predict_batch(pts)
            mse = self.predict_var_batch(pts)
            for pointprim, predictedPoint, m in zip(self.history['pointData'], pred, mse):
                currentPredictions.append(copy.deepcopy(predictedPoint))
                pointprim['predicted'].append(predictedPoint)
                pointprim['mse'].append(m)
                try:
                    pointprim['gradient'] = np.gradient(pointprim['predicted'])
                except Exception:
                    pass
        if self.history['lastPredictedPoints'] != []:
            self.history['chisquared'].append(self.chisquared(self.history['lastPredictedPoints'], currentPredictions))
            self.history['rsquared'].append(self.rsquared(self.history['lastPredictedPoints'], currentPredictions))
            self.history['adjrsquared'].append(self.adjrsquares(self.history['rsquared'][-1],
                                                                len(self.history['pointData'])))
        self.history['lastPredictedPoints'] = copy.deepcopy(currentPredictions)

    def rsquared(self, actual, observed):
        return np.corrcoef(observed, actual)[0, 1] ** 2

    def adjrsquares(self, rsquared, obs):
        return 1 - (1 - rsquared) * ((obs - 1) / (obs - self.k))

    def chisquared(self, actual, observed):
        actual = np.array(actual)
        observed = np.array(observed)
        return np.sum(np.abs(np.power((observed - actual), 2) / actual))

    def plot(self, *a, **kw):
        raise NotImplementedError("plot()/saveFigure() are GUI code outside the hot path (DESIGN.md, out of scope); "
                                  "use predict_batch()/predict_var_batch() to fill a grid")

    saveFigure = plot
