"""``regression_kriging`` -- the reference's nugget-regularised model class
(pyKriging/regressionkrige.py:20-725) on top of the B200 ``matrixops`` mixin.

It mirrors ``kriging`` 1:1 with one extra hyper-parameter ``Lambda`` (candidate layout
``[theta_1..k, p_1..k, lambda]``), the reference's different bounds and population sizes, and its
failure semantics: ``updateModel`` swallows a failed factorisation (regressionkrige.py:167-172), so
``regneglikelihood`` is evaluated on the STALE factor of the last successful candidate and repeats
that candidate's likelihood.
"""
import copy

import numpy as np

from . import _lib
from .krige import kriging


class regression_kriging(kriging):
    _pk_mode = _lib.MODE_REGRESSION
    _train_pop = {'pso': 150, 'ga': 50}  # regressionkrige.py:376, 391

    def _init_bounds(self):
        # regressionkrige.py:38-41
        self.thetamin = 1e-4
        self.thetamax = 100
        self.pmin = 1.81231
        self.pmax = 2

    # regressionkrige.py:75-82: no k == 1 np.float64 special case
    def normX(self, X):
        X = copy.deepcopy(X)
        for i in range(self.k):
            X[i] = (X[i] - self.normRange[i][0]) / float(self.normRange[i][1] - self.normRange[i][0])
        return X

    def _bounds(self):
        lowerBound = [self.thetamin] * self.k + [self.pmin] * self.k + [0]
        upperBound = [self.thetamax] * self.k + [self.pmax] * self.k + [1]
        return lowerBound, upperBound

    def _set_hyper_all(self, values):
        self._set_hyper(values)
        self.Lambda = values[-1]

    def update(self, values):
        self._set_hyper(values)
        self.Lambda = values[-1]
        self.updateModel()

    def updateModel(self):
        try:
            self.regupdatePsi()
        except Exception:
            pass  # regressionkrige.py:169-170

    def predict(self, X, norm=True):
        X = copy.deepcopy(X)
        if norm:
            X = self.normX(X)
        return self.inversenormy(self.predict_normalized(X))

    def predict_var(self, X, norm=True):
        X = copy.deepcopy(X)
        if norm:
            X = self.normX(X)
        return self.predicterr_normalized(X)

    def _infill_result(self, newpoint):
        return np.array(newpoint, dtype=float)  # regressionkrige.py:290 (no inversenormX)

    def fittingObjective(self, candidates, args):
        """NegLnLike of every candidate [theta, p, lambda] in ONE batched device call, followed by the
        host post-pass that reproduces the reference's stale-factor semantics."""
        if len(candidates) == 0:
            return []
        cands, get_cand = self._pk_candidate_rows(candidates)
        k = self.k
        out = self._pk_nll_population(np.ascontiguousarray(cands[:, :k]), np.ascontiguousarray(cands[:, k:2 * k]),
                                      np.ascontiguousarray(cands[:, -1]), _lib.MODE_REGRESSION)
        return self._pk_fitness_poststate(out, get_cand)

    def _pk_fitness_poststate(self, out, get_cand):
        """regressionkrige.py:163-172, 428-452: a failed factorisation is swallowed, so the candidate repeats the
        likelihood of the most recent successful one (possibly from before this call); the literal 10000 appears only
        while no factorisation has succeeded on this instance at all."""
        k = self.k
        info, nll = out["info"], out["nll"]
        C = len(nll)
        okmask = info == 0
        prev = np.maximum.accumulate(np.where(okmask, np.arange(C), -1))   # index of the last success at or before i
        before = np.float64(self._fit_vals[0]) if self._has_U else None    # factor present before this call?
        fitness = [None] * C
        vals = list(nll)
        for i in range(C):
            j = prev[i]
            fitness[i] = vals[j] if j >= 0 else (10000 if before is None else before)
        last = get_cand(C - 1)
        self._set_hyper(last)
        self.Lambda = last[-1]
        if okmask.any():
            j = int(prev[-1])
            c = last if j == C - 1 else get_cand(j)
            v = (out["nll"][j], out["mu"][j], out["sigma2"][j], out["lndet"][j])
            self._pk_record_batch_state(v, (np.array(c[:k], dtype=float), np.array(c[k:2 * k], dtype=float), c[-1]),
                                        _lib.MODE_REGRESSION)
        else:
            self._pk_record_batch_state(None, None, _lib.MODE_REGRESSION)
            if self._has_U:  # regneglikelihood re-evaluated on the stale factor refreshes these
                self.neglikelihood()
        return fitness

    def fittingObjective_local(self, entry):
        """regressionkrige.py:454-473 (stale-factor rule included: fittingObjective applies it)."""
        return self.fittingObjective([np.asarray(entry, dtype=float).tolist()], {})[0]
