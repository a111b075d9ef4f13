"""``Cross_Validation`` -- leave-one-out / leave-q-out diagnostics of a kriging model
(pyKriging/CrossValidation.py:13-232), a consumer of the hot path.

Same class and method names and the same arithmetic as the reference: n independent (n-1)-point models,
each TRAINED from scratch (``train(optimiser)``), the left-out sample predicted with ``predict`` /
``predict_var``, the standardised cross-validated residual of Jones et al. (1998).  Every model is a
``pykriging_b200.kriging``: each generation of each training is one batched device call, each polish
gradient one stencil population.  Differences, all deliberate:

* ``predict_var`` returns a scalar here (the reference indexes the 1x1 ``np.matrix`` its ``np.linalg.solve``
  calls produce, ``CrossValidation.py:92`` -- ``[0, 0]``); the value is the same.
* the plotting branches (``plot=1``, ``QQ_plot``) are GUI code outside the hot path and raise.
* ``train_args`` (e.g. ``seed=``, ``pop_size=``) are forwarded to ``train`` so that a run can be reproduced;
  the reference's trainings are unseeded.  With a ``seed`` model i is trained with ``seed + i``.
"""
import copy
import random

import numpy as np

from .krige import kriging


def splitArrays(krigeModel, q=5):
    """pyKriging/utilities.py:17-30: q random folds of the model's samples -> (trainX, trainy, testX, testy)."""
    ind = np.arange(krigeModel.n)
    np.random.shuffle(ind)
    for fold in np.array_split(ind, q):
        X = copy.deepcopy(krigeModel.X)
        y = copy.deepcopy(krigeModel.y)
        yield np.delete(X, fold, axis=0), np.delete(y, fold, axis=0), X[fold], y[fold]


def mse(actual, predicted):
    return (actual - predicted) ** 2  # utilities.py:32-33


class Cross_Validation(object):
    model_class = kriging

    def __init__(self, model, name=None, **train_args):
        self.model = model
        self.X = self.model.X   # model units, as in the reference (CrossValidation.py:22-23)
        self.y = self.model.y
        self.n, self.k = np.shape(self.X)
        self.predict_list, self.predict_varr, self.scvr = [], [], []
        self.name = name
        self.train_args = dict(train_args)
        self.models = []

    def _train(self, model, optimiser, index=0):
        args = dict(self.train_args)
        args.pop("optimizer", None)
        if args.get("seed") is not None:
            args["seed"] = args["seed"] + index
        return model.train(optimizer=optimiser, **args)

    # -- CrossValidation.py:28-55 ---------------------------------------------------------------------------
    def calculate_RMSE_Rsquared(self, optimiser, nt):
        sample = random.sample(list(range(len(self.X))), nt)
        model = self.model_class(self.X, self.y, name='%s' % self.name)
        self._train(model, optimiser)
        yi_p = np.asarray(model.predict_batch(self.X[sample]), dtype=float)
        yi = np.asarray(self.y, dtype=float)[sample]
        # the reference's expression, kept literally (it squares the SUM of the residuals)
        RMSE = np.sqrt((sum(yi - yi_p) ** 2.) / float(nt))
        nt_f = float(nt)
        Rsquared = ((nt_f * sum(yi * yi_p) - sum(yi) * sum(yi_p)) /
                    (np.sqrt((nt_f * sum(yi * yi) - sum(yi) ** 2.) * (nt_f * sum(yi_p * yi_p) - sum(yi_p) ** 2.)))) ** 2.
        return ['RMSE = %f' % RMSE, 'Rsquared = %f' % Rsquared]

    # -- CrossValidation.py:57-123 / 125-196 ---------------------------------------------------------------
    def _scvr(self, y_, optimiser, plot):
        if plot not in (0, 1):
            raise ValueError('value for plot should be either 0 or 1')
        y_normalised = (y_ - np.min(y_)) / (np.max(y_) - np.min(y_))
        for i in range(self.n):
            idx = [j for j in range(self.n) if j != i]
            m = self.model_class(self.X[idx], y_[idx], name='%s' % self.name)
            self._train(m, optimiser, i)
            self.models.append(m)
            self.predict_list.append(m.predict(self.X[i]))
            self.predict_varr.append(m.predict_var(self.X[i]))
            self.scvr.append((y_normalised[i] - m.normy(self.predict_list[i])) / self.predict_varr[i])
        if plot == 1:
            raise NotImplementedError("the SCVR scatter plots are GUI code outside the hot path; "
                                      "predict_list / predict_varr / scvr hold the data")
        return self.predict_list, self.predict_varr, self.scvr

    def calculate_SCVR(self, optimiser='pso', plot=0):
        return self._scvr(np.copy(self.y), optimiser, plot)

    def calculate_transformed_SCVR(self, transformation, optimiser='pso', plot=0):
        y_ = np.copy(self.y)
        if transformation == 'logarithmic':
            y_ = np.log(y_)
        elif transformation == 'inverse':
            y_ = -(1.0 / y_)
        return self._scvr(y_, optimiser, plot)

    def QQ_plot(self):
        raise NotImplementedError("QQ_plot is GUI code outside the hot path; self.scvr holds the residuals")

    # -- CrossValidation.py:208-221 (the reference ignores q and always uses 5 folds; q is honoured here) ------
    def leave_n_out(self, q=5):
        mseArray = []
        for trainX, trainy, testX, testy in splitArrays(self.model, q):
            testk = self.model_class(trainX, trainy)
            self._train(testk, self.train_args.get("optimizer", 'pso'))
            pred = np.asarray(testk.predict_batch(testX), dtype=float)
            mseArray.extend(mse(np.asarray(testy, dtype=float), pred).tolist())
            del testk
        return np.average(mseArray), np.std(mseArray)
