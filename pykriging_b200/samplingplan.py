"""Minimal ``samplingplan`` (reference: pyKriging/samplingplan.py).  Only the random Latin hypercube
``rlh`` (samplingplan.py:17-43) is provided: it is what the model classes themselves call (test
points, calcuatemeanMSE).  The Morris-Mitchell optimiser (``optimallhc``) is input generation
outside the hot path."""
import numpy as np


class samplingplan(object):
    def __init__(self, k=2):
        self.samplingplan = []
        self.k = k

    def rlh(self, n, Edges=0):
        X = np.zeros((n, self.k))
        for i in range(0, self.k):
            X[:, i] = np.transpose(np.random.permutation(np.arange(1, n + 1, 1)))
        if Edges == 1:
            X = (X - 1) / (n - 1)
        else:
            X = (X - 0.5) / n
        return X
