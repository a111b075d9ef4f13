"""TEST INFRASTRUCTURE ONLY -- CPU oracle for the pyKriging fit/predict hot path.

A NumPy restatement of the reference's algorithm (capaulson/pyKriging,
``pyKriging/matrixops.py`` + the hot methods of ``krige.py`` / ``regressionkrige.py``).
Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import it; the product package ``pykriging_b200`` never
does (it fails loudly when the CUDA library is missing).

Pinning: the reference holds NO tests or golden vectors of its own, so
this oracle is pinned against outputs of the *unmodified reference executed in the build
container* (``oracle/ref_loader.py`` + ``oracle/make_golden.py`` -> ``tests/golden/*.json``);
``tests/test_oracle.py`` checks (i) oracle == committed golden vectors everywhere and
(ii) oracle == live reference, bit for bit, when /root/reference is present.
The GA/PSO driver restatement (``oracle/inspyred_port.py``) has no such anchor:
inspyred is an un-vendored, un-pinned dependency -> "parity unpinned" for that part.

Two flavours are provided for every quantity:
  * ``*_ref``  -- follows the reference's NumPy expression order exactly (same ufuncs, same
                 LAPACK entry points: ``dpotrf`` through ``np.linalg.cholesky``, ``dgesv``
                 through ``np.linalg.solve``), so results are bit-identical to the reference
                 on the same BLAS.  This is also what the CPU baseline times.
  * ``*_fast`` -- an independent formulation (triangular solves, chunked Psi) used to check
                 big cases (n=4096, 1e5 query points) in seconds.
"""
import math

import numpy as np

EPS_DIAG = float(np.spacing(1))  # matrixops.py:30 adds eye*np.spacing(1)
PENALTY = 10000  # krige.py:447-450 (a Python int)


# --------------------------------------------------------------------------------------
# normalisation (krige.py:78-133, regressionkrige.py:75-125)
# --------------------------------------------------------------------------------------
def normalize_data(X, y):
    """Min-max normalisation exactly as ``normalizeData`` (krige.py:117-133).

    Returns (Xn, yn, normRange, ynormRange).  The arithmetic is the reference's:
    ``(x - lo) / float(hi - lo)`` element by element.
    """
    X = np.array(X, dtype=float, copy=True)
    y = np.array(y, dtype=float, copy=True)
    n, k = X.shape
    norm_range = [[min(X[:, d]), max(X[:, d])] for d in range(k)]
    for d in range(k):
        lo, hi = norm_range[d]
        X[:, d] = (X[:, d] - lo) / float(hi - lo)
    ylo, yhi = min(y), max(y)
    y = (y - ylo) / (yhi - ylo)
    return X, y, norm_range, [ylo, yhi]


def norm_x(x, norm_range):
    """krige.py:78-90 (1-D point in real-world units -> model units)."""
    x = np.array(x, dtype=float, copy=True)
    for d in range(len(norm_range)):
        lo, hi = norm_range[d]
        x[..., d] = (x[..., d] - lo) / float(hi - lo)
    return x


def inverse_norm_x(x, norm_range):
    """krige.py:92-101."""
    x = np.array(x, dtype=float, copy=True)
    for d in range(len(norm_range)):
        lo, hi = norm_range[d]
        x[..., d] = (x[..., d] * float(hi - lo)) + lo
    return x


def norm_y(y, ynorm_range):
    """krige.py:103-108."""
    return (y - ynorm_range[0]) / (ynorm_range[1] - ynorm_range[0])


def inverse_norm_y(y, ynorm_range):
    """krige.py:110-115."""
    return (y * (ynorm_range[1] - ynorm_range[0])) + ynorm_range[0]


# --------------------------------------------------------------------------------------
# distance tensor + Psi (matrixops.py:18-42)
# --------------------------------------------------------------------------------------
def distance_tensor(X):
    """matrixops.py:18-22: |X[i]-X[j]| for i<j, zeros elsewhere; shape (n,n,k).

    Vectorised (the reference uses an O(n^2) Python loop); values are identical because each
    entry is a single subtraction + abs.
    """
    X = np.asarray(X, dtype=float)
    n = X.shape[0]
    D = np.abs(X[:, None, :] - X[None, :, :])
    iu = np.triu(np.ones((n, n), dtype=bool), 1)
    D[~iu] = 0.0
    return D


def psi_ref(distance, theta, pl, diag_extra=EPS_DIAG):
    """Correlation matrix as the reference builds it (matrixops.py:24-30 / 34-40).

    ``diag_extra`` = np.spacing(1) for the interpolating model, ``Lambda`` for regression.
    Expression order follows the reference: power -> *theta -> sum(axis=2) -> negate -> exp
    -> triu(.,1) -> + transpose -> + I -> + I*extra.
    """
    theta = np.asarray(theta, dtype=float)
    pl = np.asarray(pl, dtype=float)
    n = distance.shape[0]
    new_psi = np.exp(-np.sum(theta * np.power(distance, pl), axis=2))
    upper = np.triu(new_psi, 1)
    eye = np.eye(n)
    return upper + upper.T + eye + eye * diag_extra


def chol_upper_ref(Psi):
    """matrixops.py:31-32: U = cholesky(Psi).T; raises np.linalg.LinAlgError when not PD."""
    return np.linalg.cholesky(Psi).T


def neglikelihood_ref(U, y):
    """matrixops.py:45-56 -- concentrated ln-likelihood with the reference's six dgesv solves.

    Returns (NegLnLike, mu, SigmaSqr, LnDetPsi).
    """
    n = U.shape[0]
    one = np.ones(n)
    y = np.asarray(y, dtype=float)
    ln_det = 2 * np.sum(np.log(np.abs(np.diag(U))))
    a = np.linalg.solve(U.T, one)
    b = np.linalg.solve(U, a)
    c = one.dot(b)
    d = np.linalg.solve(U.T, y)
    e = np.linalg.solve(U, d)
    mu = one.dot(e) / c
    r = y - one * mu
    sigma2 = r.dot(np.linalg.solve(U, np.linalg.solve(U.T, r))) / n
    with np.errstate(all="ignore"):
        nll = -1.0 * (-(n / 2.0) * np.log(sigma2) - 0.5 * ln_det)
    return nll, mu, sigma2, ln_det


def psi_fast(X, theta, pl, diag_extra=EPS_DIAG, chunk=256):
    """Independent Psi builder that never materialises the (n,n,k) tensor."""
    X = np.asarray(X, dtype=float)
    theta = np.asarray(theta, dtype=float)
    pl = np.asarray(pl, dtype=float)
    n, k = X.shape
    Psi = np.empty((n, n))
    for r0 in range(0, n, chunk):
        r1 = min(n, r0 + chunk)
        D = np.abs(X[r0:r1, None, :] - X[None, :, :])
        Psi[r0:r1] = np.exp(-np.sum(theta * np.power(D, pl), axis=2))
    Psi[np.diag_indices(n)] = 1.0 + diag_extra
    return Psi


def neglikelihood_fast(Psi, y):
    """Independent formulation: L = chol(Psi); a = L^-1 1, d = L^-1 y.

    Returns (nll, mu, sigma2, lndet, info) with info = 0 on success, >0 = LAPACK-style index of
    the first non-positive pivot.
    """
    from scipy.linalg import cho_factor, solve_triangular  # noqa: F401
    from scipy.linalg.lapack import dpotrf

    n = Psi.shape[0]
    L, info = dpotrf(Psi, lower=1, clean=1, overwrite_a=0)
    if info != 0:
        return float(PENALTY), float("nan"), float("nan"), float("nan"), int(info)
    y = np.asarray(y, dtype=float)
    rhs = np.stack([np.ones(n), y], axis=1)
    ad = solve_triangular(L, rhs, lower=True)
    a, d = ad[:, 0], ad[:, 1]
    mu = a.dot(d) / a.dot(a)
    r = d - mu * a
    sigma2 = r.dot(r) / n
    ln_det = 2.0 * np.sum(np.log(np.diag(L)))
    with np.errstate(all="ignore"):
        nll = (n / 2.0) * np.log(sigma2) + 0.5 * ln_det
    return nll, mu, sigma2, ln_det, 0


# --------------------------------------------------------------------------------------
# point predictors (matrixops.py:68-95, krige.py:201-234)
# --------------------------------------------------------------------------------------
def psi_vector(X, x, theta, pl):
    """matrixops.py:69-70: psi_i = exp(-sum(theta*|X_i-x|**pl)) (row-wise np.sum, as the loop)."""
    X = np.asarray(X, dtype=float)
    n = X.shape[0]
    psi = np.zeros((n, 1))
    for i in range(n):
        psi[i] = np.exp(-np.sum(theta * np.power(np.abs(X[i] - x), pl)))
    return psi


def predict_normalized_ref(X, y, U, mu, theta, pl, x):
    """matrixops.py:68-77."""
    n = X.shape[0]
    one = np.ones(n)
    psi = psi_vector(X, x, theta, pl)
    z = y - one * mu
    a = np.linalg.solve(U.T, z)
    b = np.linalg.solve(U, a)
    return (mu + psi.T.dot(b))[0]


def predicterr_normalized_ref(X, U, sigma2, theta, pl, x, extra=0.0):
    """matrixops.py:79-95 (extra=0) and the dead variant :97-110 (extra = Lambda)."""
    psi = psi_vector(X, x, theta, pl)
    ssqr = sigma2 * (1 + extra - psi.T.dot(np.linalg.solve(U, np.linalg.solve(U.T, psi))))
    ssqr = np.abs(ssqr[0])
    return np.power(ssqr, 0.5)[0]


def expimp_from(yhat, S, y_min, w=None, strict_else=True):
    """krige.py:207-217 / 227-233 given the two predictor outputs.

    ``strict_else``: kriging uses ``else`` (NaN S falls into the formula), the regression
    class uses ``elif S > 0`` and would leave EI undefined for NaN -- we return NaN there.
    """
    if S <= 0.0:
        return 0.0
    if not strict_else and not (S > 0.0):
        return float("nan")
    u = y_min - yhat
    one = u * (0.5 + 0.5 * math.erf((1.0 / np.sqrt(2.0)) * (u / S)))
    two = (S * (1.0 / np.sqrt(2.0 * np.pi))) * (np.exp(-(1.0 / 2.0) * (u ** 2.0 / S ** 2.0)))
    if w is None:
        return one + two
    return w * one + (1.0 - w) * two


def predict_batch_fast(X, y, theta, pl, diag_extra, Xq, y_min=None, w=None):
    """Vectorised independent predictor for many query points (model units).

    Returns dict(yhat, s, ei, mu, sigma2).  Uses L^-1 psi via triangular solves.
    """
    from scipy.linalg import solve_triangular

    X = np.asarray(X, dtype=float)
    y = np.asarray(y, dtype=float)
    Xq = np.atleast_2d(np.asarray(Xq, dtype=float))
    theta = np.asarray(theta, dtype=float)
    pl = np.asarray(pl, dtype=float)
    n = X.shape[0]
    Psi = psi_fast(X, theta, pl, diag_extra)
    L = np.linalg.cholesky(Psi)
    ad = solve_triangular(L, np.stack([np.ones(n), y], axis=1), lower=True)
    a, d = ad[:, 0], ad[:, 1]
    mu = a.dot(d) / a.dot(a)
    r = d - mu * a
    sigma2 = r.dot(r) / n
    wvec = solve_triangular(L, r, lower=True, trans="T")
    yhat = np.empty(len(Xq))
    s = np.empty(len(Xq))
    for q0 in range(0, len(Xq), 4096):
        q1 = min(len(Xq), q0 + 4096)
        D = np.abs(X[None, :, :] - Xq[q0:q1, None, :])
        P = np.exp(-np.sum(theta * np.power(D, pl), axis=2))  # (q, n)
        yhat[q0:q1] = mu + P.dot(wvec)
        V = solve_triangular(L, P.T, lower=True)
        s[q0:q1] = np.sqrt(np.abs(sigma2 * (1.0 - np.sum(V * V, axis=0))))
    out = {"yhat": yhat, "s": s, "mu": mu, "sigma2": sigma2}
    if y_min is not None:
        out["ei"] = np.array([expimp_from(yh, ss, y_min, w) for yh, ss in zip(yhat, s)])
    return out


# --------------------------------------------------------------------------------------
# stateful models mirroring the failure semantics
# --------------------------------------------------------------------------------------
class OracleKriging(object):
    """Mirror of ``kriging`` hot-path state machine (krige.py:23-76,158-258,429-472).

    X, y are in REAL-WORLD units (as handed to ``kriging(X, y)``); normalisation is applied
    here exactly as the reference does.  ``regression=True`` gives ``regression_kriging``
    semantics (Lambda, swallowed factorisation failures -> stale U).
    """

    def __init__(self, X, y, regression=False, normalize=True):
        if normalize:
            self.X, self.y, self.normRange, self.ynormRange = normalize_data(X, y)
        else:
            self.X = np.array(X, dtype=float, copy=True)
            self.y = np.array(y, dtype=float, copy=True)
            self.normRange = [[0.0, 1.0]] * self.X.shape[1]
            self.ynormRange = [0.0, 1.0]
        self.n, self.k = self.X.shape
        self.regression = bool(regression)
        self.theta = np.ones(self.k)
        self.pl = np.ones(self.k) * 2.0
        self.Lambda = 1  # matrixops.py:15 (also wipes regression's Lambda=0)
        self.U = None
        self.Psi = np.zeros((self.n, self.n))
        self.mu = None
        self.SigmaSqr = None
        self.NegLnLike = None
        self.LnDetPsi = None
        self.distance = distance_tensor(self.X)

    # -- matrixops.updatePsi / regupdatePsi wrapped by updateModel ------------------------
    def updateModel(self):
        extra = self.Lambda if self.regression else EPS_DIAG
        try:
            self.Psi = psi_ref(self.distance, self.theta, self.pl, extra)
            self.U = chol_upper_ref(self.Psi)
        except Exception:
            if self.regression:
                return  # regressionkrige.py:167-172 swallows
            raise Exception("bad params")  # krige.py:175-178

    def update(self, values):
        for i in range(self.k):
            self.theta[i] = values[i]
            self.pl[i] = values[i + self.k]
        if self.regression:
            self.Lambda = values[-1]
        self.updateModel()

    def neglikelihood(self):
        self.NegLnLike, self.mu, self.SigmaSqr, self.LnDetPsi = neglikelihood_ref(self.U, self.y)

    def fittingObjective(self, candidates):
        fitness = []
        for entry in candidates:
            f = PENALTY
            for i in range(self.k):
                self.theta[i] = entry[i]
                self.pl[i] = entry[i + self.k]
            if self.regression:
                self.Lambda = entry[-1]
            try:
                self.updateModel()
                self.neglikelihood()
                f = self.NegLnLike
            except Exception:
                f = PENALTY
            fitness.append(f)
        return fitness

    # -- krige.py:135-156 ----------------------------------------------------------------
    def addPoint(self, newX, newy, norm=True):
        """Append a sample, rebuild the distance tensor (updateData) and refactorise (updateModel).  mu / SigmaSqr
        are NOT refreshed, exactly as in the reference.  (The reference retrains when updateModel raises; the oracle
        lets the exception through -- the tests do not go there.)"""
        if norm:
            newX = norm_x(newX, self.normRange)
            newy = norm_y(newy, self.ynormRange)
        self.X = np.append(self.X, [newX], axis=0)
        self.y = np.append(self.y, newy)
        self.n = self.X.shape[0]
        self.distance = distance_tensor(self.X)
        self.updateModel()

    # -- krige.py:236-258 ----------------------------------------------------------------
    def infill_objective_mse(self, candidates, args=None):
        return [-1 * self.predicterr_normalized(np.asarray(entry, dtype=float)) for entry in candidates]

    def infill_objective_ei(self, candidates, args=None):
        return [-1 * self.expimp(np.asarray(entry, dtype=float)) for entry in candidates]

    # -- predictors -----------------------------------------------------------------------
    def predict_normalized(self, x):
        return predict_normalized_ref(self.X, self.y, self.U, self.mu, self.theta, self.pl, x)

    def predicterr_normalized(self, x):
        return predicterr_normalized_ref(self.X, self.U, self.SigmaSqr, self.theta, self.pl, x)

    def predict(self, x):
        return inverse_norm_y(self.predict_normalized(norm_x(x, self.normRange)), self.ynormRange)

    def predict_var(self, x):
        return self.predicterr_normalized(norm_x(x, self.normRange))

    def expimp(self, x, w=None):
        S = self.predicterr_normalized(x)
        y_min = np.min(self.y)
        if S <= 0.0:
            return 0.0
        return expimp_from(self.predict_normalized(x), S, y_min, w, strict_else=not self.regression)


# --------------------------------------------------------------------------------------
# synthetic inputs shared by tests and bench
# --------------------------------------------------------------------------------------
def rlh(n, k, seed):
    """Random Latin hypercube a la samplingplan.rlh (samplingplan.py:17-43, Edges=0)."""
    rng = np.random.default_rng(seed)
    X = np.zeros((n, k))
    for d in range(k):
        X[:, d] = rng.permutation(np.arange(1, n + 1, 1))
    return (X - 0.5) / n


def f_branin(X):
    """testfunctions.py:46-64."""
    X = np.atleast_2d(X)
    x, y = X[:, 0], X[:, 1]
    X1 = 15 * x - 5
    X2 = 15 * y
    b = 5.1 / (4 * np.pi ** 2)
    c = 5 / np.pi
    ff = 1 / (8 * np.pi)
    return ((X2 - b * X1 ** 2 + c * X1 - 6) ** 2 + 10 * (1 - ff) * np.cos(X1) + 10) + 5 * x


def f_squared(X, offset=0.25):
    """testfunctions.py:18-30."""
    X = np.atleast_2d(X)
    return np.sqrt(np.sum((X - offset) ** 2, axis=1))


def f_stybtang_norm(X):
    """testfunctions.py:144-160."""
    X = np.atleast_2d(X) * 10 - 5
    return np.sum(X ** 4 - 16 * X ** 2 + 5 * X, axis=1) / 2.0


def candidates_uniform(C, k, seed, regression=False):
    """What generate_population draws (krige.py:321-322) -- 'set U' of SURVEY 8d."""
    rng = np.random.default_rng(seed)
    if regression:
        th = rng.uniform(1e-4, 100.0, (C, k))
        p = rng.uniform(1.81231, 2.0, (C, k))
        lam = rng.uniform(0.0, 1.0, (C, 1))
        return np.concatenate([th, p, lam], axis=1)
    th = rng.uniform(1e-5, 100.0, (C, k))
    p = rng.uniform(1.0, 2.0, (C, k))
    return np.concatenate([th, p], axis=1)


def candidates_loguniform(C, k, seed, lo=-5.0, hi=2.0):
    """'set L' of SURVEY 8d: theta = 10**U[lo,hi], p ~ U[1,2] (late-generation-like)."""
    rng = np.random.default_rng(seed)
    th = 10.0 ** rng.uniform(lo, hi, (C, k))
    p = rng.uniform(1.0, 2.0, (C, k))
    return np.concatenate([th, p], axis=1)
