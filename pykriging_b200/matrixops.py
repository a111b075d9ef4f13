"""Drop-in replacement of ``pyKriging.matrixops.matrixops`` (reference: pyKriging/matrixops.py:1-110).

Same class name, same method names, same attribute names -- ``kriging`` / ``regression_kriging``
inherit from it exactly as they do in the reference (krige.py:22, regressionkrige.py:20) -- but every
method is a thin call into the C ABI (``include/pykriging_b200.h``); no linear algebra happens in
Python and there is no CPU fallback.

State machine kept from the reference:
  * ``updatePsi`` / ``regupdatePsi`` overwrite ``Psi`` always, ``U`` only when the factorisation
    succeeds; on failure they raise ``numpy.linalg.LinAlgError`` like ``np.linalg.cholesky`` does.
  * ``neglikelihood`` / ``regneglikelihood`` use whatever factor is current (possibly stale).
  * the predictors use the CURRENT ``theta`` / ``pl`` for the psi vector and the CURRENT (possibly
    stale) ``mu`` / ``SigmaSqr``.
``Psi`` and ``U`` live on the device and are copied to host arrays only on attribute access.
The reference's ``distance`` tensor (n x n x k, 84 MB at n=1024, k=10) is never materialised.
"""
import numpy as np

from . import _lib, sharding


class matrixops(object):
    _pk_device = 0  # class-level default GPU; override per instance with `device=` or set_device()

    def __init__(self):
        self.LnDetPsi = None
        self.psi = np.zeros((self.n, 1))
        self.one = np.ones(self.n)
        self.mu = None
        self.SigmaSqr = None
        self.Lambda = 1
        self._pk_reset_factor()
        self.updateData()

    # -- device plumbing ------------------------------------------------------------------------
    def _pk_handle(self):
        h = self.__dict__.get("_pk")
        if h is None:
            h = _lib.Handle(self.__dict__.get("_pk_device", type(self)._pk_device))
            self.__dict__["_pk"] = h
        return h

    def set_device(self, device):
        """Move the model to another GPU of this process (drops device state)."""
        old = self.__dict__.pop("_pk", None)
        if old is not None:
            old.close()
        self.__dict__["_pk_device"] = int(device)
        self._pk_reset_factor()
        self.updateData()

    # -- multi-GPU: one process per GPU, candidates / query points sharded over ranks ------------------
    def enable_sharding(self, group=None, min_points=4096):
        """Shard every batched ``fittingObjective`` call (and batched predictions of at least
        ``min_points`` points) over the ranks of ``group`` (default process group).  Every rank must
        hold the same model and call the same methods with the same arguments; each evaluates its
        slice on its own GPU and one all_gather of scalars returns the full result everywhere."""
        self.__dict__["_pk_shard"] = (group, int(min_points))

    def disable_sharding(self):
        self.__dict__.pop("_pk_shard", None)
        sharding.release_buffers()

    def _pk_candidate_rows(self, candidates):
        """A population handed over as a list of lists -> (C x width float array, getter of row i).  With sharding enabled
        only THIS rank's slice of the list is converted (the other rows of the array stay uninitialised: the sharded
        evaluation reads nothing but its slice, sharding.shard_bounds) -- at 8 ranks the list -> array conversion of a
        1024 x 20 population, 0.6 ms in CPython, would otherwise be a quarter of a 2.1 ms generation on every rank."""
        shard = self.__dict__.get("_pk_shard")
        if shard is not None and isinstance(candidates, (list, tuple)) and len(candidates) >= 64:
            rank, world = sharding.rank_and_world(shard[0])
            if world > 1:
                C = len(candidates)
                lo, hi = sharding.shard_bounds(C, rank, world)
                mine = np.asarray(candidates[lo:hi], dtype=float)
                if mine.ndim == 2:
                    rows = np.empty((C, mine.shape[1]))
                    rows[lo:hi] = mine
                    return rows, (lambda i: rows[i] if lo <= i < hi else np.asarray(candidates[i], dtype=float))
        rows = np.asarray(candidates, dtype=float)
        return rows, (lambda i: rows[i])

    def _pk_nll_population(self, theta, p, lam, mode):
        """NegLnLike / mu / SigmaSqr / LnDetPsi / info of a population (dict of arrays), sharded if enabled."""
        h = self._pk_handle()
        shard = self.__dict__.get("_pk_shard")
        if shard is None:
            return h.nll_batch(theta, p, lam, mode)

        def eval_slice(lo, hi):
            h.set_shard_offset(lo)
            o = h.nll_batch(theta[lo:hi], p[lo:hi], None if lam is None else lam[lo:hi], mode)
            return np.stack([o["nll"], o["mu"], o["sigma2"], o["lndet"], o["info"].astype(np.float64)], axis=1)

        h.set_shard_hint(population=theta.shape[0])
        try:
            if sharding.uses_nccl(shard[0]):
                rows = sharding.sharded_nll_device(h, theta, p, lam, mode, shard[0])
            else:
                rows = sharding.sharded_rows(theta.shape[0], eval_slice, shard[0], ncols=5)
        finally:
            h.set_shard_hint()
        return {"nll": rows[:, 0].copy(), "mu": rows[:, 1].copy(), "sigma2": rows[:, 2].copy(),
                "lndet": rows[:, 3].copy(), "info": rows[:, 4].astype(np.int32)}

    def _pk_reset_factor(self):
        self._has_U = False          # reference: self.U = None
        self._fit_vals = None        # (NegLnLike, mu, SigmaSqr, LnDetPsi) of the current factor
        self._psi_host = None
        self._U_host = None
        self._psi_ever = False
        self._pending_factor = None  # (theta, pl, Lambda, mode): factor to restore lazily
        self._factor_params = None   # (theta, pl, Lambda, mode) of the factor cached on the device

    def _pk_ensure_factor(self):
        """After a batched fittingObjective the reference's self.U is the factor of the last
        successfully evaluated candidate.  The batched call already recorded its scalars
        (``_fit_vals``); the factor itself is re-created on the device only if somebody needs it
        (U access or a prediction)."""
        pend = self._pending_factor
        if pend is not None:
            self._pending_factor = None
            th, pl, lam, mode = pend
            h = self._pk_handle()
            _, info = h.fit(th, pl, lam, mode)
            if info != 0:
                # The batch factorised this candidate; pk_fit runs other kernels (gangs, the L^-T rows) and a pivot on
                # the knife edge of the relative test can come out on the other side.  Retry with LAPACK's literal rule
                # (pivot <= 0); if even that fails, say so instead of predicting with an older factor.
                tol = h.pivot_tol
                h.set_pivot_tolerance(0.0)
                try:
                    _, info = h.fit(th, pl, lam, mode)
                finally:
                    h.set_pivot_tolerance(tol)
                if info != 0:
                    self._factor_params = None
                    raise _lib.PkError("the factor of the last successfully evaluated candidate could not be re-created "
                                       "(info = %d): Psi is numerically singular" % info)
            self._factor_params = (th.copy(), pl.copy(), float(lam), mode)
            self._U_host = None
            self._psi_host = None

    def _pk_record_batch_state(self, last_ok_vals, last_ok_params, mode):
        """Post-state of the reference's candidate loop."""
        self._psi_ever = True
        self._psi_host = None
        if last_ok_vals is not None:
            nll, mu, s2, lndet = last_ok_vals
            self.LnDetPsi = np.float64(lndet)
            self.mu = np.float64(mu)
            self.SigmaSqr = np.float64(s2)
            self.NegLnLike = np.float64(nll)
            self._fit_vals = np.array([nll, mu, s2, lndet], dtype=float)
            self._has_U = True
            self._U_host = None
            th, pl, lam = last_ok_params
            self._pending_factor = (np.array(th, dtype=float), np.array(pl, dtype=float), float(lam), mode)

    # -- reference attributes that live on the device ---------------------------------------------
    @property
    def Psi(self):
        if not self._psi_ever:
            return np.zeros((self.n, self.n), dtype=float)  # matrixops.py:9
        if self._psi_host is None:
            self._psi_host = self._pk_handle().get_psi()
        return self._psi_host

    @Psi.setter
    def Psi(self, value):  # the reference assigns zeros in __init__/updatePsi; ignore
        pass

    @property
    def U(self):
        self._pk_ensure_factor()
        if not self._has_U:
            return None  # matrixops.py:13
        if self._U_host is None:
            self._U_host = self._pk_handle().get_U()
        return self._U_host

    @U.setter
    def U(self, value):
        if value is None:
            self._has_U = False

    # -- matrixops.py:18-22 -----------------------------------------------------------------------
    def updateData(self):
        """Upload X, y (O(nk) bytes); replaces the O(n^2 k) distance tensor build."""
        self._pk_handle().set_data(np.ascontiguousarray(self.X, dtype=float),
                                   np.ascontiguousarray(self.y, dtype=float))
        self._pending_factor = None
        self._factor_params = None  # the device drops its cached factor with the old data

    # -- matrixops.py:24-32 / 34-42 ---------------------------------------------------------------
    def _pk_factorise(self, mode):
        self.one = np.ones(self.n)
        self.psi = np.zeros((self.n, 1))
        self._pending_factor = None
        lam = float(self.Lambda) if mode == _lib.MODE_REGRESSION else 0.0
        out4, info = self._pk_handle().fit(np.asarray(self.theta, dtype=float), np.asarray(self.pl, dtype=float),
                                           lam, mode)
        self._psi_host = None
        self._psi_ever = True
        if info != 0:
            raise np.linalg.LinAlgError("Matrix is not positive definite")
        self._has_U = True
        self._U_host = None
        self._fit_vals = out4
        self._factor_params = (np.array(self.theta, dtype=float), np.array(self.pl, dtype=float), lam, mode)

    def _pk_try_append(self, newX, newy, mode):
        """addPoint without the rebuild: when the device holds the factor of the
        CURRENT theta / pl / Lambda for the data before the append, border it with the new sample (O(n^2))
        instead of updateData + updatePsi (O(n^2 k) pow + n^3/3 flop).  True = done: the state is what
        updateData(); updatePsi() would have left (U current, mu / SigmaSqr still those of the old model, as in
        the reference).  False = nothing happened, take the reference's route."""
        fp = self._factor_params
        if fp is None or not self._has_U or self._pending_factor is not None:
            return False
        lam = float(self.Lambda) if mode == _lib.MODE_REGRESSION else 0.0
        if fp[3] != mode or fp[2] != lam or not np.array_equal(fp[0], np.asarray(self.theta, dtype=float)) \
                or not np.array_equal(fp[1], np.asarray(self.pl, dtype=float)):
            return False
        appended, out4, _ = self._pk_handle().append_point(np.asarray(newX, dtype=float), float(newy))
        if not appended:
            return False
        self.one = np.ones(self.n)
        self.psi = np.zeros((self.n, 1))
        self._psi_host = None
        self._U_host = None
        self._fit_vals = out4
        self.__dict__["_pk_appends"] = self.__dict__.get("_pk_appends", 0) + 1
        return True

    def updatePsi(self):
        self._pk_factorise(_lib.MODE_INTERP)

    def regupdatePsi(self):
        self._pk_factorise(_lib.MODE_REGRESSION)

    # -- matrixops.py:45-56 / 58-66 ---------------------------------------------------------------
    def neglikelihood(self):
        if not self._has_U:
            raise AttributeError("'NoneType' object has no attribute 'T'")  # self.U is None (matrixops.py:48)
        nll, mu, s2, lndet = self._fit_vals
        self.LnDetPsi = np.float64(lndet)
        self.mu = np.float64(mu)
        self.SigmaSqr = np.float64(s2)
        self.NegLnLike = np.float64(nll)

    def regneglikelihood(self):
        self.neglikelihood()

    # -- predictors: matrixops.py:68-110 -------------------------------------------------------------
    def _pk_predict(self, X, want, y_min=0.0, ei_weight=-1.0, var_extra=0.0):
        self._pk_ensure_factor()
        if not self._has_U:
            raise AttributeError("'NoneType' object has no attribute 'T'")
        Xq = np.ascontiguousarray(np.atleast_2d(np.asarray(X, dtype=float)))
        mu = float(self.mu)
        s2 = float(self.SigmaSqr) if self.SigmaSqr is not None else float("nan")
        h = self._pk_handle()
        theta, pl = np.asarray(self.theta, dtype=float), np.asarray(self.pl, dtype=float)
        shard = self.__dict__.get("_pk_shard")
        if shard is None or Xq.shape[0] < shard[1]:
            return h.predict_batch(Xq, theta, pl, mu, s2, y_min=y_min, ei_weight=ei_weight, var_extra=var_extra,
                                   want=want)
        names = [nm for nm in ("yhat", "s", "ei") if nm in want]

        def eval_slice(lo, hi):
            o = h.predict_batch(Xq[lo:hi], theta, pl, mu, s2, y_min=y_min, ei_weight=ei_weight,
                                var_extra=var_extra, want=want)
            return np.stack([o[nm] for nm in names], axis=1)

        h.set_shard_hint(query_points=Xq.shape[0])
        try:
            rows = sharding.sharded_rows(Xq.shape[0], eval_slice, shard[0], ncols=len(names))
        finally:
            h.set_shard_hint()
        out = {"yhat": None, "s": None, "ei": None}
        for c, nm in enumerate(names):
            out[nm] = rows[:, c].copy()
        return out

    def predict_normalized(self, x):
        return np.float64(self._pk_predict(x, ("yhat",))["yhat"][0])

    def predicterr_normalized(self, x):
        return np.float64(self._pk_predict(x, ("s",))["s"][0])

    def regression_predicterr_normalized(self, x):
        return np.float64(self._pk_predict(x, ("s",), var_extra=float(self.Lambda))["s"][0])

    # -- batched forms (what the per-point loops of the reference should call) -----------------------
    def predict_normalized_batch(self, X):
        return self._pk_predict(X, ("yhat",))["yhat"]

    def predicterr_normalized_batch(self, X):
        return self._pk_predict(X, ("s",))["s"]

    def expimp_batch(self, X, w=None):
        """expimp / weightedexpimp (krige.py:201-234) for many model-unit points at once."""
        y_min = float(np.min(self.y))
        return self._pk_predict(X, ("ei",), y_min=y_min, ei_weight=-1.0 if w is None else float(w))["ei"]

    def predict_all_batch(self, X, w=None):
        y_min = float(np.min(self.y))
        return self._pk_predict(X, ("yhat", "s", "ei"), y_min=y_min, ei_weight=-1.0 if w is None else float(w))

    # -- arg-max over a query set: what infill is after (krige.py:248-258, 292) --------------------------
    def _pk_argmax(self, X, name, **kw):
        """(max value, index into X) of one predictor output over the query set X.  With sharding enabled every rank
        evaluates only ITS rows of X and the ranks exchange one (value, index) pair each -- ONE all_gather of 16 bytes
        per rank; the full arrays never leave the GPU that computed them.  Ties resolve to the lowest index, as max()
        over the reference's in-order Python list does."""
        Xq = np.ascontiguousarray(np.atleast_2d(np.asarray(X, dtype=float)))
        shard = self.__dict__.get("_pk_shard")
        if shard is None:
            vals = self._pk_predict(Xq, (name,), **kw)[name]
            i = int(np.argmax(vals))
            return float(vals[i]), i
        self._pk_ensure_factor()
        if not self._has_U:
            raise AttributeError("'NoneType' object has no attribute 'T'")
        h = self._pk_handle()
        rank, world = sharding.rank_and_world(shard[0])
        lo, hi = sharding.shard_bounds(Xq.shape[0], rank, world)
        h.set_shard_hint(query_points=Xq.shape[0])
        err, vals = None, np.zeros(0)
        try:
            if hi > lo:
                vals = h.predict_batch(Xq[lo:hi], np.asarray(self.theta, dtype=float), np.asarray(self.pl, dtype=float),
                                       float(self.mu), float(self.SigmaSqr), want=(name,), **kw)[name]
        except Exception as e:  # noqa: BLE001 -- reaches every rank through the collective below
            err = e
        finally:
            h.set_shard_hint()
        try:
            return sharding.global_argmax(vals, lo, shard[0], failed=err is not None)
        except sharding.ShardError:
            if err is not None:
                raise err
            raise

    def expimp_argmax(self, X, w=None):
        """Largest expected improvement over the model-unit points X: (EI, index)."""
        return self._pk_argmax(X, "ei", y_min=float(np.min(self.y)), ei_weight=-1.0 if w is None else float(w))

    def predicterr_argmax(self, X):
        """Largest predicted RMSE over the model-unit points X: (s, index)."""
        return self._pk_argmax(X, "s")
