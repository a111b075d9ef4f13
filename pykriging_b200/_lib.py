"""ctypes binding of libpykriging_b200.so -- the C ABI declared in include/pykriging_b200.h.

There is deliberately no fallback: if the shared library has not been built, or no CUDA device
is usable, this module raises.  Build with ``python -m pykriging_b200.build`` (or
``__graft_entry__.build()``).
"""
import ctypes
import os

import numpy as np

_PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_PKG, "libpykriging_b200.so")

MODE_INTERP = 0
MODE_REGRESSION = 1

# every symbol include/pykriging_b200.h declares (tests check the library exports all of them)
EXPORTS = [
    "pk_version", "pk_last_error", "pk_create", "pk_destroy", "pk_set_data", "pk_nll_batch", "pk_fit",
    "pk_predict_batch", "pk_append_point", "pk_get_psi", "pk_get_U", "pk_psi_batch", "pk_synchronize", "pk_stream",
    "pk_launch_count", "pk_set_workspace_limit", "pk_set_pivot_tolerance", "pk_set_profiling", "pk_phase_times",
    "pk_fp64_peak", "pk_debug_math", "pk_set_cholesky_variant", "pk_set_predict_variant", "pk_set_overlap",
    "pk_set_shard_hint", "pk_set_shard_offset", "pk_set_task_options", "pk_task_schedule", "pk_opt_set", "pk_opt_get", "pk_opt_evaluate",
    "pk_ga_step", "pk_pso_step", "pk_set_predict_contraction",
]

_lib = None


class PkError(RuntimeError):
    pass


def load():
    """Load the shared library (once) and declare the C signatures."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise PkError(
            "libpykriging_b200.so is missing (%s). Build it with `python -m pykriging_b200.build`; "
            "pykriging_b200 has no CPU fallback." % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)
    c_dp = ctypes.c_void_p  # double* / int32_t* passed as raw addresses (host or device)
    hp = ctypes.c_void_p
    lib.pk_version.restype = ctypes.c_int
    lib.pk_last_error.restype = ctypes.c_char_p
    lib.pk_last_error.argtypes = [hp]
    lib.pk_create.argtypes = [ctypes.POINTER(hp), ctypes.c_int]
    lib.pk_destroy.argtypes = [hp]
    lib.pk_set_data.argtypes = [hp, ctypes.c_int, ctypes.c_int, c_dp, c_dp]
    lib.pk_nll_batch.argtypes = [hp, ctypes.c_int, c_dp, c_dp, c_dp, ctypes.c_int, c_dp, c_dp, c_dp, c_dp, c_dp]
    lib.pk_fit.argtypes = [hp, c_dp, c_dp, ctypes.c_double, ctypes.c_int, c_dp, c_dp]
    lib.pk_predict_batch.argtypes = [hp, ctypes.c_int64, c_dp, c_dp, c_dp, ctypes.c_double, ctypes.c_double,
                                     ctypes.c_double, ctypes.c_double, ctypes.c_double, c_dp, c_dp, c_dp]
    lib.pk_append_point.argtypes = [hp, c_dp, ctypes.c_double, c_dp, c_dp, c_dp]
    lib.pk_append_point.restype = ctypes.c_int
    lib.pk_get_psi.argtypes = [hp, c_dp]
    lib.pk_get_U.argtypes = [hp, c_dp]
    lib.pk_psi_batch.argtypes = [hp, ctypes.c_int, c_dp, c_dp, c_dp, ctypes.c_int, c_dp]
    lib.pk_synchronize.argtypes = [hp]
    lib.pk_stream.restype = ctypes.c_void_p
    lib.pk_stream.argtypes = [hp]
    lib.pk_launch_count.restype = ctypes.c_int64
    lib.pk_launch_count.argtypes = [hp]
    lib.pk_set_workspace_limit.argtypes = [hp, ctypes.c_int64]
    lib.pk_set_pivot_tolerance.argtypes = [hp, ctypes.c_double]
    lib.pk_set_pivot_tolerance.restype = ctypes.c_int
    lib.pk_set_profiling.argtypes = [hp, ctypes.c_int]
    lib.pk_set_profiling.restype = ctypes.c_int
    lib.pk_phase_times.argtypes = [hp, c_dp]
    lib.pk_phase_times.restype = ctypes.c_int
    lib.pk_fp64_peak.argtypes = [hp, ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_double)]
    lib.pk_fp64_peak.restype = ctypes.c_int
    lib.pk_debug_math.argtypes = [hp, ctypes.c_int64, c_dp, c_dp, c_dp, c_dp]
    lib.pk_debug_math.restype = ctypes.c_int
    lib.pk_set_cholesky_variant.argtypes = [hp, ctypes.c_int]
    lib.pk_set_cholesky_variant.restype = ctypes.c_int
    lib.pk_set_predict_variant.argtypes = [hp, ctypes.c_int]
    lib.pk_set_predict_variant.restype = ctypes.c_int
    lib.pk_set_shard_hint.argtypes = [hp, ctypes.c_int64, ctypes.c_int64]
    lib.pk_set_shard_hint.restype = ctypes.c_int
    lib.pk_set_shard_offset.argtypes = [hp, ctypes.c_int64]
    lib.pk_set_shard_offset.restype = ctypes.c_int
    lib.pk_set_overlap.argtypes = [hp, ctypes.c_int]
    lib.pk_set_overlap.restype = ctypes.c_int
    lib.pk_set_predict_contraction.argtypes = [hp, ctypes.c_int]
    lib.pk_set_predict_contraction.restype = ctypes.c_int
    lib.pk_set_task_options.argtypes = [hp, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int]
    lib.pk_set_task_options.restype = ctypes.c_int
    lib.pk_task_schedule.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_int, c_dp, ctypes.c_int]
    lib.pk_task_schedule.restype = ctypes.c_int
    lib.pk_opt_set.argtypes = [hp, ctypes.c_int, ctypes.c_int, ctypes.c_int, c_dp]
    lib.pk_opt_get.argtypes = [hp, ctypes.c_int, ctypes.c_int, ctypes.c_int, c_dp]
    lib.pk_opt_evaluate.argtypes = [hp, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, c_dp, c_dp, c_dp, c_dp, c_dp]
    lib.pk_ga_step.argtypes = [hp, ctypes.c_int, c_dp, ctypes.c_int, c_dp, c_dp]
    lib.pk_pso_step.argtypes = [hp, c_dp, c_dp, c_dp, ctypes.c_double, ctypes.c_double, ctypes.c_double, c_dp, c_dp]
    for name in ("pk_opt_set", "pk_opt_get", "pk_opt_evaluate", "pk_ga_step", "pk_pso_step"):
        getattr(lib, name).restype = ctypes.c_int
    for name in ("pk_create", "pk_destroy", "pk_set_data", "pk_nll_batch", "pk_fit", "pk_predict_batch",
                 "pk_get_psi", "pk_get_U", "pk_psi_batch", "pk_synchronize", "pk_set_workspace_limit"):
        getattr(lib, name).restype = ctypes.c_int
    _lib = lib
    return lib


def _addr(a):
    """Raw address of a numpy array (host) or of anything with .data_ptr() (torch CUDA tensor)."""
    if a is None:
        return None
    if hasattr(a, "data_ptr"):
        return ctypes.c_void_p(a.data_ptr())
    assert isinstance(a, np.ndarray) and a.flags["C_CONTIGUOUS"]
    return ctypes.c_void_p(a.ctypes.data)


def _f64(a):
    if hasattr(a, "data_ptr"):
        return a
    return np.ascontiguousarray(a, dtype=np.float64)


def task_schedule(nt, single_below=4, blocks_per_task=2):
    """Task list of one candidate in task mode (no device needed): array of (type, j, row block, blocks)."""
    lib = load()
    n = lib.pk_task_schedule(int(nt), int(single_below), int(blocks_per_task), None, 0)
    if n < 0:
        raise PkError("pk_task_schedule: invalid schedule for nt=%d" % nt)
    out = np.zeros((n, 4), dtype=np.int32)
    lib.pk_task_schedule(int(nt), int(single_below), int(blocks_per_task), _addr(out), n)
    return out


class Handle(object):
    """Owner of one pk_handle (one per process and GPU; not thread-safe)."""

    def __init__(self, device=0):
        self.lib = load()
        self._h = ctypes.c_void_p()
        if self.lib.pk_create(ctypes.byref(self._h), int(device)) != 0:
            raise PkError("pk_create failed: %s" % self.lib.pk_last_error(None).decode())
        self.device = int(device)
        self.n = 0
        self.k = 0
        self.pivot_tol = 2.0 * 2.220446049250313e-16  # the library's default (pk_set_pivot_tolerance)

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            self.lib.pk_destroy(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc, what):
        if rc != 0:
            raise PkError("%s failed: %s" % (what, self.lib.pk_last_error(self._h).decode()))

    # -- data -----------------------------------------------------------------------------------
    def set_data(self, X, y):
        X = _f64(X)
        y = _f64(y)
        n, k = X.shape
        self._check(self.lib.pk_set_data(self._h, n, k, _addr(X), _addr(y)), "pk_set_data")
        self.n, self.k = int(n), int(k)

    # -- likelihood -----------------------------------------------------------------------------
    def nll_batch(self, theta, p, lam=None, mode=MODE_INTERP, want=("nll", "mu", "sigma2", "lndet", "info")):
        """Host-array convenience wrapper: returns dict of numpy arrays."""
        theta = _f64(theta)
        p = _f64(p)
        C = theta.shape[0]
        lam_a = _f64(lam) if lam is not None else None
        out = {}
        for name in ("nll", "mu", "sigma2", "lndet"):
            out[name] = np.empty(C, dtype=np.float64) if name in want else None
        out["info"] = np.empty(C, dtype=np.int32) if "info" in want else None
        self._check(self.lib.pk_nll_batch(self._h, C, _addr(theta), _addr(p), _addr(lam_a), int(mode),
                                          _addr(out["nll"]), _addr(out["mu"]), _addr(out["sigma2"]),
                                          _addr(out["lndet"]), _addr(out["info"])), "pk_nll_batch")
        return out

    def nll_batch_raw(self, C, theta, p, lam, mode, nll, mu=None, sigma2=None, lndet=None, info=None):
        """Pointer-level call (host numpy arrays or CUDA tensors); no allocation, no implicit sync for
        device buffers."""
        self._check(self.lib.pk_nll_batch(self._h, int(C), _addr(theta), _addr(p), _addr(lam), int(mode), _addr(nll),
                                          _addr(mu), _addr(sigma2), _addr(lndet), _addr(info)), "pk_nll_batch")

    def fit(self, theta, p, lam=0.0, mode=MODE_INTERP):
        theta = _f64(theta)
        p = _f64(p)
        out4 = np.empty(4, dtype=np.float64)
        info = np.zeros(1, dtype=np.int32)
        self._check(self.lib.pk_fit(self._h, _addr(theta), _addr(p), float(lam), int(mode), _addr(out4), _addr(info)),
                    "pk_fit")
        return out4, int(info[0])

    def append_point(self, x_new, y_new):
        """Bordered O(n^2) update of the cached factor with one more sample (model units).  Returns
        (appended, out4, info); appended False = nothing changed, take the set_data + fit route."""
        x = _f64(x_new)
        out4 = np.empty(4, dtype=np.float64)
        info = np.zeros(1, dtype=np.int32)
        app = np.zeros(1, dtype=np.int32)
        self._check(self.lib.pk_append_point(self._h, _addr(x), float(y_new), _addr(out4), _addr(info), _addr(app)),
                    "pk_append_point")
        if app[0]:
            self.n += 1
        return bool(app[0]), out4, int(info[0])

    # -- predictors -----------------------------------------------------------------------------
    def predict_batch(self, Xq, theta, p, mu, sigma2, y_min=0.0, ei_weight=-1.0, var_extra=0.0,
                      want=("yhat", "s", "ei")):
        Xq = _f64(Xq)
        m = Xq.shape[0]
        theta = np.ascontiguousarray(theta, dtype=np.float64)
        p = np.ascontiguousarray(p, dtype=np.float64)
        out = {name: (np.empty(m, dtype=np.float64) if name in want else None) for name in ("yhat", "s", "ei")}
        self._check(self.lib.pk_predict_batch(self._h, m, _addr(Xq), _addr(theta), _addr(p), float(mu), float(sigma2),
                                              float(y_min), float(ei_weight), float(var_extra), _addr(out["yhat"]),
                                              _addr(out["s"]), _addr(out["ei"])), "pk_predict_batch")
        return out

    def predict_batch_raw(self, m, Xq, theta, p, mu, sigma2, y_min, ei_weight, var_extra, yhat, s, ei):
        self._check(self.lib.pk_predict_batch(self._h, int(m), _addr(Xq), _addr(theta), _addr(p), float(mu),
                                              float(sigma2), float(y_min), float(ei_weight), float(var_extra),
                                              _addr(yhat), _addr(s), _addr(ei)), "pk_predict_batch")

    def get_psi(self):
        out = np.empty((self.n, self.n), dtype=np.float64)
        self._check(self.lib.pk_get_psi(self._h, _addr(out)), "pk_get_psi")
        return out

    def get_U(self):
        out = np.empty((self.n, self.n), dtype=np.float64)
        self._check(self.lib.pk_get_U(self._h, _addr(out)), "pk_get_U")
        return out

    def psi_batch(self, theta, p, lam=None, mode=MODE_INTERP, out=None):
        theta = _f64(theta)
        p = _f64(p)
        C = theta.shape[0]
        lam_a = _f64(lam) if lam is not None else None
        if out is None:
            out = np.empty((C, self.n, self.n), dtype=np.float64)
        self._check(self.lib.pk_psi_batch(self._h, C, _addr(theta), _addr(p), _addr(lam_a), int(mode), _addr(out)),
                    "pk_psi_batch")
        return out

    # -- device-resident optimiser population ------------------------------------------------------
    OPT_POP, OPT_CHILD, OPT_PREV, OPT_ARCHIVE = 0, 1, 2, 3

    def opt_set(self, which, rows):
        rows = np.ascontiguousarray(rows, dtype=np.float64)
        self._check(self.lib.pk_opt_set(self._h, int(which), rows.shape[0], rows.shape[1], _addr(rows)), "pk_opt_set")

    def opt_get(self, which, row0, nrows, dims):
        out = np.empty((int(nrows), int(dims)), dtype=np.float64)
        self._check(self.lib.pk_opt_get(self._h, int(which), int(row0), int(nrows), _addr(out)), "pk_opt_get")
        return out

    def opt_evaluate(self, which, count, mode, row0=0):
        out = {name: np.empty(count, dtype=np.float64) for name in ("nll", "mu", "sigma2", "lndet")}
        out["info"] = np.empty(count, dtype=np.int32)
        self._check(self.lib.pk_opt_evaluate(self._h, int(which), int(row0), int(count), int(mode), _addr(out["nll"]),
                                             _addr(out["mu"]), _addr(out["sigma2"]), _addr(out["lndet"]), _addr(out["info"])),
                    "pk_opt_evaluate")
        return out

    def ga_step(self, nc, keep, parents, cut):
        keep = None if keep is None else np.ascontiguousarray(keep, dtype=np.int32)
        parents = np.ascontiguousarray(parents, dtype=np.int32)
        cut = np.ascontiguousarray(cut, dtype=np.int32)
        self._check(self.lib.pk_ga_step(self._h, int(nc), _addr(keep), int(parents.shape[0]), _addr(parents), _addr(cut)),
                    "pk_ga_step")

    def pso_step(self, r, nbest, improved, inertia, cognitive, social, lo, hi):
        r = np.ascontiguousarray(r, dtype=np.float64)
        nbest = np.ascontiguousarray(nbest, dtype=np.int32)
        improved = None if improved is None else np.ascontiguousarray(improved, dtype=np.int32)
        lo = np.ascontiguousarray(lo, dtype=np.float64)
        hi = np.ascontiguousarray(hi, dtype=np.float64)
        self._check(self.lib.pk_pso_step(self._h, _addr(r), _addr(nbest), _addr(improved), float(inertia), float(cognitive),
                                         float(social), _addr(lo), _addr(hi)), "pk_pso_step")

    # -- misc -----------------------------------------------------------------------------------
    def synchronize(self):
        self._check(self.lib.pk_synchronize(self._h), "pk_synchronize")

    @property
    def stream(self):
        return self.lib.pk_stream(self._h)

    @property
    def launch_count(self):
        return int(self.lib.pk_launch_count(self._h))

    def set_profiling(self, enabled):
        self._check(self.lib.pk_set_profiling(self._h, int(bool(enabled))), "pk_set_profiling")

    def phase_times(self):
        """ms spent since the last call in {psi, cholesky, epilogue, predict} (needs set_profiling(True))."""
        ms = np.zeros(4, dtype=np.float64)
        self._check(self.lib.pk_phase_times(self._h, _addr(ms)), "pk_phase_times")
        return dict(zip(("psi", "cholesky", "epilogue", "predict"), ms.tolist()))

    def fp64_peak(self):
        a, b = ctypes.c_double(), ctypes.c_double()
        self._check(self.lib.pk_fp64_peak(self._h, ctypes.byref(a), ctypes.byref(b)), "pk_fp64_peak")
        return {"dmma_tflops": a.value, "dfma_tflops": b.value}

    def debug_math(self, d, p):
        d = np.ascontiguousarray(d, dtype=np.float64)
        p = np.ascontiguousarray(p, dtype=np.float64)
        a, b = np.empty_like(d), np.empty_like(d)
        self._check(self.lib.pk_debug_math(self._h, d.size, _addr(d), _addr(p), _addr(a), _addr(b)), "pk_debug_math")
        return a, b

    def set_pivot_tolerance(self, tol):
        self._check(self.lib.pk_set_pivot_tolerance(self._h, float(tol)), "pk_set_pivot_tolerance")
        self.pivot_tol = float(tol)

    def set_cholesky_variant(self, variant):
        """0 auto, 1 per-block-column launches, 2 persistent CTA per candidate, 3 the same with gangs of CTAs per candidate,
        4 task mode (persistent CTAs pulling tile tasks; bit-identical to 2)."""
        self._check(self.lib.pk_set_cholesky_variant(self._h, int(variant)), "pk_set_cholesky_variant")

    def set_task_options(self, single_below=-1, blocks_per_task=0, cohort=0, skew=0):
        """Task-mode knobs (pk_set_task_options; defaults = automatic): 64-row tasks for block columns with at most
        `single_below` row blocks left; up to `blocks_per_task` row blocks per panel task; `cohort` candidates walk the
        task list together (0 = all), member m running m % skew entries behind.  None of them changes a bit of the result."""
        self._check(self.lib.pk_set_task_options(self._h, int(single_below), int(blocks_per_task), int(cohort), int(skew)),
                    "pk_set_task_options")

    def set_overlap(self, enabled):
        """Experiment knob, default OFF (measured slower): Psi construction of the second part of a population on a
        second stream under the factorisation of the first."""
        self._check(self.lib.pk_set_overlap(self._h, int(bool(enabled))), "pk_set_overlap")

    def set_predict_variant(self, variant):
        """0 auto, 1 shared-memory psi tile kernel, 2 register-resident kernel (n <= 512), 3 split pipeline, 4 split
        pipeline without the device-side grid detection."""
        self._check(self.lib.pk_set_predict_variant(self._h, int(variant)), "pk_set_predict_variant")

    def set_predict_contraction(self, mode):
        """The |L^-1 psi|^2 contraction of the split predictor: 0 automatic, 1 FP64 DMMA, 2 INT8 tensor cores (tcgen05)."""
        self._check(self.lib.pk_set_predict_contraction(self._h, int(mode)), "pk_set_predict_contraction")

    def set_shard_hint(self, population=0, query_points=0):
        """The next calls evaluate a slice of a population / query set of these sizes (0 = no hint): kernel
        selection follows the whole, so a shard reproduces the unsharded result bit for bit."""
        self._check(self.lib.pk_set_shard_hint(self._h, int(population), int(query_points)), "pk_set_shard_hint")

    def set_shard_offset(self, first_candidate=0):
        """The next pk_nll_batch calls evaluate candidates first_candidate, first_candidate + 1, ... of the hinted
        population (the kernel a candidate takes can depend on its position: stragglers of a short last wave)."""
        self._check(self.lib.pk_set_shard_offset(self._h, int(first_candidate)), "pk_set_shard_offset")

    def set_workspace_limit(self, nbytes):
        self._check(self.lib.pk_set_workspace_limit(self._h, int(nbytes)), "pk_set_workspace_limit")
