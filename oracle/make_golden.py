"""TEST INFRASTRUCTURE ONLY -- generates tests/golden/*.json by running the UNMODIFIED
reference (``/root/reference``, through ``oracle/ref_loader.py``) in the build container.

    python oracle/make_golden.py            # rewrites tests/golden/

The reference has no tests or golden vectors; these files are the anchor
that pins both ``oracle/oracle_np.py`` and the CUDA path.  Floats are stored as C99 hex
strings (``float.hex``) so they round-trip bit-exactly.  Environment of the generating run
(numpy / BLAS build) is recorded in each file: the last 2-3 digits of ill-conditioned cases
are BLAS-dependent, which is why tests compare with a conditioning-aware
tolerance rather than bit-exactly against these numbers.
"""
import hashlib
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

from oracle import oracle_np as onp  # noqa: E402
from oracle.ref_loader import load_reference, load_sampling_plan  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def hx(v):
    if v is None:
        return None
    if isinstance(v, (int, np.integer)) and not isinstance(v, bool):
        return {"int": int(v)}
    return float(v).hex()


def hxa(a):
    return [float(v).hex() for v in np.asarray(a, dtype=float).ravel()]


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a, dtype=np.float64).tobytes()).hexdigest()


def env():
    import scipy

    cfg = {}
    try:
        cfg = np.show_config(mode="dicts")["Build Dependencies"]["blas"]
    except Exception:
        pass
    return {"numpy": np.__version__, "scipy": scipy.__version__,
            "blas": {k: cfg.get(k) for k in ("name", "version")}}


def per_candidate(model, cands, regression):
    """update(c); neglikelihood() for each candidate separately -> nll/mu/sigma2/lndet/cond."""
    rows = []
    for c in cands:
        row = {"candidate": hxa(c)}
        try:
            if regression:
                # regression updateModel swallows failures: detect them through Psi directly
                model.update(list(c))
                np.linalg.cholesky(model.Psi)
                model.regneglikelihood()
            else:
                model.update(list(c))
                model.neglikelihood()
            row.update(ok=True, nll=hx(model.NegLnLike), mu=hx(model.mu), sigma2=hx(model.SigmaSqr),
                       lndet=hx(model.LnDetPsi), cond=float(np.linalg.cond(np.asarray(model.Psi))))
        except Exception:
            row.update(ok=False)
        rows.append(row)
    return rows


def predictions(model, cand, points_model_units, regression, weights=(0.3,)):
    """Set the state explicitly (update + neglikelihood) then evaluate points."""
    model.update(list(cand))
    if regression:
        model.regneglikelihood()
    else:
        model.neglikelihood()
    out = {"candidate": hxa(cand), "mu": hx(model.mu), "sigma2": hx(model.SigmaSqr),
           "nll": hx(model.NegLnLike), "y_min": hx(np.min(model.y)), "points": []}
    for x in points_model_units:
        x = np.array(x, dtype=float)
        xr = model.inversenormX(x)
        rec = {
            "x_model": hxa(x), "x_real": hxa(xr),
            "predict_normalized": hx(model.predict_normalized(x)),
            "predicterr_normalized": hx(model.predicterr_normalized(x)),
            "predict": hx(model.predict(np.array(xr))),
            "predict_var": hx(model.predict_var(np.array(xr))),
            "expimp": hx(model.expimp(x)),
        }
        for w in weights:
            rec["weightedexpimp_%g" % w] = hx(model.weightedexpimp(x, w))
        out["points"].append(rec)
    pts = [list(map(float, p)) for p in points_model_units]
    out["infill_objective_mse"] = hxa(model.infill_objective_mse(pts, {}))
    out["infill_objective_ei"] = hxa(model.infill_objective_ei(pts, {}))
    return out


def case(name, X, y, cands, pred_cands, points, regression=False, store_inputs=True, recipe=None,
         sequence=None):
    K, R, _ = load_reference()
    cls = R if regression else K
    model = cls(np.array(X), np.array(y))
    doc = {"name": name, "env": env(), "regression": regression, "n": int(model.n), "k": int(model.k),
           "X_sha256": sha(X), "y_sha256": sha(y), "recipe": recipe,
           "normRange": [[hx(a), hx(b)] for a, b in model.normRange],
           "ynormRange": [hx(v) for v in model.ynormRange],
           "Xn_sha256": sha(model.X), "yn_sha256": sha(model.y)}
    if store_inputs:
        doc["X"] = hxa(X)
        doc["y"] = hxa(y)
    # 1) the batched entry point on a FRESH model: list in -> list out (penalties / stale-U)
    fresh = cls(np.array(X), np.array(y))
    seq = sequence if sequence is not None else cands
    fit = fresh.fittingObjective([list(map(float, c)) for c in seq], {})
    doc["fittingObjective"] = {"candidates": [hxa(c) for c in seq], "fitness": [hx(f) for f in fit],
                               "post_theta": hxa(fresh.theta), "post_pl": hxa(fresh.pl),
                               "post_mu": hx(fresh.mu), "post_sigma2": hx(fresh.SigmaSqr),
                               "post_nll": hx(getattr(fresh, "NegLnLike", None))}
    if regression:
        doc["fittingObjective"]["post_Lambda"] = hx(fresh.Lambda)
    # 2) each candidate in isolation
    doc["per_candidate"] = per_candidate(model, cands, regression)
    # 3) predictions on explicitly-set states
    doc["predictions"] = [predictions(model, c, points, regression) for c in pred_cands]
    path = os.path.join(OUT, name + ".json")
    with open(path, "w") as f:
        json.dump(doc, f, indent=0, separators=(",", ":"))
    print("wrote", path, os.path.getsize(path), "bytes")


def main():
    os.makedirs(OUT, exist_ok=True)
    _, _, T = load_reference()
    tf = T()
    only = set(sys.argv[1:])  # optional: names of the cases to (re)generate

    def case_if(name, *a, **kw):
        if not only or name in only:
            case(name, *a, **kw)

    # ---- C1: 2-D Branin, 20-point optimal LHC (BASELINE configs[0]) ---------------------
    X = load_sampling_plan(2, 20)
    y = tf.branin(X)
    fixed = [[10, 0.5, 1.5, 1.9], [50, 50, 1, 1], [1, 1, 2, 2], [1e-5, 1e-5, 2, 2],
             [1.38569, 0.29047, 1.97596, 2.0], [100, 100, 2, 2], [1e-5, 100, 1, 2], [3, 7, 1.0, 1.5]]
    cands = np.array(fixed + onp.candidates_uniform(24, 2, 101).tolist()
                     + onp.candidates_loguniform(24, 2, 102).tolist())
    rng = np.random.default_rng(7)
    pts = [[0.3, 0.6], [0.0, 0.0], [1.0, 1.0], [0.5, 0.5]] + rng.uniform(0, 1, (12, 2)).tolist()
    pts.append(list(onp.normalize_data(X, y)[0][3]))  # exactly on a sample point (s -> 0)
    case_if("c1_branin_n20_k2", X, y, cands, [fixed[0], fixed[4], fixed[7]], pts)

    # ---- C2: 3-D 'squared', 40-point optimal LHC (BASELINE configs[1]) ------------------
    X = load_sampling_plan(3, 40)
    y = tf.squared(X)
    fixed3 = [[1, 1, 1, 2, 2, 2], [0.5, 2.0, 8.0, 1.2, 1.7, 1.99], [1e-5, 1e-5, 1e-5, 2, 2, 2],
              [100, 1e-5, 3, 1, 2, 1.5], [0.05, 0.05, 0.05, 1.9, 1.9, 1.9]]
    cands = np.array(fixed3 + onp.candidates_uniform(32, 3, 201).tolist()
                     + onp.candidates_loguniform(32, 3, 202, lo=-3.0).tolist())
    rng = np.random.default_rng(8)
    pts = [[0.25, 0.25, 0.25], [0, 0, 0], [1, 1, 1]] + rng.uniform(0, 1, (13, 3)).tolist()
    case_if("c2_squared_n40_k3", X, y, cands, [fixed3[1], fixed3[4]], pts)

    # ---- regression kriging, small (2-D noisy Branin, 20 pt) ----------------------------
    X = load_sampling_plan(2, 20)
    rng = np.random.default_rng(9)
    y = tf.branin(X) + rng.normal(0.0, 15.0, 20)  # branin_noise semantics, seeded
    fixedr = [[1e-4, 1e-4, 2, 2, 0.0], [1, 1, 2, 2, 0.1], [1e-4, 1e-4, 2, 2, 0.0], [10, 0.5, 1.9, 1.95, 0.01],
              [5, 5, 1.85, 2.0, 1.0], [1e-4, 1e-4, 2, 2, 0.0], [50, 2, 1.9, 1.9, 0.5]]
    cands = np.array(fixedr + onp.candidates_uniform(25, 2, 301, regression=True).tolist())
    rng = np.random.default_rng(10)
    pts = [[0.3, 0.6], [0.0, 1.0]] + rng.uniform(0, 1, (10, 2)).tolist()
    case_if("reg_branin_noise_n20_k2", X, y, cands, [fixedr[3], fixedr[4]], pts, regression=True)

    # ---- mid size, drives the blocked large-n path: n=300, k=10 (rlh seed) ---------------
    X = onp.rlh(300, 10, 1234)
    y = onp.f_stybtang_norm(X)
    cands = np.concatenate([onp.candidates_uniform(6, 10, 401), onp.candidates_loguniform(10, 10, 402),
                            np.array([[1e-5] * 10 + [2.0] * 10])])
    rng = np.random.default_rng(11)
    pts = rng.uniform(0, 1, (6, 10)).tolist()
    case_if("mid_stybtang_n300_k10", X, y, cands, [cands[8]], pts, store_inputs=False,
         recipe="X=oracle_np.rlh(300,10,1234); y=oracle_np.f_stybtang_norm(X)")

    # ---- mid size regression, n=257 (ragged vs. every tile size), k=6 ---------------------
    X = onp.rlh(257, 6, 11)
    rng = np.random.default_rng(12)
    y = onp.f_squared(X) + rng.normal(0.0, 0.05, 257)
    # uniform candidates give Psi ~ I (cond < 2: too easy for the tiled regression path); the log-uniform ones with
    # a small nugget reach cond(Psi) 1e5 .. 1e9, and one candidate without nugget is numerically singular
    rng = np.random.default_rng(502)
    hard = np.column_stack([10 ** rng.uniform(-3.5, -0.5, (12, 6)), rng.uniform(1.81231, 2.0, (12, 6)),
                            10 ** rng.uniform(-9.0, -3.0, 12)])
    cands = np.concatenate([onp.candidates_uniform(8, 6, 501, regression=True), hard,
                            np.array([[1e-4] * 6 + [2.0] * 6 + [0.0]])])
    rng = np.random.default_rng(13)
    pts = rng.uniform(0, 1, (6, 6)).tolist()
    case_if("mid_reg_squared_n257_k6", X, y, cands, [cands[0], cands[9]], pts, regression=True, store_inputs=False,
         recipe="X=oracle_np.rlh(257,6,11); y=oracle_np.f_squared(X)+default_rng(12).normal(0,0.05,257)")

    # ---- headline shape C3: n=1024, k=10; a handful of candidates through the reference ---
    X = onp.rlh(1024, 10, 1234)
    y = onp.f_stybtang_norm(X)
    cands = np.concatenate([onp.candidates_uniform(3, 10, 1), onp.candidates_loguniform(5, 10, 2),
                            onp.candidates_loguniform(8, 10, 3, lo=-4.0, hi=0.5)])
    rng = np.random.default_rng(14)
    pts = rng.uniform(0, 1, (3, 10)).tolist()
    case_if("c3_stybtang_n1024_k10", X, y, cands, [cands[4]], pts, store_inputs=False,
         recipe="X=oracle_np.rlh(1024,10,1234); y=oracle_np.f_stybtang_norm(X)")


if __name__ == "__main__":
    main()
