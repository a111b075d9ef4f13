"""GPU tests of the reference-facing classes (``kriging`` / ``regression_kriging`` on the CUDA
``matrixops``) against the CPU oracle: batched fittingObjective incl. penalty / stale-U semantics and
post-state, update/predict/expimp, addPoint, infill objectives, and a full ``train()`` under a fixed
optimiser seed against the inspyred restatement driving the oracle likelihood."""
from random import Random

import numpy as np
import pytest
from scipy.optimize import minimize

from oracle import inspyred_port as ip
from oracle import oracle_np as onp
from tests.golden_util import SMALL_CASES, fh, fha, load_case, rel_err

pytestmark = pytest.mark.gpu


def _classes():
    from pykriging_b200.krige import kriging
    from pykriging_b200.regressionkrige import regression_kriging

    return kriging, regression_kriging


@pytest.mark.parametrize("name", SMALL_CASES)
def test_fitting_objective_sequence_and_poststate(name):
    """Same list in -> same list out as the unmodified reference (golden), incl. the int 10000 penalty
    (krige.py:447-450) and the regression stale-factor repeats (regressionkrige.py:167-172)."""
    doc = load_case(name)
    K, R = _classes()
    m = (R if doc["regression"] else K)(doc["X_arr"], doc["y_arr"])
    fo = doc["fittingObjective"]
    cands = [list(fha(c)) for c in fo["candidates"]]
    got = m.fittingObjective(cands, {})
    want = [fh(v) for v in fo["fitness"]]
    assert len(got) == len(want)
    for g, w in zip(got, want):
        if isinstance(w, int):
            assert isinstance(g, int) and g == 10000
        else:
            assert not isinstance(g, int)
            assert rel_err(g, w) < 1e-7  # some golden candidates have cond(Psi) up to 1e7
    assert np.array_equal(m.theta, fha(fo["post_theta"]))
    assert np.array_equal(m.pl, fha(fo["post_pl"]))
    if fo["post_mu"] is not None:
        assert rel_err(m.mu, fh(fo["post_mu"])) < 1e-7
        assert rel_err(m.SigmaSqr, fh(fo["post_sigma2"])) < 1e-7
        assert rel_err(m.NegLnLike, fh(fo["post_nll"])) < 1e-7
    if doc["regression"]:
        assert m.Lambda == fh(fo["post_Lambda"])


@pytest.mark.parametrize("name", SMALL_CASES)
def test_update_predict_expimp_match_golden(name):
    doc = load_case(name)
    K, R = _classes()
    m = (R if doc["regression"] else K)(doc["X_arr"], doc["y_arr"])
    assert m.U is None and m.Lambda == 1  # fresh-model state of the reference
    for blk in doc["predictions"]:
        c = fha(blk["candidate"])
        m.update(c)
        (m.regneglikelihood if doc["regression"] else m.neglikelihood)()
        assert rel_err(m.mu, fh(blk["mu"])) < 1e-9
        assert rel_err(m.SigmaSqr, fh(blk["sigma2"])) < 1e-9
        assert rel_err(m.NegLnLike, fh(blk["nll"])) < 1e-9
        pts = np.array([fha(p["x_model"]) for p in blk["points"]])
        for p in blk["points"][:6]:
            x = fha(p["x_model"])
            xr = fha(p["x_real"])
            assert rel_err(m.predict_normalized(x), fh(p["predict_normalized"])) < 1e-9
            assert rel_err(m.predict(xr), fh(p["predict"])) < 1e-9
            assert abs(m.predicterr_normalized(x) - fh(p["predicterr_normalized"])) < 1e-7
            assert abs(m.predict_var(xr) - fh(p["predict_var"])) < 1e-7
            assert abs(m.expimp(x) - fh(p["expimp"])) < 1e-7
            assert abs(m.weightedexpimp(x, 0.3) - fh(p["weightedexpimp_0.3"])) < 1e-7
        mse = m.infill_objective_mse(pts.tolist(), {})
        ei = m.infill_objective_ei(pts.tolist(), {})
        assert np.max(np.abs(np.array(mse) - fha(blk["infill_objective_mse"]))) < 1e-7
        assert np.max(np.abs(np.array(ei) - fha(blk["infill_objective_ei"]))) < 1e-7
        U = m.U
        assert U.shape == (m.n, m.n) and np.allclose(U.T @ U, m.Psi, rtol=1e-12, atol=1e-13)


def test_bad_params_raise_and_keep_state():
    K, _ = _classes()
    doc = load_case("c1_branin_n20_k2")
    m = K(doc["X_arr"], doc["y_arr"])
    m.update([10, 0.5, 1.5, 1.9])
    m.neglikelihood()
    U0, mu0 = m.U.copy(), m.mu
    with pytest.raises(Exception, match="bad params"):
        m.update([1e-5, 1e-5, 2, 2])
    assert np.array_equal(m.U, U0) and m.mu == mu0
    assert np.array_equal(m.theta, [1e-5, 1e-5])
    assert m.fittingObjective_local([1e-5, 1e-5, 2, 2]) == 10000
    f = m.fittingObjective_local([10, 0.5, 1.5, 1.9])
    assert rel_err(f, -26.774947339276977) < 1e-9  # SURVEY section 4 known answer


def test_addpoint_matches_oracle():
    K, _ = _classes()
    X = onp.rlh(30, 2, 3)
    y = onp.f_branin(X)
    m = K(X, y)
    m.update([5.0, 3.0, 1.9, 1.7])
    m.neglikelihood()
    newx = np.array([0.31, 0.77])
    newy = float(onp.f_branin(newx)[0])
    m.addPoint(newx, newy)
    assert m.n == 31
    Xn = m.X
    yn = m.y
    ref = onp.predict_batch_fast(Xn, yn, m.theta, m.pl, onp.EPS_DIAG, np.array([[0.5, 0.5]]))
    m.neglikelihood()
    assert rel_err(m.mu, ref["mu"]) < 1e-9
    assert rel_err(m.predict_normalized(np.array([0.5, 0.5])), ref["yhat"][0]) < 1e-9


def _oracle_train(X, y, optimizer, seed, pop, regression=False):
    """The reference's train() (krige.py:355-427) restated on the CPU oracle + inspyred port."""
    ora = onp.OracleKriging(X, y, regression=regression)
    k = ora.k
    if regression:
        lo = [1e-4] * k + [1.81231] * k + [0]
        hi = [100] * k + [2] * k + [1]
    else:
        lo = [1e-5] * k + [1] * k
        hi = [100] * k + [2] * k

    def gen(random, args):
        b = args["_ec"].bounder
        return [random.uniform(l, h) for l, h in zip(b.lower_bound, b.upper_bound)]

    def term(population, num_generations, num_evaluations, args):
        mg = args.setdefault('max_generations', 10)
        pb = args.setdefault('previous_best', None)
        me = args.setdefault('max_evaluations', 30000)
        cb = np.around(max(population).fitness, decimals=4)
        if pb is None or pb != cb:
            args['previous_best'] = cb
            args['generation_count'] = 0
            return False or (num_evaluations >= me)
        if args['generation_count'] >= mg:
            return True
        args['generation_count'] += 1
        return False or (num_evaluations >= me)

    def evaluator(candidates, args):
        return ora.fittingObjective(candidates)

    if optimizer == 'pso':
        ea = ip.PSO(Random(seed))
        ea.terminator = term
        final = ea.evolve(generator=gen, evaluator=evaluator, pop_size=pop, maximize=False,
                          bounder=ip.Bounder(lo, hi), max_evaluations=30000, neighborhood_size=20, num_inputs=k)
        final.sort(reverse=True)
    else:
        ea = ip.GA(Random(seed))
        ea.terminator = term
        final = ea.evolve(generator=gen, evaluator=evaluator, pop_size=pop, maximize=False,
                          bounder=ip.Bounder(lo, hi), max_evaluations=30000, num_elites=10, mutation_rate=.05)
    return ora, ea, final


@pytest.mark.parametrize("optimizer,case,pop", [("ga", "c1_branin_n20_k2", 300), ("pso", "c2_squared_n40_k3", 256)])
def test_train_selects_same_best_under_fixed_seed(optimizer, case, pop):
    """BASELINE configs[0]/[1]: the global search must pick the same best theta/p as the CPU restatement
    under a fixed optimiser seed (inspyred restatement: parity unpinned, see oracle/inspyred_port.py)."""
    doc = load_case(case)
    K, _ = _classes()
    seed = 2024
    ora, ea_ref, final_ref = _oracle_train(doc["X_arr"], doc["y_arr"], optimizer, seed, pop)
    m = K(doc["X_arr"], doc["y_arr"])
    m.train(optimizer=optimizer, seed=seed, pop_size=pop)
    ea = m._last_ea
    best_ref = final_ref[0]
    best = max(ea.population)
    # The search is chaotic in the last bits: once two candidates of near-equal fitness swap rank the
    # trajectories may diverge.  Require identical generation count and best candidate when the traces
    # agree, and always require the same best fitness to 1e-6 (4-decimal rounding drives termination).
    assert rel_err(best.fitness, best_ref.fitness) < 1e-6
    if ea.num_generations == ea_ref.num_generations:
        assert np.allclose(best.candidate, best_ref.candidate, rtol=1e-9, atol=1e-12)
    # SLSQP polish from the same start on the oracle objective
    lo, hi = m._bounds()
    res = minimize(lambda e: ora.fittingObjective([list(e)])[0], best_ref.candidate, method='SLSQP',
                   bounds=list(zip(lo, hi)), options={'disp': False})
    got = np.concatenate([m.theta, m.pl])
    f_got = m.fittingObjective_local(got)
    assert f_got <= res.fun + 1e-3 * abs(res.fun) + 1e-6


def test_regression_train_runs_and_predicts():
    _, R = _classes()
    doc = load_case("reg_branin_noise_n20_k2")
    m = R(doc["X_arr"], doc["y_arr"])
    m.train(optimizer='ga', seed=5)
    assert 0.0 <= m.Lambda <= 1.0
    m.regneglikelihood()
    v = m.predict(np.array(doc["X_arr"][3]))
    assert np.isfinite(v)
    assert m.predict_var(np.array(doc["X_arr"][3])) >= 0.0


def test_infill_returns_points_in_bounds():
    K, _ = _classes()
    doc = load_case("c1_branin_n20_k2")
    m = K(doc["X_arr"], doc["y_arr"])
    m.update([1.38569, 0.29047, 1.97596, 2.0])
    m.neglikelihood()
    pts = m.infill(2, method='ei', addPoint=True, seed=3)
    assert pts.shape == (2, 2) and m.n == 20
    lo = np.array([r[0] for r in m.normRange])
    hi = np.array([r[1] for r in m.normRange])
    assert np.all(pts >= lo - 1e-12) and np.all(pts <= hi + 1e-12)
    pts2 = m.infill(1, method='error', addPoint=False, seed=4)
    assert pts2.shape == (1, 2)


@pytest.mark.parametrize("case", ["c2_squared_n40_k3", "mid_stybtang_n300_k10"])
def test_slsqp_batched_stencil_matches_sequential(case):
    """the SLSQP polish (krige.py:402-427) with its finite-difference stencil
    evaluated as one population per gradient must produce the iterates of the sequential run bit for bit
    (fused small-n kernel and tiled DMMA path)."""
    doc = load_case(case)
    K, _ = _classes()
    k = doc["X_arr"].shape[1]
    start = [3.0 + 0.5 * d for d in range(k)] + [1.6] * k
    res = []
    for batched in (False, True):
        m = K(doc["X_arr"], doc["y_arr"])
        m.slsqp_batched = batched
        lo, hi = m._bounds()
        r = m._pk_slsqp(start, [[a, b] for a, b in zip(lo, hi)])
        res.append((r, m))
    (r0, m0), (r1, m1) = res
    assert np.array_equal(r0.x, r1.x), (r0.x, r1.x, r0.nit, r1.nit)
    assert r0.fun == r1.fun and r0.nit == r1.nit and r0.nfev == r1.nfev, (r0.fun, r1.fun, r0.nit, r1.nit, r0.nfev, r1.nfev)
    assert m1._pk_stencil_calls == r1.njev and getattr(m0, "_pk_stencil_calls", 0) == 0  # one population per gradient
    # and the polished point is no worse than the oracle's own polish from the same start
    ora = onp.OracleKriging(doc["X_arr"], doc["y_arr"])
    f_start = ora.fittingObjective([start])[0]
    assert r1.fun <= f_start
    assert rel_err(ora.fittingObjective([list(r1.x)])[0], r1.fun) < 1e-8


def test_single_candidate_equals_its_value_inside_a_stencil_population():
    """f(x) evaluated alone (C = 1) and inside a 2k+1 population must be the same bits: SLSQP differences
    them at a step of 1.5e-8."""
    doc = load_case("mid_stybtang_n300_k10")
    K, _ = _classes()
    m = K(doc["X_arr"], doc["y_arr"])
    k = m.k
    x = np.array([2.0 + d for d in range(k)] + [1.3 + 0.05 * d for d in range(k)])
    pts = [x.tolist()]
    for i in range(2 * k):
        xi = x.copy()
        xi[i] += 1.4901161193847656e-08
        pts.append(xi.tolist())
    alone = m.fittingObjective_local(x)
    pop = m.fittingObjective(pts, {})
    assert alone == pop[0]
    assert len(set(pop)) > 1


@pytest.mark.parametrize("regression", [False, True])
@pytest.mark.parametrize("n0", [30, 60, 150])
def test_incremental_addpoint_matches_rebuild(n0, regression):
    """addPoint through the bordered O(n^2) factor update must leave the model in
    the state the reference's rebuild (updateData + updatePsi, krige.py:135-156) leaves it in -- several
    appends in a row, across a 64-row tile boundary (n0 = 60: the library falls back to the rebuild there),
    interpolating and regression models."""
    K, R = _classes()
    cls = R if regression else K
    k = 3
    X = onp.rlh(n0, k, 21)
    y = onp.f_squared(X)
    hyper = [4.0, 2.5, 6.0, 1.9, 1.85, 1.95] + ([0.01] if regression else [])
    new_pts = onp.rlh(7, k, 99) * 0.9 + 0.05
    models = []
    for incremental in (True, False):
        m = cls(X, y)
        m.incremental_addpoint = incremental
        m.update(hyper)
        (m.regneglikelihood if regression else m.neglikelihood)()
        mu_before = m.mu
        for q in new_pts:
            m.addPoint(q, float(onp.f_squared(q[None, :])[0]))
            assert m.mu == mu_before  # addPoint refreshes U, not mu / SigmaSqr
        (m.regneglikelihood if regression else m.neglikelihood)()
        models.append(m)
    inc, full = models
    assert inc.n == full.n == n0 + 7
    # the bordered path really ran (n0 = 60: the append that crosses n = 64 is a rebuild)
    assert inc.__dict__.get("_pk_appends", 0) == (6 if n0 == 60 else 7) and "_pk_appends" not in full.__dict__
    assert rel_err(inc.NegLnLike, full.NegLnLike) < 1e-10
    assert rel_err(inc.mu, full.mu) < 1e-10 and rel_err(inc.SigmaSqr, full.SigmaSqr) < 1e-10
    assert np.allclose(inc.U, full.U, rtol=1e-9, atol=1e-12)
    assert np.allclose(inc.Psi, full.Psi, rtol=0, atol=1e-15)
    q = np.array([0.37, 0.52, 0.61])
    assert rel_err(inc.predict(q), full.predict(q)) < 1e-9
    assert abs(inc.predict_var(q) - full.predict_var(q)) < 1e-8 * max(1.0, float(full.SigmaSqr) ** 0.5)
    assert abs(inc.expimp(inc.normX(q.copy())) - full.expimp(full.normX(q.copy()))) < 1e-8
    # and against the CPU oracle on the grown data set
    ora = onp.OracleKriging(inc.X, inc.y, regression=regression, normalize=False)
    ora.update(hyper)
    ora.neglikelihood()
    assert rel_err(inc.NegLnLike, ora.NegLnLike) < 1e-9


def test_cross_validation_scvr_matches_direct_prediction():
    """the leave-one-out driver (CrossValidation.py:57-123) on the device path.
    Whatever hyper-parameters the n trainings select, every left-out prediction must equal the oracle's
    predictor evaluated with THAT model's theta / p on the (n-1)-point data set."""
    from pykriging_b200.CrossValidation import Cross_Validation

    K, _ = _classes()
    X = onp.rlh(14, 2, 8)
    y = onp.f_branin(X)
    base = K(X, y)
    cv = Cross_Validation(base, name="cv", seed=11, pop_size=24, max_evaluations=600)
    pred, var, scvr = cv.calculate_SCVR(optimiser='ga')
    assert len(pred) == len(var) == len(scvr) == 14 and len(cv.models) == 14
    for i, m in enumerate(cv.models):
        assert m.n == 13
        m.neglikelihood()  # predictions after train() use the mu / SigmaSqr train() left behind (App. A.5)
    ynorm = (base.y - base.y.min()) / (base.y.max() - base.y.min())
    for i, m in enumerate(cv.models):
        xq = m.normX(np.array(base.X[i], dtype=float))
        ref = onp.predict_batch_fast(m.X, m.y, m.theta, m.pl, onp.EPS_DIAG, np.array([xq]))
        assert rel_err(m.predict_normalized(xq), ref["yhat"][0]) < 1e-8
        assert np.isfinite(scvr[i]) and np.isfinite(pred[i]) and var[i] >= 0.0
    # a smooth function sampled on a 14-point plan: the cross-validated predictions track the samples
    resid = np.array([m.normy(p) for m, p in zip(cv.models, pred)]) - ynorm
    assert np.sqrt(np.mean(resid ** 2)) < 0.35
    rmse, r2 = cv.calculate_RMSE_Rsquared('ga', 6)
    assert rmse.startswith('RMSE = ') and r2.startswith('Rsquared = ')
    avg, std = cv.leave_n_out(q=3)
    assert np.isfinite(avg) and np.isfinite(std)
