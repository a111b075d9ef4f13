"""Helper process of tests/test_gpu_level_a.py (INTEGRATION.md "Level A"): the reference's UNMODIFIED model classes
(pyKriging/krige.py, pyKriging/regressionkrige.py -- from /root/reference in the build container, from the copy the
committed recipe oracle/make_ref.py left under oracle/_ref on the GPU box) with `pyKriging.matrixops` replaced by
`pykriging_b200.matrixops`, i.e. the one-line change of krige.py:7.  Runs the reference's own `fittingObjective`,
`update`, `neglikelihood`, `predict`, `predict_var`, `expimp`, `addPoint` on a golden case and prints the results as
JSON (hex floats).  A process of its own because `pyKriging.matrixops` can be bound only once per interpreter."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import ref_loader  # noqa: E402
from tests.golden_util import fha, load_case  # noqa: E402


def hx(v):
    if isinstance(v, (int, np.integer)) and not isinstance(v, bool):
        return {"int": int(v)}
    return float(v).hex()


def main():
    name = sys.argv[1]
    doc = load_case(name)
    K, R = ref_loader.load_reference_on_b200_mixin()
    import pykriging_b200.matrixops as b200

    cls = R if doc["regression"] else K
    assert cls.__mro__[1] is b200.matrixops and cls.__module__.startswith("pyKriging.")
    out = {"class": "%s.%s" % (cls.__module__, cls.__name__), "mixin": "%s.%s" % (cls.__mro__[1].__module__, cls.__mro__[1].__name__),
           "source": os.path.abspath(sys.modules[cls.__module__].__file__)}
    X, y = np.array(doc["X_arr"]), np.array(doc["y_arr"])
    m = cls(X, y)
    out["fresh_U_is_None"] = m.U is None
    fo = doc["fittingObjective"]
    cands = [list(fha(c)) for c in fo["candidates"]]
    fit = m.fittingObjective(cands, {})          # the reference's own Python loop (krige.py:436-451) on the new mixin
    out["fitness"] = [hx(f) for f in fit]
    out["post"] = {"theta": [hx(v) for v in m.theta], "pl": [hx(v) for v in m.pl],
                   "mu": None if m.mu is None else hx(m.mu), "sigma2": None if m.SigmaSqr is None else hx(m.SigmaSqr)}
    preds = []
    for blk in doc["predictions"]:
        c = fha(blk["candidate"])
        m.update(list(c))
        (m.regneglikelihood if doc["regression"] else m.neglikelihood)()
        rec = {"mu": hx(m.mu), "sigma2": hx(m.SigmaSqr), "nll": hx(m.NegLnLike), "points": []}
        for pt in blk["points"][:6]:
            x, xr = fha(pt["x_model"]), fha(pt["x_real"])
            rec["points"].append({"predict_normalized": hx(m.predict_normalized(x)), "predict": hx(m.predict(np.array(xr))),
                                  "predict_var": hx(m.predict_var(np.array(xr))), "expimp": hx(m.expimp(x))})
        pts = [list(map(float, fha(pt["x_model"]))) for pt in blk["points"]]
        rec["infill_objective_mse"] = [hx(v) for v in m.infill_objective_mse(pts, {})]
        rec["infill_objective_ei"] = [hx(v) for v in m.infill_objective_ei(pts, {})]
        preds.append(rec)
    out["predictions"] = preds
    # addPoint through the reference's own method (np.append + updateData + updateModel, krige.py:135-156)
    n0 = m.n
    xr = np.array(fha(doc["predictions"][0]["points"][1]["x_real"]))
    m.addPoint(xr, m.predict(np.array(xr)))
    out["addpoint"] = {"n": int(m.n), "n0": int(n0), "predict_at_new": hx(m.predict(np.array(xr)))}
    print("LEVEL_A_JSON " + json.dumps(out))


if __name__ == "__main__":
    main()
