"""Helpers for the golden fixtures written by oracle/make_golden.py (hex floats, recipes)."""
import hashlib
import json
import os

import numpy as np

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
ALL_CASES = ["c1_branin_n20_k2", "c2_squared_n40_k3", "reg_branin_noise_n20_k2",
             "mid_stybtang_n300_k10", "mid_reg_squared_n257_k6", "c3_stybtang_n1024_k10"]
SMALL_CASES = ALL_CASES[:3]


def fh(v):
    if v is None:
        return None
    if isinstance(v, dict):
        return v["int"]
    return float.fromhex(v)


def fha(lst):
    return np.array([float.fromhex(v) for v in lst], dtype=float)


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a, dtype=np.float64).tobytes()).hexdigest()


def load_case(name):
    with open(os.path.join(GOLDEN_DIR, name + ".json")) as f:
        doc = json.load(f)
    n, k = doc["n"], doc["k"]
    if "X" in doc:
        X = fha(doc["X"]).reshape(n, k)
        y = fha(doc["y"])
    else:
        from oracle import oracle_np as onp

        env = {"oracle_np": onp, "np": np, "default_rng": np.random.default_rng}
        for stmt in doc["recipe"].split(";"):
            exec(stmt.strip(), env)
        X, y = env["X"], env["y"]
    assert sha(X) == doc["X_sha256"] and sha(y) == doc["y_sha256"], "golden inputs do not reproduce"
    doc["X_arr"], doc["y_arr"] = X, y
    return doc


def tol_for_cond(cond, base=1e-9):
    """rel. error ~ cond*1e-18 (mu, sigma2); the 1e-9 contract holds to cond~1e9."""
    return max(base, 50.0 * cond * 2.2e-16)


def rel_err(a, b):
    a = np.asarray(a, dtype=float)
    b = np.asarray(b, dtype=float)
    return np.abs(a - b) / np.maximum(np.abs(b), 1e-300)
