"""CPU: the FieldAnalysis oracle restatement (oracle/field_oracle.py) against the committed golden vectors that the
UNMODIFIED reference produced (tests/golden/field_golden.npz, made by tests/golden/make_field_golden.py)."""
import hashlib
import warnings

import numpy as np
import pytest

from oracle import field_oracle
from tests.golden.field_cases import CASES, case_frame

GOLD = np.load("tests/golden/field_golden.npz")
EXACT = ["strip_rows", "strip_cols", "profile_len"]
SKIP = {"input_sha1"} | set(EXACT)
TOP_KEYS = {"top_position_index_x_y", "top_horizontal_distance_from_cax_mm", "top_vertical_distance_from_cax_mm",
            "top_horizontal_distance_from_beam_center_mm", "top_vertical_distance_from_beam_center_mm"}
TOL = 1e-7          # mm / samples / %: the restatement repeats the arithmetic


def oracle_kwargs(ak):
    ak = dict(ak)
    ak.pop("is_FFF", None)
    if "interpolation" in ak and ak["interpolation"] is not None:
        ak["interpolation"] = "Linear"
    return ak


@pytest.mark.parametrize("name", CASES)
def test_oracle_matches_reference_golden(name):
    a, ps, sid, ak = case_frame(name)
    sha = np.frombuffer(hashlib.sha1(a.tobytes()).digest(), dtype=np.uint8)
    assert np.array_equal(sha, GOLD[f"{name}/input_sha1"]), "synthetic input drifted from the one the golden was made with"
    dpmm = (25.4 / ps) / 25.4 * sid / 1000.0
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        o = field_oracle.field_analyze(a, dpmm, **oracle_kwargs(ak))
    for k in EXACT:
        assert np.array_equal(np.asarray(o[k]), GOLD[f"{name}/{k}"]), k
    keys = [k.split("/", 1)[1] for k in GOLD.files if k.startswith(name + "/") and k.split("/", 1)[1] not in SKIP]
    assert len(keys) >= 25
    for k in keys:
        # the five "top" fields: where the reference's L-BFGS-B run on the fitted parabola stops (x0, the first iterate, or the
        # constrained optimum; oracle/field_oracle.py::_lbfgsb_top).  The first iterate carries the reference's finite-difference
        # gradient noise (~1e-8 px), hence the looser bound
        tol = 1e-6 if k in TOP_KEYS else TOL
        np.testing.assert_allclose(np.asarray(o[k], dtype=float), GOLD[f"{name}/{k}"], rtol=0, atol=tol, err_msg=k)


def test_lbfgsb_top_against_scipy_minimize():
    """The restated stopping rule against scipy's own L-BFGS-B on parabolas spanning the three regimes: early exits reproduce the
    optimiser's output to its gradient noise; in the converged regime the optimiser stops within pgtol / |curvature| of the
    constrained optimum (its documented projected-gradient tolerance), which is what the restatement returns."""
    from scipy.optimize import minimize

    lo, hi = 500.0, 780.0
    xm, sc = (lo + hi) / 2, (hi - lo) / 2
    checked = {"x0": 0, "x1": 0, "opt": 0}
    for H in (-4e-6, -1.5e-6, 7e-6, 2e-5, 1e-4):
        for g0 in (3e-6, 9e-6, 2e-5, 4e-5, 1e-4, 3e-4, 1e-3):
            for sgn in (1, -1):
                v = xm + sgn * g0 / abs(H)                 # stationary point of the parabola (a maximum of it when H > 0)
                a = -H / 2
                c2, c1, c0 = a * sc * sc, 2 * a * sc * (xm - v), 1.0 + a * (xm - v) ** 2 - a * 0
                coef = np.array([a, -2 * a * v, a * v * v + 1.0])
                res = minimize(lambda x: -(coef[0] * x**2 + coef[1] * x + coef[2]), x0=(lo + abs(hi - lo) / 2,), bounds=((lo, hi),))
                par = lambda x: coef[0] * x**2 + coef[1] * x + coef[2]
                cands = [lo, hi] + ([v] if lo <= v <= hi and a != 0 else [])
                best = max(cands, key=par)
                c2, c1, c0 = a * sc * sc, 2 * a * (xm - v) * sc, par(xm)
                ours = field_oracle.SP._lbfgsb_top(c2, c1, c0, xm, sc, lo, hi, best)
                x0 = lo + abs(hi - lo) / 2
                if ours == x0:
                    assert res.x[0] == x0
                    checked["x0"] += 1
                elif abs(ours - x0) < 0.01:
                    assert abs(res.x[0] - ours) < 1e-6
                    checked["x1"] += 1
                else:
                    assert abs(res.x[0] - ours) <= 1.5e-5 / abs(H) + 1e-6, (H, g0, res.x[0], ours)
                    checked["opt"] += 1
    assert min(checked.values()) >= 5, checked
