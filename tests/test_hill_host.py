"""CPU: the Hill-function regression (pylinac_b200/core/hill.py, mirror of core/hill.py:11-82) against scipy's curve_fit -- the
solver the reference calls -- on seeded penumbrae, plus the closed-form helpers."""
import math
import warnings

import numpy as np
import pytest
from scipy.optimize import curve_fit
from scipy.special import erf

from pylinac_b200.core.hill import Hill, hill_func


@pytest.mark.parametrize("trial", range(16))
def test_fit_matches_curve_fit(trial):
    rng = np.random.default_rng(100 + trial)
    n = int(rng.integers(12, 90))
    x0 = rng.uniform(40, 900)
    x = np.arange(int(x0) - n // 2, int(x0) + n // 2 + 1).astype(float)
    sign = 1 if trial % 2 == 0 else -1
    y = 0.03 + 0.97 * 0.5 * (1 + erf(sign * (x - x0) / rng.uniform(2, 9))) + rng.normal(0, 0.003, len(x))
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        p, _ = curve_fit(hill_func, x, y, p0=(min(y), max(y), np.median(x), 0))
    ours = Hill.fit(x, y)
    ref = Hill.from_params(p)
    assert abs(ours.inflection_idx()["index (exact)"] - ref.inflection_idx()["index (exact)"]) < 1e-6
    np.testing.assert_allclose(ours.params, p, rtol=5e-6)
    xm = float(np.median(x))
    # both solvers stop within their 1.5e-8 tolerances of the same least-squares minimum: the curves agree to that level
    assert abs(ours.y(xm) - ref.y(xm)) < 1e-6 and abs(ours.gradient_at(xm) - ref.gradient_at(xm)) < 1e-5 * max(1.0, abs(ref.gradient_at(xm)))


def test_closed_forms():
    h = Hill.from_params(np.array([0.1, 1.0, 50.0, 8.0]))
    assert abs(h.y(50.0) - 0.55) < 1e-15
    assert abs(h.x(h.y(43.0)) - 43.0) < 1e-10
    infl = h.inflection_idx()
    assert abs(infl["index (exact)"] - 50.0 * math.pow(7 / 9, 1 / 8)) < 1e-12 and infl["index (rounded)"] == 48
    eps = 1e-6
    assert abs(h.gradient_at(47.0) - (h.y(47.0 + eps) - h.y(47.0 - eps)) / (2 * eps)) < 1e-7
    with pytest.raises(TypeError):
        Hill.fit([1.0, 2.0, 3.0], [0.0, 0.5, 1.0])
