"""CPU: the host-side look-up pieces of pylinac_b200/core/profile.py against the scipy objects the reference uses
(UnivariateSpline(k=1, s=0), interp1d linear / cubic; core/profile.py:249-288, 656-670) and the metric plug-ins on a stand-in
profile (no GPU: the peak search is replaced by scipy.signal for this test only)."""
import numpy as np
import pytest
from scipy.interpolate import UnivariateSpline, interp1d

from pylinac_b200.core import profile as P


def test_linear_spline_matches_univariate_spline():
    rng = np.random.default_rng(1)
    for xk in (np.arange(300.0), np.cumsum(rng.uniform(0.2, 2.0, 80))):
        yk = rng.random(len(xk)) * 1000
        f = UnivariateSpline(x=xk, y=yk, k=1, s=0)
        q = rng.uniform(xk[0], xk[-1], 500)
        np.testing.assert_allclose(P._linear_spline(xk, yk, q), f(q), rtol=1e-13, atol=1e-10)
        assert isinstance(P._linear_spline(xk, yk, float(q[0])), float)
        np.testing.assert_allclose(P._linear_spline(xk, yk, xk), yk, rtol=0, atol=1e-12)


def test_interp1d_linear_matches_scipy_on_unsorted_samples():
    rng = np.random.default_rng(2)
    xs = np.sort(rng.random(60))[::-1] + rng.normal(0, 1e-3, 60)      # a falling edge with noise: not monotone
    ys = np.arange(60.0)
    f = interp1d(x=xs, y=ys)
    for q in rng.uniform(xs.min() + 1e-6, xs.max() - 1e-6, 50):
        assert abs(P._interp1d_linear(xs, ys, float(q)) - float(f(q))) < 1e-10
    with pytest.raises(ValueError):
        P._interp1d_linear(xs, ys, xs.max() + 1.0)


def test_cubic_spline_and_stationary_point_match_scipy():
    rng = np.random.default_rng(3)
    x = np.arange(400.0)
    y = np.exp(-0.5 * ((x - 137.3) / 4.0) ** 2) - 0.7 * np.exp(-0.5 * ((x - 301.8) / 5.0) ** 2) + rng.normal(0, 1e-4, 400)
    M = P._not_a_knot_cubic(x, y)
    f = interp1d(x, y, kind="cubic")
    q = rng.uniform(0, 399, 300)

    def S(v):
        i = min(int(v), 398)
        t = v - x[i]
        b = (y[i + 1] - y[i]) - (2 * M[i] + M[i + 1]) / 6
        return y[i] + b * t + M[i] / 2 * t * t + (M[i + 1] - M[i]) / 6 * t**3

    assert max(abs(S(v) - float(f(v))) for v in q) < 1e-12
    from scipy.optimize import minimize

    left = P._cubic_stationary_near(x, y, M, int(np.argmax(y)), want_max=True)
    right = P._cubic_stationary_near(x, y, M, int(np.argmin(y)), want_max=False)
    assert abs(left - minimize(lambda v: -f(v), x0=float(np.argmax(y))).x[0]) < 1e-3
    assert abs(right - minimize(f, x0=float(np.argmin(y))).x[0]) < 1e-3
    assert abs(left - 137.3) < 0.05 and abs(right - 301.8) < 0.05
