"""Hill-function regression of a penumbra (mirror of pylinac/core/hill.py:11-82).

The reference fits ``a + (b - a) / (1 + (c / x)**d)`` with scipy ``curve_fit`` (MINPACK Levenberg-Marquardt, forward-difference
Jacobian, ftol = xtol = 1.49e-8) from ``p0 = (min y, max y, median x, 0)``.  This is host scalar work on a few dozen samples
(SURVEY.md section 2 row 12), restated as a scaled Levenberg-Marquardt iteration with the analytic Jacobian and MINPACK's
trust-region bookkeeping (column scaling by the running column norms, gain-ratio step control); both iterations stop at the same
least-squares minimum, so the fitted parameters agree to the solver tolerance (~1e-7 relative; tests/test_gpu_profiles_ext.py
compares the inflection position with the reference's own fit)."""
from __future__ import annotations

import math

import numpy as np


def hill_func(x, a: float, b: float, c: float, d: float):
    """a: low level, b: high level, c: approximate inflection position, d: slope (hill.py:69-82)"""
    return a + (b - a) / (1.0 + (c / x) ** d)


def _model_and_jacobian(x: np.ndarray, p: np.ndarray):
    a, b, c, d = p
    with np.errstate(all="ignore"):
        u = (c / x) ** d
        den = 1.0 + u
        f = a + (b - a) / den
        J = np.empty((len(x), 4))
        J[:, 0] = u / den
        J[:, 1] = 1.0 / den
        g = -(b - a) / (den * den)
        J[:, 2] = g * d * u / c
        J[:, 3] = g * u * np.log(c / x)
    return f, J


def _levenberg_marquardt(x: np.ndarray, y: np.ndarray, p0, ftol: float = 1.49012e-8, xtol: float = 1.49012e-8,
                         max_iter: int = 1000) -> np.ndarray:
    """min ||f(x; p) - y||^2, MINPACK lmder flow: scaled trust region of radius delta, LM parameter from the step-length
    condition | ||D s|| - delta | <= 0.1 delta, gain ratio rho deciding acceptance and the next radius."""
    p = np.asarray(p0, dtype=np.float64).copy()
    f, J = _model_and_jacobian(x, p)
    r = f - y
    fnorm = float(np.linalg.norm(r))
    D = np.linalg.norm(J, axis=0)
    D[D == 0] = 1.0
    xnorm = float(np.linalg.norm(D * p))
    delta = 100.0 * xnorm if xnorm > 0 else 100.0
    par = 0.0
    for _ in range(max_iter):
        g = J.T @ r
        if fnorm == 0 or not np.all(np.isfinite(g)):
            break
        # gradient-orthogonality test (gtol = 0 in curve_fit: only an exactly zero gradient stops here)
        while True:
            s, par = _lm_step(J, r, D, delta, par)
            pnorm = float(np.linalg.norm(D * s))
            p_new = p + s
            f_new, J_new = _model_and_jacobian(x, p_new)
            r_new = f_new - y
            fnorm1 = float(np.linalg.norm(r_new)) if np.all(np.isfinite(r_new)) else np.inf
            actred = 1.0 - (fnorm1 / fnorm) ** 2 if 0.1 * fnorm1 < fnorm else -1.0
            Js = J @ s
            temp1 = float(np.linalg.norm(Js)) / fnorm
            temp2 = math.sqrt(par) * pnorm / fnorm
            prered = temp1 * temp1 + 2.0 * temp2 * temp2
            dirder = -(temp1 * temp1 + temp2 * temp2)
            ratio = actred / prered if prered != 0 else 0.0
            if ratio <= 0.25:
                temp = 0.5 if actred >= 0 else 0.5 * dirder / (dirder + 0.5 * actred)
                if 0.1 * fnorm1 >= fnorm or temp < 0.1:
                    temp = 0.1
                delta = temp * min(delta, pnorm / 0.1)
                par /= temp
            elif par == 0 or ratio >= 0.75:
                delta = pnorm / 0.5
                par *= 0.5
            if ratio >= 1e-4:                       # successful step
                p, r, J, fnorm = p_new, r_new, J_new, fnorm1
                D = np.maximum(D, np.linalg.norm(J, axis=0))
                xnorm = float(np.linalg.norm(D * p))
            if abs(actred) <= ftol and prered <= ftol and 0.5 * ratio <= 1:
                return p
            if delta <= xtol * xnorm:
                return p
            if abs(actred) <= np.finfo(float).eps and prered <= np.finfo(float).eps and 0.5 * ratio <= 1:
                return p
            if ratio >= 1e-4:
                break
    return p


def _lm_step(J: np.ndarray, r: np.ndarray, D: np.ndarray, delta: float, par: float):
    """MINPACK lmpar: the step s = -(J^T J + par D^2)^-1 J^T r whose scaled length is within 10 % of delta (par = 0 when the
    Gauss-Newton step already is)."""
    A = J.T @ J
    g = J.T @ r

    def step(lam):
        M = A + lam * np.diag(D * D)
        try:
            return -np.linalg.solve(M, g)
        except np.linalg.LinAlgError:
            return -np.linalg.lstsq(M, g, rcond=None)[0]

    rank_ok = np.linalg.matrix_rank(J) == J.shape[1]
    if rank_ok:
        s = step(0.0)
        if np.all(np.isfinite(s)) and np.linalg.norm(D * s) <= 1.1 * delta:
            return s, 0.0
    # bracket and solve phi(par) = ||D s(par)|| - delta = 0 (monotone decreasing in par): Newton on the secular equation, safeguarded
    gnorm = float(np.linalg.norm(g / D))
    paru = gnorm / delta if gnorm > 0 else np.finfo(float).tiny / min(delta, 0.1)
    parl = 0.0
    par = min(max(par, parl), paru)
    if par == 0:
        par = gnorm / max(float(np.linalg.norm(D * step(paru))), np.finfo(float).tiny) if paru > 0 else 1e-3
        par = min(max(par, 1e-12 * paru), paru)
    s = step(par)
    for _ in range(20):
        dxnorm = float(np.linalg.norm(D * s))
        fp = dxnorm - delta
        if abs(fp) <= 0.1 * delta:
            break
        # d||D s||/dpar = -(q^T (A + par D^2)^-1 D^2 ... ) evaluated through one extra solve
        M = A + par * np.diag(D * D)
        q = D * D * s
        try:
            w = np.linalg.solve(M, q)
        except np.linalg.LinAlgError:
            break
        dphi = -float(s @ (D * D * w)) / dxnorm if dxnorm > 0 else 0.0
        if fp > 0:
            parl = max(parl, par)
        else:
            paru = min(paru, par)
        # Newton on 1/phi-like update (More'): par += (phi / delta) * (dxnorm / -dphi)
        if dphi < 0:
            par_new = par + (fp / delta) * (dxnorm / -dphi)
        else:
            par_new = 0.5 * (parl + paru)
        if not (parl < par_new < paru):
            par_new = max(0.5 * (parl + paru), 1e-3 * paru) if parl == 0 else 0.5 * (parl + paru)
        par = par_new
        s = step(par)
    return s, par


class Hill:
    """hill.py:11-66"""

    params: np.ndarray

    @classmethod
    def fit(cls, x_data, y_data) -> "Hill":
        x = np.asarray(x_data, dtype=np.float64)
        y = np.asarray(y_data, dtype=np.float64)
        if len(x) < 4:
            raise TypeError(f"The number of func parameters=4 must not exceed the number of data points={len(x)}")
        inst = cls()
        inst.params = _levenberg_marquardt(x, y, (y.min(), y.max(), np.median(x), 0.0))
        return inst

    @classmethod
    def from_params(cls, params) -> "Hill":
        inst = cls()
        inst.params = params
        return inst

    def inflection_idx(self) -> dict:
        c, d = self.params[2], self.params[3]
        idx = c * math.pow((d - 1) / (d + 1), 1 / d)
        return {"index (exact)": idx, "index (rounded)": int(round(idx))}

    def gradient_at(self, x: float) -> float:
        a, b, c, d = self.params
        cxd = math.pow(c / x, d)
        return (b - a) * d * cxd / (math.pow(cxd + 1, 2) * x)

    def x(self, y: float) -> float:
        a, b, c, d = self.params
        return c * math.pow((y - a) / (b - y), 1 / d)

    def y(self, x: float) -> float:
        a, b, c, d = self.params
        return a + (b - a) / (1 + (c / x) ** d)
