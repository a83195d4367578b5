"""Seeded 1-D profiles for the SingleProfile parity tests (toy open-field profiles with penumbra, horns, noise, tilt)."""
from __future__ import annotations

import numpy as np
from scipy import special

CASES = ["default", "dpmm_linear", "inflection", "inflection_dpmm", "no_interp", "norm_max", "norm_geo", "no_ground", "geo_centering",
         "narrow", "spline", "spline_dpmm_inflection", "hill", "hill_dpmm_fff", "hill_spline_none_norm", "xvals_linear",
         "xvals_none_uneven", "xvals_spline_hill"]


def _field(n, left, right, pen, seed, horns=0.0, tilt=0.0, floor=0.02, noise=0.004):
    x = np.arange(n, dtype=float)
    prof = 0.5 * (special.erf((x - left) / pen) - special.erf((x - right) / pen))
    mid = (left + right) / 2
    prof = prof * (1 + horns * ((x - mid) / (right - left)) ** 2 + tilt * (x - mid) / (right - left))
    rng = np.random.default_rng(seed)
    return prof + floor + rng.normal(0, noise, n)


def case_profile(name):
    """-> (values, SingleProfile kwargs (enum values as in the reference), query parameters)"""
    q = {"fwxm_x": 50, "penumbra": (20, 80), "in_field_ratio": 0.8, "slope_exclusion_ratio": 0.2}
    if name == "default":
        return _field(400, 100.3, 310.6, 6.0, 1), {}, q
    if name == "dpmm_linear":
        return _field(640, 180.2, 470.9, 8.0, 2, horns=0.15), {"dpmm": 2.56}, dict(q, fwxm_x=30, penumbra=(10, 90))
    if name == "inflection":
        return _field(500, 120.7, 390.1, 7.0, 3, horns=0.2), {"edge_detection_method": "Inflection Derivative"}, q
    if name == "inflection_dpmm":
        return _field(1280, 400.4, 900.2, 9.0, 4, horns=0.1, tilt=0.03), {"dpmm": 2.976, "edge_detection_method": "Inflection Derivative",
                                                                          "interpolation_resolution_mm": 0.1}, dict(q, in_field_ratio=0.7, slope_exclusion_ratio=0.3)
    if name == "no_interp":
        return _field(800, 250.5, 590.5, 10.0, 5), {"interpolation": None, "dpmm": 1.0}, q
    if name == "norm_max":
        return _field(400, 90.0, 300.0, 5.0, 6, horns=0.3), {"normalization_method": "Max"}, dict(q, fwxm_x=70)
    if name == "norm_geo":
        return _field(401, 110.0, 320.0, 5.0, 7), {"normalization_method": "Geometric center", "interpolation_factor": 5}, q
    if name == "no_ground":
        return _field(400, 100.0, 310.0, 6.0, 8, floor=0.1), {"ground": False, "normalization_method": None}, q
    if name == "geo_centering":
        return _field(600, 150.0, 420.0, 7.0, 9, tilt=0.05), {"centering": "Geometric center", "dpmm": 2.0, "interpolation_resolution_mm": 0.2}, q
    if name == "narrow":
        return _field(300, 130.0, 170.0, 4.0, 10), {"interpolation_factor": 20}, dict(q, in_field_ratio=0.9)
    if name == "spline":
        return _field(400, 100.3, 310.6, 6.0, 21), {"interpolation": "Spline"}, q
    if name == "spline_dpmm_inflection":
        return _field(512, 140.2, 380.9, 7.0, 22, horns=0.15), {"interpolation": "Spline", "dpmm": 2.56,
                                                                "edge_detection_method": "Inflection Derivative"}, q
    if name == "hill":
        return _field(500, 120.7, 390.1, 7.0, 23, horns=0.2), {"edge_detection_method": "Inflection Hill"}, q
    if name == "hill_dpmm_fff":
        return _field(640, 170.4, 470.2, 9.0, 24, horns=-1.3), {"edge_detection_method": "Inflection Hill", "dpmm": 2.56,
                                                                 "hill_window_ratio": 0.15}, dict(q, penumbra=(10, 90))
    if name == "hill_spline_none_norm":
        return _field(450, 110.0, 340.0, 6.0, 25), {"edge_detection_method": "Inflection Hill", "interpolation": "Spline",
                                                     "normalization_method": None, "ground": False}, q
    if name == "xvals_linear":
        return _field(300, 80.0, 220.0, 5.0, 26), {"x_values": np.linspace(-75.0, 74.5, 300)}, q
    if name == "xvals_none_uneven":
        x = np.cumsum(np.where(np.arange(260) % 7 == 3, 1.5, 1.0)) - 140.0
        return _field(260, 70.0, 190.0, 5.0, 27), {"x_values": x, "interpolation": None}, q
    if name == "xvals_spline_hill":
        return _field(200, 50.0, 150.0, 4.0, 28, horns=0.1), {"x_values": np.arange(200) * 0.5 + 3.0, "interpolation": "Spline",
                                                                "edge_detection_method": "Inflection Hill", "interpolation_factor": 4}, q
    raise KeyError(name)
