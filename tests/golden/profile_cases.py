"""Seeded 1-D profiles for the SingleProfile parity tests (toy open-field profiles with penumbra, horns, noise, tilt)."""
from __future__ import annotations

import numpy as np
from scipy import special

CASES = ["default", "dpmm_linear", "inflection", "inflection_dpmm", "no_interp", "norm_max", "norm_geo", "no_ground", "geo_centering",
         "narrow"]


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
    raise KeyError(name)
