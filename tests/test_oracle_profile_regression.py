"""CPU: the SingleProfile oracle (oracle/field_oracle.py:SP) against the reference's OWN frozen regression vectors
(tests_basic/core/profile_regression_fixtures.py; pinned to 1e-9 by tests_basic/core/test_profile.py:2546-2687), variants
without x_values, interpolation NONE and LINEAR -- extracted by tests/golden/make_profile_regression.py."""
import math

import numpy as np
import pytest

from oracle.field_oracle import SP

REG = np.load("tests/golden/profile_regression.npz")
N = len(REG["names"])
VARIANTS = {"no_x": None, "linear_no_x": "Linear"}


def oracle_metrics(sp):
    fd = sp.field_data(0.8, 0.2)
    fv = np.asarray(fd["field values"], dtype=float)
    dmax, dmin = fv.max(), fv.min()
    pd = 100 * (fv - fv[::-1]) / fd["beam center value (@rounded)"]
    s1, s2 = fv / fv[::-1], fv[::-1] / fv
    pdq = np.maximum(np.abs(s1), np.abs(s2)) * np.where(np.abs(s1) > np.abs(s2), np.sign(s1), np.sign(s2))
    n = len(fv)
    al, ar = fv[: math.floor(n / 2)].sum(), fv[math.ceil(n / 2):].sum()
    return {"varian_flatness_difference": 100 * abs(dmax - dmin) / (dmax + dmin),
            "siemens_flatness_difference": 100 * abs(dmax - dmin) / (dmax + dmin),
            "elekta_flatness_ratio": 100 * (dmax / dmin),
            "varian_symmetry_point_difference": pd[np.argmax(np.abs(pd))],
            "elekta_symmetry_pdq": pdq[np.argmax(np.abs(pdq))],
            "siemens_symmetry_area": 100 * (al - ar) / (al + ar)}


@pytest.mark.parametrize("variant", list(VARIANTS))
@pytest.mark.parametrize("k", range(N))
def test_oracle_matches_frozen_reference_metrics(k, variant):
    sp = SP(REG[f"{k}/values"], interpolation=VARIANTS[variant])
    got = oracle_metrics(sp)
    for key, exp in zip(REG[f"{k}/{variant}/keys"], REG[f"{k}/{variant}/vals"]):
        assert abs(got[str(key)] - exp) <= 1e-9, (str(REG["names"][k]), str(key), got[str(key)], exp)
