"""Extract the reference's own frozen SingleProfile regression vectors (tests_basic/core/profile_regression_fixtures.py:
20 exported water-tank / array profiles with protocol metrics pinned to 1e-9 by tests_basic/core/test_profile.py:2546-2687)
into tests/golden/profile_regression.npz: all six variants (interpolation NONE / LINEAR / SPLINE, with the exported detector
positions as x_values and without) plus the frozen field-data geometry.

Run here (the container that has /root/reference):  python -m tests.golden.make_profile_regression
"""
from __future__ import annotations

import importlib.util
import sys

import numpy as np


def main():
    spec = importlib.util.spec_from_file_location("profile_regression_fixtures",
                                                  "/root/reference/tests_basic/core/profile_regression_fixtures.py")
    mod = importlib.util.module_from_spec(spec)
    sys.modules["profile_regression_fixtures"] = mod      # dataclasses resolves the module of the class being defined
    spec.loader.exec_module(mod)
    store = {"names": np.array([f.name for f in mod.PROFILE_REGRESSION_FIXTURES])}
    for k, f in enumerate(mod.PROFILE_REGRESSION_FIXTURES):
        store[f"{k}/values"] = np.asarray(f.values, dtype=np.float64)
        store[f"{k}/x_values"] = np.asarray(f.x_values, dtype=np.float64)
        store[f"{k}/field_data/keys"] = np.array(sorted(f.expected_field_data))
        store[f"{k}/field_data/vals"] = np.array([f.expected_field_data[m] for m in sorted(f.expected_field_data)], dtype=np.float64)
        for variant, attr in (("x", "expected_metrics"), ("linear_x", "expected_metrics_linear"), ("spline_x", "expected_metrics_spline"),
                              ("no_x", "expected_metrics_no_x"), ("linear_no_x", "expected_metrics_linear_no_x"),
                              ("spline_no_x", "expected_metrics_spline_no_x")):
            d = getattr(f, attr)
            store[f"{k}/{variant}/keys"] = np.array(sorted(d))
            store[f"{k}/{variant}/vals"] = np.array([d[m] for m in sorted(d)], dtype=np.float64)
    np.savez_compressed("tests/golden/profile_regression.npz", **store)
    print(len(mod.PROFILE_REGRESSION_FIXTURES), "fixtures;", sorted(mod.PROFILE_REGRESSION_FIXTURES[0].expected_metrics_no_x))


if __name__ == "__main__":
    sys.exit(main())
