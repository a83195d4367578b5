"""oracle/vmat_oracle.py against goldens of the UNMODIFIED reference (tests/golden/make_vmat_golden.py): DRGS / DRMLC / DLG."""
import warnings

import numpy as np
import pytest

from oracle import vmat_oracle
from oracle.pf_oracle import mlc_arrangement
from tests.golden.vmat_cases import DLG_CASES, VMAT_CASES, dlg_case, vmat_case

GOLD = np.load("tests/golden/vmat_golden.npz", allow_pickle=False)
DEFAULT_OFFSETS = {"DRGS": (-60, -40, -20, 0, 20, 40, 60), "DRMLC": (-45, -15, 15, 45)}
MLC_RUNS = {"MILLENNIUM": [(10, 10), (40, 5), (10, 10)], "HD_MILLENNIUM": [(14, 5), (32, 2.5), (14, 5)]}


def oracle_kwargs(klass, ck, ak):
    roi = ak.get("roi_config")
    return dict(offsets_mm=[v["offset_mm"] for v in roi.values()] if roi else DEFAULT_OFFSETS[klass], tolerance=ak.get("tolerance", 1.5),
                segment_size_mm=ak.get("segment_size_mm", (5, 100)), ground=ck.get("ground", True), check_inv=ck.get("check_inversion", True),
                invert_image_order=ak.get("invert_image_order", False))


@pytest.mark.parametrize("name", VMAT_CASES)
def test_vmat_oracle_matches_reference(name):
    klass, a, b, ps, sid, ck, ak = vmat_case(name)
    dpmm = (1 / ps) * sid / 1000.0
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        o = vmat_oracle.vmat_analyze(a, b, dpmm, **oracle_kwargs(klass, ck, ak))
    g = lambda k: GOLD[f"{name}/{k}"]
    assert o["open_is_first"] == int(g("open_is_first"))
    assert bool(o["center_warning"]) == bool(int(g("n_user_warnings")))
    for k in ("r_corr", "r_dev", "stdev", "center_x", "center_y"):
        np.testing.assert_allclose(o[k], g(k), rtol=1e-10, atol=1e-10, equal_nan=True, err_msg=k)
    assert np.array_equal(o["passed_seg"], g("passed_seg"))
    assert o["passed"] == bool(g("passed"))
    for k in ("max_r_deviation", "avg_abs_r_deviation", "avg_r_deviation"):
        np.testing.assert_allclose(o[k], g(k), rtol=1e-9, atol=1e-9, equal_nan=True, err_msg=k)


def test_contrived_case_meets_the_reference_tests_expectation():
    """tests_basic/test_vmat.py:708-720: R_corr 100 +- 1, R_dev 0 +- 1, segment centres (506, 640) / (685, 640) +- 5 px, passes"""
    g = lambda k: GOLD[f"drmlc_contrived/{k}"]
    assert abs(g("center_x")[0] - 506) < 5 and abs(g("center_x")[2] - 685) < 5 and abs(g("center_y")[0] - 640) < 5
    assert np.all(np.abs(g("r_corr")[[0, 2]] - 100) < 1) and np.all(np.abs(g("r_dev")[[0, 2]]) < 1)
    assert bool(g("passed")) and float(g("max_r_deviation")) < 0.1


@pytest.mark.parametrize("name", DLG_CASES)
def test_dlg_oracle_matches_reference(name):
    frame, ps, sid, gaps, mlc, yfs, pw = dlg_case(name)
    dpmm = (1 / ps) * sid / 1000.0
    _, centers, widths = mlc_arrangement(MLC_RUNS[mlc])
    o = vmat_oracle.dlg_analyze(frame, dpmm, gaps, centers, widths, yfs, pw)
    for k in ("measured_dlg", "measured_dlg_per_leaf", "planned_dlg_per_leaf", "slope", "intercept"):
        np.testing.assert_allclose(o[k], GOLD[f"{name}/{k}"], rtol=1e-12, atol=1e-12, err_msg=k)
