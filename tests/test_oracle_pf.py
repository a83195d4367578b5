"""CPU: the oracle restatement (oracle/pf_oracle.py) against the committed golden vectors that the
UNMODIFIED reference produced (tests/golden/pf_golden.npz, made by tests/golden/make_pf_golden.py)."""
import hashlib
import warnings

import numpy as np
import pytest

from oracle import pf_oracle
from tests.golden.pf_cases import CASES, case_frame

GOLD = np.load("tests/golden/pf_golden.npz") if __import__("os").path.exists("tests/golden/pf_golden.npz") else None

EXACT = ["orientation", "n_meas", "meas_leaf", "meas_picket", "picket_idx", "max_error_picket", "passed", "number_of_pickets"]
CLOSE = ["meas_position", "meas_error", "meas_width_mm", "picket_spacing", "fits", "percent_passing", "max_error",
         "abs_median_error", "offsets_from_cax_mm", "mean_picket_spacing", "mlc_skew", "picket_widths"]


def run_oracle(name):
    a, ps, sid, ck, ak = case_frame(name)
    ck = dict(ck)
    if ck.get("mlc") == "HD":
        ck["mlc"] = "HD Millennium"
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        return a, pf_oracle.pf_analyze(a, (1 / ps) * sid / 1000.0, **ck, **ak)


@pytest.mark.parametrize("name", CASES)
def test_oracle_matches_reference_golden(name):
    a, _, _, _, _ = case_frame(name)
    sha = np.frombuffer(hashlib.sha1(a.tobytes()).digest(), dtype=np.uint8)
    assert np.array_equal(sha, GOLD[f"{name}/input_sha1"]), "synthetic input drifted from the one the golden was made with"
    if f"{name}/raises" in GOLD:
        with pytest.raises(ValueError):
            run_oracle(name)
        return
    _, o = run_oracle(name)
    for k in EXACT:
        assert np.array_equal(np.asarray(o[k]), GOLD[f"{name}/{k}"]), k
    for k in CLOSE:
        # the restatement performs the same float operations in the same order: bit-identical expected
        np.testing.assert_array_equal(np.asarray(o[k]), GOLD[f"{name}/{k}"], err_msg=k)
    assert str(o["max_error_leaf"]) == str(GOLD[f"{name}/max_error_leaf"])
    assert [str(x) for x in o["failed_leaves"]] == [str(x) for x in GOLD[f"{name}/failed_leaves"]]
