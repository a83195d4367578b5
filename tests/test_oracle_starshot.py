"""CPU: the Starshot oracle restatement (oracle/starshot_oracle.py) against the committed golden vectors that the
UNMODIFIED reference produced (tests/golden/starshot_golden.npz, made by tests/golden/make_starshot_golden.py)."""
import hashlib
import warnings

import numpy as np
import pytest

from oracle import starshot_oracle
from tests.golden.starshot_cases import CASES, case_frame

GOLD = np.load("tests/golden/starshot_golden.npz")
EXACT = ["iterations", "profile_len", "peak_idx", "n_lines", "passed"]
CLOSE = ["radius_px", "peak_xy", "wobble_center", "wobble_radius_px", "wobble_radius_mm", "angles"]


@pytest.mark.parametrize("name", CASES)
def test_oracle_matches_reference_golden(name):
    a, ps, sid, ak = case_frame(name)
    sha = np.frombuffer(hashlib.sha1(a.tobytes()).digest(), dtype=np.uint8)
    assert np.array_equal(sha, GOLD[f"{name}/input_sha1"]), "synthetic input drifted from the one the golden was made with"
    dpmm = (1 / ps) * sid / 1000.0
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        if f"{name}/raises" in GOLD:
            with pytest.raises(RuntimeError):
                starshot_oracle.starshot_analyze(a, dpmm, **ak)
            return
        o = starshot_oracle.starshot_analyze(a, dpmm, **ak)
    for k in EXACT:
        assert np.array_equal(np.asarray(o[k]), GOLD[f"{name}/{k}"]), k
    for k in CLOSE:
        # same scipy calls on the same data in the same order: identical up to the last bit of dpmm (dpi / 25.4 vs 1 / pixel size)
        np.testing.assert_allclose(np.asarray(o[k]), GOLD[f"{name}/{k}"], rtol=1e-14, atol=0, err_msg=k)
