"""CPU: the Winston-Lutz 2-D oracle restatement (oracle/wl_oracle.py) against golden vectors produced by the UNMODIFIED reference
WinstonLutz2D run with the restated skimage functions of oracle/skimage_shim.py (tests/golden/make_wl_golden.py).
The skimage boundary itself is unpinned (no skimage in this container); everything above it is pinned bit for bit."""
import hashlib
import warnings

import numpy as np
import pytest

from oracle import wl_oracle
from tests.golden.wl_cases import CASES, case_frame

GOLD = np.load("tests/golden/wl_golden.npz")
KEYS = ["field_cax", "bb", "epid", "cax2bb_vector", "cax2bb_distance", "cax2epid_vector", "cax2epid_distance"]


@pytest.mark.parametrize("name", CASES)
def test_oracle_matches_reference_golden(name):
    a, ps, sid, g, c, p, ak = case_frame(name)
    sha = np.frombuffer(hashlib.sha1(a.tobytes()).digest(), dtype=np.uint8)
    assert np.array_equal(sha, GOLD[f"{name}/input_sha1"]), "synthetic input drifted from the one the golden was made with"
    dpmm = (1 / ps) * sid / 1000.0
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        if f"{name}/raises" in GOLD:
            with pytest.raises(ValueError):
                wl_oracle.wl2d_analyze(a, dpmm, **ak)
            return
        o = wl_oracle.wl2d_analyze(a, dpmm, **ak)
    assert np.array_equal(o["shape"], GOLD[f"{name}/shape"])
    for k in KEYS:
        np.testing.assert_allclose(np.asarray(o[k], dtype=float), GOLD[f"{name}/{k}"], rtol=0, atol=1e-9, err_msg=k)
