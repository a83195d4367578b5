"""CPU: the FieldAnalysis oracle restatement (oracle/field_oracle.py) against the committed golden vectors that the
UNMODIFIED reference produced (tests/golden/field_golden.npz, made by tests/golden/make_field_golden.py)."""
import hashlib
import warnings

import numpy as np
import pytest

from oracle import field_oracle
from tests.golden.field_cases import CASES, case_frame

GOLD = np.load("tests/golden/field_golden.npz")
EXACT = ["strip_rows", "strip_cols", "profile_len"]
SKIP = {"input_sha1"} | set(EXACT)
TOP_KEYS = {"top_position_index_x_y", "top_horizontal_distance_from_cax_mm", "top_vertical_distance_from_cax_mm",
            "top_horizontal_distance_from_beam_center_mm", "top_vertical_distance_from_beam_center_mm"}
TOL = 1e-7          # mm / samples / %: the restatement repeats the arithmetic; the "top" uses the exact parabola vertex


def oracle_kwargs(ak):
    ak = dict(ak)
    ak.pop("is_FFF", None)
    if "interpolation" in ak and ak["interpolation"] is not None:
        ak["interpolation"] = "Linear"
    return ak


@pytest.mark.parametrize("name", CASES)
def test_oracle_matches_reference_golden(name):
    a, ps, sid, ak = case_frame(name)
    sha = np.frombuffer(hashlib.sha1(a.tobytes()).digest(), dtype=np.uint8)
    assert np.array_equal(sha, GOLD[f"{name}/input_sha1"]), "synthetic input drifted from the one the golden was made with"
    dpmm = (25.4 / ps) / 25.4 * sid / 1000.0
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        o = field_oracle.field_analyze(a, dpmm, **oracle_kwargs(ak))
    for k in EXACT:
        assert np.array_equal(np.asarray(o[k]), GOLD[f"{name}/{k}"]), k
    keys = [k.split("/", 1)[1] for k in GOLD.files if k.startswith(name + "/") and k.split("/", 1)[1] not in SKIP]
    assert len(keys) >= 25
    for k in keys:
        if k in TOP_KEYS:
            continue
        np.testing.assert_allclose(np.asarray(o[k], dtype=float), GOLD[f"{name}/{k}"], rtol=0, atol=TOL, err_msg=k)
    # the "top": the reference's L-BFGS-B run stops wherever its finite-difference gradient noise lets it (tens of pixels from
    # the vertex on these flat tops, with the same function value to ~1e-9); ours is the exact vertex, so it can only be higher
    for axis, key in enumerate(("top_parabola_h", "top_parabola_v")):
        c2, c1, c0, xm, sc = o[key]
        poly = lambda x: c2 * ((x - xm) / sc) ** 2 + c1 * ((x - xm) / sc) + c0
        ours, ref = o["top_position_index_x_y"][axis], GOLD[f"{name}/top_position_index_x_y"][axis]
        assert poly(ours) >= poly(ref) - 1e-9
