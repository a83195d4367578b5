"""GPU parity: Starshot pipeline in CUDA (through the C-ABI) vs the committed reference goldens and the oracle port.

Bars (BASELINE.json north_star): bit-exact for the integer quantities (start point, iteration count of the recursive search,
profile length, FWXM peak indices, line count, pass flag); <= 0.01 px for the wobble centre / radius (we assert 1e-6 px: the
fp64 profile arithmetic and the Nelder-Mead iteration repeat the reference's operation order)."""
import warnings

import numpy as np
import pytest

from tests.golden.starshot_cases import CASES, case_frame

pytestmark = pytest.mark.gpu

GOLD = np.load("tests/golden/starshot_golden.npz")
POS_TOL_PX = 1e-6     # required: 0.01 px


def gpu_run(name):
    from pylinac_b200 import starshot as ss

    a, ps, sid, ak = case_frame(name)
    dpmm = (1 / ps) * sid / 1000.0
    return ss.analyze_batch(a[None], dpmm, **ak)[0], dpmm


@pytest.mark.parametrize("name", CASES)
def test_starshot_matches_reference_golden(name):
    r, dpmm = gpu_run(name)
    if f"{name}/raises" in GOLD:
        assert r.status != 0
        with pytest.raises(RuntimeError):
            r.raise_for_status()
        return
    assert r.status == 0, r.status
    g = lambda k: GOLD[f"{name}/{k}"]
    row = r.r
    # ---- bit-exact integers
    assert int(row["iterations"]) == int(g("iterations"))
    assert int(row["profile_len"]) == int(g("profile_len"))
    assert int(row["n_lines"]) == int(g("n_lines"))
    npk = int(row["n_peaks"])
    assert np.array_equal(row["peak_idx"][:npk], g("peak_idx"))
    assert bool(row["passed"]) == bool(g("passed"))
    # ---- sub-pixel quantities
    np.testing.assert_allclose(float(row["radius_px"]), float(g("radius_px")), rtol=0, atol=1e-9)
    np.testing.assert_allclose(np.stack([row["peak_x"][:npk], row["peak_y"][:npk]], axis=1), g("peak_xy"), rtol=0, atol=POS_TOL_PX)
    np.testing.assert_allclose([float(row["wobble_x"]), float(row["wobble_y"])], g("wobble_center"), rtol=0, atol=POS_TOL_PX)
    np.testing.assert_allclose(float(row["wobble_radius_px"]), float(g("wobble_radius_px")), rtol=0, atol=POS_TOL_PX)
    np.testing.assert_allclose(float(row["wobble_radius_mm"]), float(g("wobble_radius_mm")), rtol=0, atol=POS_TOL_PX)
    np.testing.assert_allclose(row["angles"][: int(row["n_lines"])], g("angles"), rtol=0, atol=1e-6)


def test_starshot_start_point_and_inversion_match_oracle():
    from oracle import starshot_oracle

    for name in ("offset6", "inverted", "as1200"):
        a, ps, sid, ak = case_frame(name)
        dpmm = (1 / ps) * sid / 1000.0
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            o = starshot_oracle.starshot_analyze(a, dpmm, **ak)
        r, _ = gpu_run(name)
        assert bool(r.r["hist_inverted"]) == o["hist_inverted"]
        assert (int(r.r["start_x"]), int(r.r["start_y"])) == tuple(int(v) for v in o["start_point"])
        assert float(r.r["local_max"]) == o["local_max"]


def test_starshot_batch_is_frame_independent():
    from pylinac_b200 import starshot as ss

    names = ["offset6", "spokes4", "inverted", "noisy6"]
    frames = np.stack([case_frame(n)[0] for n in names])
    res = ss.analyze_batch(frames, 2.56)
    for i, n in enumerate(names):
        single, _ = gpu_run(n)
        for k in res.rows.dtype.names:
            np.testing.assert_array_equal(res.rows[k][i], single.r[k], err_msg=f"{n}/{k}")


def test_starshot_class_api():
    from pylinac_b200.starshot import Starshot

    a, ps, sid, ak = case_frame("offset6")
    s = Starshot(a, dpi=25.4 / ps, sid=sid)
    s.analyze()
    assert s.passed
    rd = s.results_data()
    np.testing.assert_allclose(rd.circle_center_x_y, GOLD["offset6/wobble_center"], atol=POS_TOL_PX)
    assert abs(rd.circle_radius_mm - float(GOLD["offset6/wobble_radius_mm"])) < 1e-6
    assert len(rd.angles) == 6 and len(s.lines) == 6
    assert "Starshot Results" in s.results()
    with pytest.raises(RuntimeError):
        b, ps, sid, ak = case_frame("wide_wobble")
        Starshot(b, dpi=25.4 / ps, sid=sid).analyze(**ak)
