"""GPU parity of pylinac_b200.core.profile.SingleProfile (the device engine of csrc/field.cu through epid_single_profile)
against golden vectors produced by the UNMODIFIED reference SingleProfile (tests/golden/make_profile_golden.py)."""
import numpy as np
import pytest

from tests.golden.make_profile_golden import FD_KEYS
from tests.golden.profile_cases import CASES, case_profile

pytestmark = pytest.mark.gpu

GOLD = np.load("tests/golden/profile_golden.npz")
TOL = 1e-9


def build(name):
    from pylinac_b200.core.profile import Centering, Edge, Interpolation, Normalization, SingleProfile

    vals, kw, q = case_profile(name)
    kw = dict(kw)
    for k, enum in (("interpolation", Interpolation), ("normalization_method", Normalization), ("edge_detection_method", Edge),
                    ("centering", Centering)):
        if k in kw:
            kw[k] = enum(kw[k])
    return SingleProfile(vals, **kw), kw, q


@pytest.mark.parametrize("name", CASES)
def test_single_profile_matches_reference(name):
    sp, kw, q = build(name)
    g = lambda k: GOLD[f"{name}/{k}"]
    assert len(sp.values) == len(g("values"))
    np.testing.assert_allclose(sp.values, g("values"), rtol=0, atol=1e-12)
    np.testing.assert_allclose(sp.x_indices, g("x_indices"), rtol=0, atol=1e-12)
    gc, bc = sp.geometric_center(), sp.beam_center()
    np.testing.assert_allclose([gc["index (exact)"], gc["value (exact)"]], g("geometric_center"), rtol=0, atol=TOL)
    np.testing.assert_allclose([bc["index (exact)"], bc["value (@rounded)"]], g("beam_center"), rtol=0, atol=TOL)
    fw = sp.fwxm_data(q["fwxm_x"])
    np.testing.assert_allclose([fw["left index (exact)"], fw["right index (exact)"], fw["center value (@rounded)"],
                                fw["left value (@rounded)"], fw["right value (@rounded)"]], g("fwxm"), rtol=0, atol=TOL)
    lo, up = q["penumbra"]
    pen = sp.penumbra(lo, up)
    np.testing.assert_allclose([pen[f"left {lo}% index (exact)"], pen[f"left {up}% index (exact)"], pen[f"right {lo}% index (exact)"],
                                pen[f"right {up}% index (exact)"]], g("penumbra"), rtol=0, atol=TOL)
    if f"{name}/inflection" in GOLD:
        inf = sp.inflection_data()
        np.testing.assert_allclose([inf["left index (exact)"], inf["right index (exact)"], inf["left value (@exact)"],
                                    inf["right value (@exact)"], inf["left value (@rounded)"], inf["right value (@rounded)"]],
                                   g("inflection"), rtol=0, atol=TOL)
    fd = sp.field_data(q["in_field_ratio"], q["slope_exclusion_ratio"])
    np.testing.assert_allclose([float(fd[k]) for k in FD_KEYS], g("field_data"), rtol=0, atol=1e-8)
    np.testing.assert_allclose(fd["field values"], g("field_values"), rtol=0, atol=1e-12)
    # np.polyfit coefficients: same least-squares problem, solved on centred / scaled abscissae
    x = np.linspace(fd["left inner index (exact)"], fd["right inner index (exact)"], 50)
    ours = np.polyval(fd["top params"], x)
    ref = np.polyval(g("top_params"), x)
    np.testing.assert_allclose(ours, ref, rtol=0, atol=1e-9)
    assert abs(sp.field_calculation(q["in_field_ratio"], "max", q["slope_exclusion_ratio"]) - g("field_values").max()) < 1e-12


def test_single_profile_rejects_what_the_gpu_path_does_not_cover():
    from pylinac_b200.core.profile import Edge, Interpolation, SingleProfile

    vals, _, _ = case_profile("default")
    with pytest.raises(NotImplementedError):
        SingleProfile(vals, interpolation=Interpolation.SPLINE)
    with pytest.raises(NotImplementedError):
        SingleProfile(vals, edge_detection_method=Edge.INFLECTION_HILL)
    sp = SingleProfile(vals)
    with pytest.raises(ValueError):
        sp.field_data(0.5, 0.6)
    with pytest.raises(ValueError):
        sp.penumbra(80, 20)
    with pytest.raises(ValueError):
        sp.inflection_data()


REG = np.load("tests/golden/profile_regression.npz")


@pytest.mark.parametrize("variant", ["no_x", "linear_no_x"])
@pytest.mark.parametrize("k", range(len(REG["names"])))
def test_single_profile_matches_the_reference_frozen_regressions(k, variant):
    """The reference's own known answers (tests_basic/core/profile_regression_fixtures.py, pinned to 1e-9 by
    tests_basic/core/test_profile.py:2546-2687): protocol metrics of SingleProfile(values) without x_values."""
    from pylinac_b200 import field_analysis as fa
    from pylinac_b200.core.profile import Interpolation, SingleProfile

    calc = {"varian_flatness_difference": fa.flatness_dose_difference, "varian_symmetry_point_difference": fa.symmetry_point_difference,
            "elekta_flatness_ratio": fa.flatness_dose_ratio, "elekta_symmetry_pdq": fa.symmetry_pdq_iec,
            "siemens_flatness_difference": fa.flatness_dose_difference, "siemens_symmetry_area": fa.symmetry_area}
    p = SingleProfile(REG[f"{k}/values"], interpolation=Interpolation.NONE if variant == "no_x" else Interpolation.LINEAR)
    for key, exp in zip(REG[f"{k}/{variant}/keys"], REG[f"{k}/{variant}/vals"]):
        got = calc[str(key)](p, in_field_ratio=0.8)
        assert abs(got - exp) <= 1e-9, (str(REG["names"][k]), str(key), got, exp)


# ---- the reference's toy-profile known answers (tests_basic/core/test_profile.py:104-161, 215-330, 2700-2722)
SIMPLE9 = np.array([0, 1, 2, 3, 4, 3, 2, 1, 0], dtype=float)
SIMPLE8 = np.array([0, 1, 2, 3, 3, 2, 1, 0], dtype=float)
SKEWED19 = np.array([0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 10, 10, 8, 6, 4, 2, 0], dtype=float)


@pytest.mark.parametrize("values,height,left,right", [
    (SIMPLE9, 50, 2, 6), (SIMPLE8, 50, 1.5, 5.5), (SIMPLE9, 25, 1, 7), (SIMPLE9, 75, 3, 5), (SKEWED19, 50, 5, 14.5)])
def test_fwxm_profile_known_edges(values, height, left, right):
    from pylinac_b200.core.profile import FWXMProfile

    p = FWXMProfile(values, fwxm_height=height)
    assert p.field_edge_idx("left") == left
    assert p.field_edge_idx("right") == right
    assert p.field_width_px == right - left


def test_fwxm_profile_known_centres():
    from pylinac_b200.core.profile import FWXMProfile

    assert FWXMProfile(SIMPLE9).center_idx == 4
    assert FWXMProfile(SIMPLE8).center_idx == 3.5


def test_multiprofile_triangle_known_peaks():
    """MultiProfileTriangle (tests_basic/core/test_profile.py:2716-2722): peaks / valleys / FWXM peaks within +-1 sample."""
    from scipy import signal as sps

    from pylinac_b200.core.profile import MultiProfile

    p = MultiProfile(sps.sawtooth(np.linspace(0, 8 * np.pi, num=200), width=0.5))
    for got, known in ((p.find_peaks()[0], (25, 75, 125, 175)), (p.find_valleys()[0], (50, 100, 150)),
                       (p.find_fwxm_peaks()[0], (25, 75, 125, 175))):
        assert len(got) == len(known)
        assert all(abs(int(a) - b) <= 1 for a, b in zip(got, known))
    assert p.values.min() != 0
    p.ground()
    assert p.values.min() == 0
