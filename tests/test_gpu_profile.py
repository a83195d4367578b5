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
    # Hill edges -> beam centre -> normalisation value: the fit tolerance propagates into the normalised values
    np.testing.assert_allclose(sp.values, g("values"), rtol=0, atol=1e-8 if f"{name}/hill" in GOLD and kw.get("normalization_method", 1) is not None else 1e-12)
    np.testing.assert_allclose(sp.x_indices, g("x_indices"), rtol=0, atol=1e-12)
    gc, bc = sp.geometric_center(), sp.beam_center()
    np.testing.assert_allclose([gc["index (exact)"], gc["value (exact)"]], g("geometric_center"), rtol=0, atol=TOL)
    np.testing.assert_allclose([bc["index (exact)"], bc["value (@rounded)"]], g("beam_center"), rtol=0,
                               atol=2e-6 if f"{name}/hill" in GOLD else TOL)
    fw = sp.fwxm_data(q["fwxm_x"])
    np.testing.assert_allclose([fw["left index (exact)"], fw["right index (exact)"], fw["center value (@rounded)"],
                                fw["left value (@rounded)"], fw["right value (@rounded)"]], g("fwxm"), rtol=0, atol=TOL)
    lo, up = q["penumbra"]
    pen = sp.penumbra(lo, up)
    np.testing.assert_allclose([pen[f"left {lo}% index (exact)"], pen[f"left {up}% index (exact)"], pen[f"right {lo}% index (exact)"],
                                pen[f"right {up}% index (exact)"]], g("penumbra"), rtol=0, atol=2e-6 if f"{name}/hill" in GOLD else TOL)
    if f"{name}/hill" in GOLD:
        # Edge.INFLECTION_HILL: the reference's curve_fit and the host Levenberg-Marquardt stop at the same least-squares minimum
        # within their 1.5e-8 tolerances (tests/test_hill_host.py); positions agree to ~1e-6, parameters to ~1e-5 relative
        inf = sp.inflection_data()
        np.testing.assert_allclose([inf["left index (exact)"], inf["right index (exact)"], inf["left value (@exact)"],
                                    inf["right value (@exact)"]], g("hill"), rtol=0, atol=2e-6)
        np.testing.assert_allclose(np.array([inf["left Hill params"], inf["right Hill params"]]), g("hill_params"), rtol=2e-5)
        np.testing.assert_allclose([pen["left gradient (exact)"], pen["right gradient (exact)"]], g("hill_gradients"), rtol=2e-5)
    if f"{name}/inflection" in GOLD:
        inf = sp.inflection_data()
        np.testing.assert_allclose([inf["left index (exact)"], inf["right index (exact)"], inf["left value (@exact)"],
                                    inf["right value (@exact)"], inf["left value (@rounded)"], inf["right value (@rounded)"]],
                                   g("inflection"), rtol=0, atol=TOL)
    fd = sp.field_data(q["in_field_ratio"], q["slope_exclusion_ratio"])
    hill = f"{name}/hill" in GOLD
    np.testing.assert_allclose([float(fd[k]) for k in FD_KEYS], g("field_data"), rtol=0, atol=2e-6 if hill else 1e-8)
    np.testing.assert_allclose(fd["field values"], g("field_values"), rtol=0, atol=1e-8 if hill else 1e-12)
    # np.polyfit coefficients: same least-squares problem, solved on centred / scaled abscissae
    x = np.linspace(fd["left inner index (exact)"], fd["right inner index (exact)"], 50)
    ours = np.polyval(fd["top params"], x)
    ref = np.polyval(g("top_params"), x)
    np.testing.assert_allclose(ours, ref, rtol=0, atol=1e-9)
    assert abs(sp.field_calculation(q["in_field_ratio"], "max", q["slope_exclusion_ratio"]) - g("field_values").max()) < (1e-8 if hill else 1e-12)


def test_single_profile_rejects_what_the_gpu_path_does_not_cover():
    from pylinac_b200.core.profile import Edge, Interpolation, SingleProfile

    vals, _, _ = case_profile("default")
    with pytest.raises(ValueError, match="monotonically increasing"):
        SingleProfile(vals, x_values=np.arange(len(vals))[::-1].copy())
    sp = SingleProfile(vals)
    with pytest.raises(ValueError):
        sp.field_data(0.5, 0.6)
    with pytest.raises(ValueError):
        sp.penumbra(80, 20)
    with pytest.raises(ValueError):
        sp.inflection_data()


REG = np.load("tests/golden/profile_regression.npz")


@pytest.mark.parametrize("variant", ["x", "linear_x", "spline_x", "no_x", "linear_no_x", "spline_no_x"])
@pytest.mark.parametrize("k", range(len(REG["names"])))
def test_single_profile_matches_the_reference_frozen_regressions(k, variant):
    """The reference's own known answers (tests_basic/core/profile_regression_fixtures.py, pinned to 1e-9 by
    tests_basic/core/test_profile.py:2546-2687): protocol metrics of SingleProfile(values[, x_values]) for interpolation NONE /
    LINEAR / SPLINE, with the exported (partly unevenly spaced) detector positions and without."""
    from pylinac_b200 import field_analysis as fa
    from pylinac_b200.core.profile import Interpolation, SingleProfile

    calc = {"varian_flatness_difference": fa.flatness_dose_difference, "varian_symmetry_point_difference": fa.symmetry_point_difference,
            "elekta_flatness_ratio": fa.flatness_dose_ratio, "elekta_symmetry_pdq": fa.symmetry_pdq_iec,
            "siemens_flatness_difference": fa.flatness_dose_difference, "siemens_symmetry_area": fa.symmetry_area}
    interp = {"x": Interpolation.NONE, "no": Interpolation.NONE, "linear": Interpolation.LINEAR, "spline": Interpolation.SPLINE}[variant.split("_")[0]]
    p = SingleProfile(REG[f"{k}/values"], x_values=None if variant.endswith("no_x") else REG[f"{k}/x_values"], interpolation=interp)
    for key, exp in zip(REG[f"{k}/{variant}/keys"], REG[f"{k}/{variant}/vals"]):
        got = calc[str(key)](p, in_field_ratio=0.8)
        assert abs(got - exp) <= 1e-9, (str(REG["names"][k]), str(key), got, exp)
    if variant == "x":          # test_field_data_geometry_matches_frozen_exports (delta 1e-4 in the reference)
        fd = p.field_data(in_field_ratio=0.8, slope_exclusion_ratio=0.2)
        for key, exp in zip(REG[f"{k}/field_data/keys"], REG[f"{k}/field_data/vals"]):
            tol = 1e-4
            if str(key) == '"top" index (exact)':
                # the frozen value is where the reference's L-BFGS-B run stopped: anywhere the slope of the fitted parabola is below
                # its projected-gradient tolerance 1e-5, i.e. within 1e-5 / (2 |p0|) of the optimum this engine returns
                tol = max(tol, 1.5e-5 / (2 * abs(float(fd["top params"][0]))))
            assert abs(fd[str(key)] - exp) <= tol, (str(REG["names"][k]), str(key), fd[str(key)], exp, tol)


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
