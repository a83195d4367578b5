"""GPU parity: FieldProfileAnalysis (pylinac_b200/field_profile_analysis.py: GPU strip sums / peak search / smoothing, host
metric formulas) against goldens from the UNMODIFIED reference (tests/golden/make_fpa_golden.py).

FWHM edge type: every quantity to 1e-9 (the same arithmetic; the strip sums are exact integers).  INFLECTION_DERIVATIVE: the
reference ends a BFGS minimisation of the cubic interpolant at its gradient tolerance, we evaluate the spline's stationary point
in closed form, so the edges agree to 1e-5 .. 2e-4 samples (bar: 0.01 px; the looser end is the BFGS tolerance on normalised profiles) and the metrics that depend on them accordingly."""
import numpy as np
import pytest

from tests.golden.fpa_cases import CASES, case, resolve_enums

pytestmark = pytest.mark.gpu

GOLD = np.load("tests/golden/fpa_golden.npz")


def run(name):
    from pylinac_b200.field_profile_analysis import FieldProfileAnalysis

    a, ps, sid, kw = case(name)
    from pylinac_b200.core.profile import Normalization

    f = FieldProfileAnalysis(a, dpi=25.4 / ps, sid=sid)
    f.analyze(**resolve_enums(kw, Normalization))
    return f, kw


@pytest.mark.parametrize("name", list(CASES))
def test_field_profile_analysis_matches_reference(name):
    f, kw = run(name)
    exact = kw.get("edge_type") == "FWHM"
    edge_tol = 1e-9 if exact else 1e-3          # samples (required: 0.01)
    for ax, p in (("x", f.x_profile), ("y", f.y_profile)):
        g = lambda k: GOLD[f"{name}/{ax}/{k}"]
        np.testing.assert_allclose(p.values, g("values"), rtol=0, atol=1e-9 if exact else 1e-6, err_msg="profile values")
        np.testing.assert_allclose([p.field_edge_idx("left"), p.field_edge_idx("right")], g("edges"), rtol=0, atol=edge_tol)
        np.testing.assert_allclose(p.center_idx, float(g("center_idx")), rtol=0, atol=edge_tol)
        np.testing.assert_allclose(p.cax_index, float(g("cax_index")), rtol=0, atol=0)
        np.testing.assert_allclose(p.field_width_mm, float(g("field_width_mm")), rtol=0, atol=edge_tol)
        assert list(p.metric_values.keys()) == list(g("metric_names"))
        for key, got, exp in zip(p.metric_values.keys(), p.metric_values.values(), g("metric_values")):
            np.testing.assert_allclose(float(got), float(exp), rtol=0, atol=1e-9 if exact else 1e-3, err_msg=f"{ax} {key}")


def test_field_profile_analysis_api():
    from pylinac_b200.field_profile_analysis import FieldProfileAnalysis, NotAnalyzed
    from pylinac_b200.metrics.profile import FlatnessRatioMetric, SymmetryAreaMetric, SymmetryPointDifferenceQuotientMetric

    a, ps, sid, _ = case("fwhm")
    f = FieldProfileAnalysis(a, dpi=25.4 / ps, sid=sid)
    with pytest.raises(NotAnalyzed):
        f.results_data()
    with pytest.raises(NotImplementedError):
        f.analyze(edge_type="Inflection Hill")
    with pytest.raises(ValueError):
        f.analyze(edge_type="FWHM", x_width=1.5)
    f.analyze(edge_type="FWHM", metrics=(FlatnessRatioMetric(), SymmetryAreaMetric(), SymmetryPointDifferenceQuotientMetric()))
    rd = f.results_data()
    assert set(rd.x_metrics) == {"Flatness (Ratio) (%)", "Symmetry (Area)", "Point Difference Quotient Symmetry (%)", "Field Width (mm)", "values"}
    assert 100 < rd.x_metrics["Flatness (Ratio) (%)"] < 110 and abs(rd.x_metrics["Symmetry (Area)"]) < 1
    assert "x_metrics" in f.results()
