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
    # Inflection Hill: derivative edges (closed-form stationary point, see above) seed the fit windows; the two least-squares solvers
    # agree to ~1e-8 on a given window (tests/test_hill_host.py)
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
    # central ROI (center_rect): mean / std / min / max of the rasterised rectangle (one epid_roi_stats launch)
    c = f.center_rect
    np.testing.assert_allclose([c.mean, c.std, c.min, c.max], GOLD[f"{name}/center"], rtol=1e-12, atol=1e-9)
    assert set(f.results_data().center) == {"mean", "stdev", "min", "max"}


def test_field_profile_analyze_batch_equals_per_frame():
    """analyze_batch: whole-batch device statistics (inversion check, centre sums, strips shared per view) == the per-frame class"""
    from pylinac_b200 import field_profile_analysis as fpa

    names = ["default_inflection", "inverted_frame", "fwhm", "offset_inflection_wide"]
    frames, ps, sid = [], None, None
    for n in names:
        a, ps, sid, _ = case(n)
        frames.append(a)
    dpmm = 1 / ps * sid / 1000.0
    stack = np.stack(frames)
    for kw in ({}, {"edge_type": "FWHM", "x_width": 0.02, "y_width": 0.03}, {"invert": True, "centering": "Manual", "edge_type": "FWHM"}):
        singles, err = [], None
        for a in frames:
            one = fpa.FieldProfileAnalysis(a, dpi=25.4 / ps, sid=sid)
            try:
                one.analyze(**kw)
            except (IndexError, ValueError) as e:       # e.g. no peak in a flipped profile: the batch path must fail the same way
                err = type(e)
                break
            singles.append(one)
        if err is not None:
            with pytest.raises(err):
                fpa.analyze_batch(stack, dpmm, sid=sid, **kw)
            continue
        batch = fpa.analyze_batch(stack, dpmm, sid=sid, **kw)
        for got, one in zip(batch, singles):
            for ax in ("x_profile", "y_profile"):
                a_, b_ = getattr(got, ax), getattr(one, ax)
                np.testing.assert_array_equal(a_.values, b_.values)
                assert a_.metric_values == b_.metric_values
            assert (got.center_rect.mean, got.center_rect.std) == (one.center_rect.mean, one.center_rect.std)
