"""CUDA VMAT (DRGS / DRMLC / DRCS) and DLG against goldens of the UNMODIFIED reference (tests/golden/make_vmat_golden.py) and
against oracle/vmat_oracle.py on seeded random pairs.  Tolerances: integer / boolean results exact; R_corr / R_dev / stdev are fp64
means of per-pixel quotients, summed in a different order than numpy's pairwise sum -> 1e-9 relative."""
import warnings

import numpy as np
import pytest

from tests.golden.vmat_cases import DLG_CASES, DRCS_CASES, VMAT_CASES, dlg_case, drcs_case, vmat_case

pytestmark = pytest.mark.gpu
GOLD = np.load("tests/golden/vmat_golden.npz", allow_pickle=False)
RTOL = 1e-9


def _check_segments(name, segs, v):
    g = lambda k: GOLD[f"{name}/{k}"]
    np.testing.assert_allclose([s.r_corr for s in segs], g("r_corr"), rtol=RTOL, atol=1e-9, equal_nan=True)
    np.testing.assert_allclose([s.r_dev for s in segs], g("r_dev"), rtol=0, atol=1e-7, equal_nan=True)
    np.testing.assert_allclose([s.stdev for s in segs], g("stdev"), rtol=1e-7, atol=1e-12, equal_nan=True)
    np.testing.assert_allclose([s.center.x for s in segs], g("center_x"), rtol=0, atol=1e-9)
    np.testing.assert_allclose([s.center.y for s in segs], g("center_y"), rtol=0, atol=1e-9)
    assert [bool(s.passed) for s in segs] == [bool(x) for x in g("passed_seg")]
    assert bool(v.passed) == bool(g("passed"))
    np.testing.assert_allclose(v.max_r_deviation, g("max_r_deviation"), rtol=0, atol=1e-7, equal_nan=True)
    np.testing.assert_allclose(v.avg_abs_r_deviation, g("avg_abs_r_deviation"), rtol=0, atol=1e-7, equal_nan=True)
    np.testing.assert_allclose(v.avg_r_deviation, g("avg_r_deviation"), rtol=0, atol=1e-7, equal_nan=True)


@pytest.mark.parametrize("name", VMAT_CASES)
def test_vmat_class_matches_reference_golden(name):
    from pylinac_b200 import vmat

    klass, a, b, ps, sid, ck, ak = vmat_case(name)
    with warnings.catch_warnings(record=True) as wl:
        warnings.simplefilter("always")
        v = getattr(vmat, klass)(image_paths=(a, b), dpi=25.4 / ps, sid=sid, **ck)
        v.analyze(**ak)
    g = lambda k: GOLD[f"{name}/{k}"]
    assert int(v.open_image is v._images[0]) == int(g("open_is_first"))
    _check_segments(name, v.segments, v)
    n_warn = sum(1 for w in wl if "VMAT field center" in str(w.message))
    assert (n_warn > 0) == bool(int(g("n_user_warnings")))
    # the image objects carry the reference's ground / inversion (sums of the processed arrays)
    np.testing.assert_allclose(np.asarray(v.open_image.array, dtype=np.float64).sum(), g("open_sum"), rtol=0, atol=0.5)
    np.testing.assert_allclose(np.asarray(v.dmlc_image.array, dtype=np.float64).sum(), g("dmlc_sum"), rtol=0, atol=0.5)
    rd = v.results_data()
    np.testing.assert_allclose(rd.max_deviation_percent, g("rd_max_deviation_percent"), atol=1e-7, equal_nan=True)
    assert len(rd.segment_data) == len(v.segments) and rd.test_type == v._result_header
    if bool(int(g("n_user_warnings"))):
        assert any("VMAT field center" in w["message"] for w in rd.warnings)


def test_vmat_batch_mixed_pairs_match_goldens_and_oracle():
    """All default-configuration 1280 x 1280 DRGS pairs in ONE batched call (either image order), each row against its golden."""
    from pylinac_b200 import vmat

    names = ["drgs_7", "drgs_inverted_swapped", "drgs_failing"]
    cases = [vmat_case(n) for n in names]
    i1 = np.stack([c[1] for c in cases])
    i2 = np.stack([c[2] for c in cases])
    dpmm = (1 / cases[0][3]) * cases[0][4] / 1000.0
    rows = vmat.analyze_batch(i1, i2, dpmm, test="DRGS")
    for name, r in zip(names, rows):
        r.raise_for_status()
        g = lambda k: GOLD[f"{name}/{k}"]
        assert int(r.open_is_first) == int(g("open_is_first"))
        np.testing.assert_allclose(r.r_corrs, g("r_corr"), rtol=RTOL)
        np.testing.assert_allclose(r.r_devs, g("r_dev"), atol=1e-7)
        np.testing.assert_allclose(r.stdevs, g("stdev"), rtol=1e-7)
        assert r.passed == bool(g("passed"))


@pytest.mark.parametrize("seed", range(6))
def test_vmat_random_pairs_match_oracle(seed):
    from oracle import synth, vmat_oracle
    from pylinac_b200 import vmat

    rng = np.random.default_rng(7000 + seed)
    fr = [synth.as1200(1000.0), synth.epid1024(), synth.as1000(1500.0)][seed % 3]
    o = type(fr)(fr.shape, fr.pixel_size, fr.sid)
    d = type(fr)(fr.shape, fr.pixel_size, fr.sid)
    test = "DRGS" if seed % 2 == 0 else "DRMLC"
    offs = (-60, -40, -20, 0, 20, 40, 60) if test == "DRGS" else (-45, -15, 15, 45)
    shift = float(rng.uniform(-6, 6))
    o.add_filtered_field((150, 150), cax_offset_mm=(0, shift), alpha=float(rng.uniform(0.5, 0.8)))
    o.gaussian(2.0)
    o.noise(0.002, seed=seed)
    for k in offs:
        d.add_filtered_field((150, 16 if test == "DRGS" else 26), cax_offset_mm=(0, shift + k), alpha=float(rng.uniform(0.28, 0.32)))
    d.gaussian(1.5)
    d.noise(0.002, seed=100 + seed)
    a, b = (o.image, d.image) if seed % 3 else (d.image, o.image)
    if seed == 4:
        a, b = o.inverted(), d.inverted()
    dpmm = o.dpmm
    tol = float(rng.uniform(1.0, 4.0))
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        want = vmat_oracle.vmat_analyze(a, b, dpmm, offsets_mm=offs, tolerance=tol)
    r = vmat.analyze_batch(a[None], b[None], dpmm, test=test, tolerance=tol)[0]
    r.raise_for_status()
    assert int(r.open_is_first) == want["open_is_first"]
    assert [bool(x) for x in r.r["inverted"]] == [bool(x) for x in want["inverted"]]
    np.testing.assert_allclose(r.r["profile_center_idx"], want["profile_center_idx"], atol=1e-9)
    assert list(r.r["field_len"]) == list(want["field_len"])
    np.testing.assert_allclose(r.r["field_std"], want["field_std"], rtol=1e-9)
    np.testing.assert_allclose(r.r_corrs, want["r_corr"], rtol=RTOL)
    np.testing.assert_allclose(r.r_devs, want["r_dev"], atol=1e-7)
    np.testing.assert_allclose(r.stdevs, want["stdev"], rtol=1e-7)
    assert r.passed == want["passed"]


@pytest.mark.parametrize("name", DRCS_CASES)
def test_drcs_matches_reference_golden(name):
    from pylinac_b200 import vmat

    klass, a, b, ps, sid, ck, ak = drcs_case(name)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        v = vmat.DRCS(image_paths=(a, b), dpi=25.4 / ps, sid=sid, **ck)
        v.analyze(**ak)
    g = lambda k: GOLD[f"{name}/{k}"]
    assert int(v.open_image is v._images[0]) == int(g("open_is_first"))
    _check_segments(name, v.segments, v)
    np.testing.assert_allclose([s.rotation for s in v.segments], g("rotation"), atol=1e-9)
    np.testing.assert_allclose([cd.angle_measured for cd in v.collimator_deviations], g("coll_angle_measured"), atol=1e-9)
    np.testing.assert_allclose([cd.angle_deviation for cd in v.collimator_deviations], g("coll_angle_deviation"), atol=1e-9)
    np.testing.assert_allclose(v.rotation_offset_deg, g("rotation_offset_deg"), atol=1e-9)
    rd = v.results_data()
    assert len(rd.collimator_data) == len(v.collimator_deviations)


def test_drcs_too_many_spokes_raises():
    from pylinac_b200 import vmat

    klass, a, b, ps, sid, ck, ak = drcs_case("drcs_basic")
    v = vmat.DRCS(image_paths=(a, b), dpi=25.4 / ps, sid=sid)
    cfg = dict(v.default_collimator_config)
    cfg["G"] = 0
    with pytest.raises(ValueError):
        v.analyze(collimator_config=cfg)


@pytest.mark.parametrize("name", DLG_CASES)
def test_dlg_matches_reference_golden(name):
    from pylinac_b200 import dlg
    from pylinac_b200.core import image
    from pylinac_b200.picketfence import MLC

    frame, ps, sid, gaps, mlc, yfs, pw = dlg_case(name)
    img = image.ArrayImage(frame, dpi=25.4 / ps, sid=sid)
    d = dlg.DLG(img)
    d.analyze(gaps=gaps, mlc=MLC[mlc], y_field_size=yfs, profile_width=pw)
    g = lambda k: GOLD[f"{name}/{k}"]
    np.testing.assert_allclose(d.measured_dlg_per_leaf, g("measured_dlg_per_leaf"), rtol=1e-12, atol=1e-9)
    np.testing.assert_allclose(d.planned_dlg_per_leaf, g("planned_dlg_per_leaf"), rtol=0, atol=0)
    np.testing.assert_allclose(d.measured_dlg, g("measured_dlg"), rtol=1e-10)
    np.testing.assert_allclose(d._lin_fit.slope, g("slope"), rtol=1e-10)
    np.testing.assert_allclose(d._lin_fit.intercept, g("intercept"), rtol=1e-10)
    # batched: the same frame three times
    dpmm = (1 / ps) * sid / 1000.0
    out = dlg.analyze_batch(np.stack([frame] * 3), dpmm, gaps, MLC[mlc], yfs, pw)
    np.testing.assert_allclose(out["measured_dlg"], np.full(3, g("measured_dlg")), rtol=1e-10)
