"""GPU parity: per-image Winston-Lutz pipeline in CUDA (through the C-ABI) vs the committed reference goldens and the oracle.

Bars (BASELINE.json north_star): bit-exact for the integer quantities (cropped shape, inversion flag, crop count, threshold
passes); <= 0.01 px for the field CAX / BB centroids (we assert 1e-9 px: the field centre of mass is an exact integer ratio
and the weighted centroid differs from numpy's only in summation order).
The goldens come from the unmodified reference run with oracle/skimage_shim.py standing in for scikit-image (absent from the
build container), so the label / regionprops boundary is unpinned against skimage itself."""
import warnings

import numpy as np
import pytest

from tests.golden.wl_cases import CASES, case_frame

pytestmark = pytest.mark.gpu

GOLD = np.load("tests/golden/wl_golden.npz")
POS_TOL_PX = 1e-9     # required: 0.01 px


def gpu_run(name):
    from pylinac_b200 import winston_lutz as wl

    a, ps, sid, g, c, p, ak = case_frame(name)
    dpmm = (1 / ps) * sid / 1000.0
    return wl.analyze_batch(a[None], dpmm, **ak)[0], a, dpmm, ak


@pytest.mark.parametrize("name", CASES)
def test_wl2d_matches_reference_golden(name):
    r, a, dpmm, ak = gpu_run(name)
    if f"{name}/raises" in GOLD:
        assert r.status != 0
        with pytest.raises(ValueError):
            r.raise_for_status()
        return
    assert r.status == 0, r.status
    g = lambda k: GOLD[f"{name}/{k}"]
    assert np.array_equal(np.array(r.shape), g("shape"))
    np.testing.assert_allclose([r.field_cax.x, r.field_cax.y], g("field_cax"), rtol=0, atol=POS_TOL_PX)
    np.testing.assert_allclose([r.bb.x, r.bb.y], g("bb"), rtol=0, atol=POS_TOL_PX)
    np.testing.assert_allclose([r.epid.x, r.epid.y], g("epid"), rtol=0, atol=0)
    np.testing.assert_allclose([r.cax2bb_vector.x, r.cax2bb_vector.y], g("cax2bb_vector"), rtol=0, atol=POS_TOL_PX)
    np.testing.assert_allclose(r.cax2bb_distance, float(g("cax2bb_distance")), rtol=0, atol=POS_TOL_PX)
    np.testing.assert_allclose([r.cax2epid_vector.x, r.cax2epid_vector.y], g("cax2epid_vector"), rtol=0, atol=POS_TOL_PX)
    np.testing.assert_allclose(r.cax2epid_distance, float(g("cax2epid_distance")), rtol=0, atol=POS_TOL_PX)


@pytest.mark.parametrize("name", ["g0", "noisy", "inverted", "as1200", "bb8"])
def test_wl2d_integer_decisions_match_oracle(name):
    from oracle import wl_oracle

    r, a, dpmm, ak = gpu_run(name)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        o = wl_oracle.wl2d_analyze(a, dpmm, **ak)
    assert bool(r.r["inverted"]) == o["inverted"]
    assert int(r.r["crop_px"]) == 2 * o["crops"]
    assert int(r.r["threshold_passes"]) == o["threshold_passes"]


def test_wl2d_edge_cleanup_crops_like_the_oracle():
    """A bright 2-pixel artefact rim must be cropped ring by ring exactly as _clean_edges does (winston_lutz.py:1109-1133)."""
    from oracle import wl_oracle
    from pylinac_b200 import winston_lutz as wl

    a, ps, sid, g, c, p, ak = case_frame("g0")
    a = a.copy()
    a[:3, :] = 65535
    a[:, -4:] = 0
    dpmm = (1 / ps) * sid / 1000.0
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        o = wl_oracle.wl2d_analyze(a, dpmm)
    r = wl.analyze_batch(a[None], dpmm)[0]
    assert r.status == 0
    assert int(r.r["crop_px"]) == 2 * o["crops"] and o["crops"] >= 2
    assert np.array_equal(np.array(r.shape), o["shape"])
    np.testing.assert_allclose([r.bb.x, r.bb.y], o["bb"], rtol=0, atol=POS_TOL_PX)
    np.testing.assert_allclose([r.field_cax.x, r.field_cax.y], o["field_cax"], rtol=0, atol=POS_TOL_PX)


def test_wl2d_batch_is_per_frame_independent():
    from pylinac_b200 import winston_lutz as wl

    names = ["g0", "g90", "couch45", "big_offset", "fff"]
    frames, dp = [], None
    for nme in names:
        a, ps, sid, *_ = case_frame(nme)
        frames.append(a)
        dp = (1 / ps) * sid / 1000.0
    res = wl.analyze_batch(np.stack(frames), dp)
    for k, nme in enumerate(names):
        assert res[k].status == 0
        np.testing.assert_allclose([res[k].bb.x, res[k].bb.y], GOLD[f"{nme}/bb"], rtol=0, atol=POS_TOL_PX)


def test_wl2d_class_api():
    from pylinac_b200 import winston_lutz as wl

    a, ps, sid, g, c, p, ak = case_frame("couch45")
    img = wl.WinstonLutz2D(a, dpi=25.4 / ps, sid=sid, gantry=g, coll=c, couch=p)
    with pytest.raises(ValueError):
        img.results_data()
    img.analyze(**ak)
    rd = img.results_data()
    assert rd.variable_axis == str(GOLD["couch45/variable_axis"])
    np.testing.assert_allclose([img.bb.x, img.bb.y], GOLD["couch45/bb"], rtol=0, atol=1e-7)
    np.testing.assert_allclose(rd.cax2bb_distance, float(GOLD["couch45/cax2bb_distance"]), rtol=0, atol=1e-7)


@pytest.mark.parametrize("name", ["standard", "gantry_only", "combo"])
def test_wl_set_matches_reference_golden(name):
    """WinstonLutz (set level, winston_lutz.py:1519-1850): per-image rows from the GPU, set-level solve on the host."""
    from pylinac_b200 import winston_lutz as wl
    from tests.golden.wlset_cases import set_frames
    from tests.test_wlset_host import GOLD as SGOLD
    from tests.test_wlset_host import check_set

    frames, ps, sid, axes = set_frames(name)
    st = wl.WinstonLutz.from_arrays(frames, axes, dpmm=float(SGOLD[f"{name}/dpmm"]))
    with pytest.raises(ValueError):
        st.results_data()
    st.analyze()
    np.testing.assert_allclose([[im.bb.x, im.bb.y] for im in st.images], SGOLD[f"{name}/bbs"], rtol=0, atol=POS_TOL_PX)
    np.testing.assert_allclose([[im.field_cax.x, im.field_cax.y] for im in st.images], SGOLD[f"{name}/fields"], rtol=0, atol=POS_TOL_PX)
    check_set(st, name, 1e-7)      # scipy L-BFGS-B on a max-of-distances objective: inputs agree to 1e-9 px
