"""CPU: the set-level Winston-Lutz host logic (3-D BB / field position solve, isocentre sizes, axis classification, distance
statistics; pylinac_b200/winston_lutz.py) against goldens produced by the UNMODIFIED reference WinstonLutz
(tests/golden/make_wlset_golden.py).  The per-image inputs (BB / field / EPID points) are taken from the golden file here, so
this test needs no GPU; tests/test_gpu_wl.py runs the same comparison with the per-image rows computed in CUDA."""
import numpy as np
import pytest

from pylinac_b200 import _native as nat
from pylinac_b200 import winston_lutz as wl
from tests.golden.make_wlset_golden import SCALARS
from tests.golden.wlset_cases import SETS

GOLD = np.load("tests/golden/wlset_golden.npz")


def rows_from_golden(name):
    bbs, fields, epids, dpmm = GOLD[f"{name}/bbs"], GOLD[f"{name}/fields"], GOLD[f"{name}/epids"], float(GOLD[f"{name}/dpmm"])
    rows = np.zeros(len(bbs), nat.WL_RESULT_DTYPE)
    rows["bb_x"], rows["bb_y"] = bbs[:, 0], bbs[:, 1]
    rows["field_x"], rows["field_y"] = fields[:, 0], fields[:, 1]
    rows["epid_x"], rows["epid_y"] = epids[:, 0], epids[:, 1]
    rows["cax2bb_x"], rows["cax2bb_y"] = (bbs[:, 0] - fields[:, 0]) / dpmm, (bbs[:, 1] - fields[:, 1]) / dpmm
    rows["cax2bb_distance"] = np.sqrt((fields[:, 0] - bbs[:, 0]) ** 2 + (fields[:, 1] - bbs[:, 1]) ** 2 + 0.0) / dpmm
    rows["cax2epid_x"], rows["cax2epid_y"] = (epids[:, 0] - fields[:, 0]) / dpmm, (epids[:, 1] - fields[:, 1]) / dpmm
    rows["cax2epid_distance"] = np.sqrt((fields[:, 0] - epids[:, 0]) ** 2 + (fields[:, 1] - epids[:, 1]) ** 2 + 0.0) / dpmm
    return rows, dpmm


def build_set(name, rows, dpmm):
    st = wl.WinstonLutz.__new__(wl.WinstonLutz)
    st._setup(np.zeros((len(rows), 4, 4), np.uint16), [tuple(float(v) for v in a) for a in SETS[name][1]], dpmm)
    refs = dict(snap_tolerance=3, gantry_reference=0, collimator_reference=0, couch_reference=0)
    st.images = [wl._SetImage(wl.WLFrameResult(rows[k]), dpmm, *st._axes[k], refs) for k in range(len(rows))]
    st._is_analyzed = True
    return st


def check_set(st, name, tol):
    rd = st.results_data()
    for k in SCALARS:
        np.testing.assert_allclose(getattr(rd, k), GOLD[f"{name}/{k}"], rtol=0, atol=tol, err_msg=k)
    sv = st.bb_shift_vector
    np.testing.assert_allclose([sv.x, sv.y, sv.z], GOLD[f"{name}/bb_shift_vector"], rtol=0, atol=tol)
    p = st.measured_bb_position
    np.testing.assert_allclose([p.x, p.y, p.z], GOLD[f"{name}/measured_bb_position"], rtol=0, atol=tol)
    p = st.measured_field_position
    np.testing.assert_allclose([p.x, p.y, p.z], GOLD[f"{name}/measured_field_position"], rtol=0, atol=tol)
    assert [im.variable_axis.value for im in st.images] == list(GOLD[f"{name}/variable_axes"])
    assert list(rd.keyed_image_details.keys()) == list(GOLD[f"{name}/keys"])
    assert st.bb_shift_instructions() == str(GOLD[f"{name}/shift_instructions"])


@pytest.mark.parametrize("name", list(SETS))
def test_set_level_solve_matches_reference(name):
    rows, dpmm = rows_from_golden(name)
    check_set(build_set(name, rows, dpmm), name, 1e-9)


def test_scale_conversion_and_axis_snap():
    assert wl.convert_scale(wl.MachineScale.IEC61217, wl.MachineScale.VARIAN_STANDARD, 90, 0, 45) == (90, 180, 135)
    assert wl.convert_scale(wl.MachineScale.VARIAN_IEC, wl.MachineScale.IEC61217, 10, 20, 30) == (10, 20, 330)
    assert wl.variable_axis(358, 0, 0) == wl.Axis.REFERENCE
    assert wl.variable_axis(358, 0, 0, snap_tolerance=1) == wl.Axis.GANTRY
    assert wl.variable_axis(0, 45, 0, collimator_reference=45) == wl.Axis.REFERENCE
    assert wl.variable_axis(10, 10, 10) == wl.Axis.GBP_COMBO
    with pytest.raises(ValueError):
        wl.is_close_degrees(1, 2, delta=-1)


@pytest.mark.parametrize("name", ["standard", "combo"])
def test_virtual_shift_matches_reference(name):
    """analyze(apply_virtual_shift=True) (winston_lutz.py:1587-1601): BBs moved by the projected shift, everything recomputed."""
    rows, dpmm = rows_from_golden(name)
    st = build_set(name, rows, dpmm)
    st._apply_virtual_shift(dict(snap_tolerance=3, gantry_reference=0, collimator_reference=0, couch_reference=0))
    v = f"{name}_vshift"
    np.testing.assert_allclose([[im.bb.x, im.bb.y] for im in st.images], GOLD[f"{v}/bbs"], rtol=0, atol=1e-9)
    np.testing.assert_allclose([im.cax2bb_distance for im in st.images], GOLD[f"{v}/cax2bb_distances"], rtol=0, atol=1e-9)
    rd = st.results_data()
    for k in SCALARS:
        np.testing.assert_allclose(getattr(rd, k), GOLD[f"{v}/{k}"], rtol=0, atol=1e-7, err_msg=k)
    sv = st.bb_shift_vector
    np.testing.assert_allclose([sv.x, sv.y, sv.z], GOLD[f"{v}/bb_shift_vector"], rtol=0, atol=1e-9)


@pytest.mark.parametrize("name", list(SETS))
def test_results_text_follows_the_reference_format(name):
    """WinstonLutz.results() (winston_lutz.py:2501-2546): the lines, rebuilt here from the scalars the UNMODIFIED reference produced for
    the same set (golden file), in the reference's format strings."""
    rows, dpmm = rows_from_golden(name)
    st = build_set(name, rows, dpmm)
    g = lambda k: float(GOLD[f"{name}/{k}"])
    n = int(g("num_total_images"))
    expect = [
        "Winston-Lutz Analysis",
        "=================================",
        f"Number of images: {n}",
        f"Maximum 2D CAX->BB distance: {g('max_2d_cax_to_bb_mm'):.2f}mm",
        f"Median 2D CAX->BB distance: {g('median_2d_cax_to_bb_mm'):.2f}mm",
        f"Mean 2D CAX->BB distance: {g('mean_2d_cax_to_bb_mm'):.2f}mm",
        f"Shift to iso: facing gantry, move BB: {str(GOLD[f'{name}/shift_instructions'])}",
        f"Gantry 3D isocenter diameter: {g('gantry_3d_iso_diameter_mm'):.2f}mm ({int(g('num_gantry_images'))}/{n} images considered)",
        f"Maximum Gantry RMS deviation (mm): {g('max_gantry_rms_deviation_mm'):.2f}mm",
        f"Maximum EPID RMS deviation (mm): {g('max_epid_rms_deviation_mm'):.2f}mm",
        f"Gantry+Collimator 3D isocenter diameter: {g('gantry_coll_3d_iso_diameter_mm'):.2f}mm ({int(g('num_gantry_coll_images'))}/{n} images considered)",
        f"Collimator 2D isocenter diameter: {g('coll_2d_iso_diameter_mm'):.2f}mm ({int(g('num_coll_images'))}/{n} images considered)",
        f"Maximum Collimator RMS deviation (mm): {g('max_coll_rms_deviation_mm'):.2f}",
        f"Couch 2D isocenter diameter: {g('couch_2d_iso_diameter_mm'):.2f}mm ({int(g('num_couch_images'))}/{n} images considered)",
        f"Maximum Couch RMS deviation (mm): {g('max_couch_rms_deviation_mm'):.2f}",
    ]
    assert st.results(as_list=True) == expect
    assert st.results() == "\n".join(expect)
    st._is_analyzed = False
    with pytest.raises(ValueError, match="not analyzed"):
        st.results()
