"""Run the UNMODIFIED reference PicketFence (stub-imported from /root/reference) on an
ndarray and flatten what it computed into the same dict layout as oracle.pf_oracle.pf_analyze.
Only usable where /root/reference exists (this container); used by make_pf_golden.py and
tests/test_oracle_vs_reference.py."""
from __future__ import annotations

import io
import types

import numpy as np

from oracle.refstub import FakeDicomDataset, import_reference


def reference_pf(frame, pixel_spacing_mm, sid, ctor_kwargs=None, analyze_kwargs=None):
    import_reference()
    from pylinac import picketfence as rpf
    from pylinac.core import image as rimage

    ds = FakeDicomDataset(frame, pixel_spacing_mm, sid=sid)
    old = rimage.retrieve_dicom_file, rimage.pixels
    rimage.retrieve_dicom_file = lambda path: ds
    rimage.pixels = types.SimpleNamespace(apply_rescale=lambda arr, md: arr)
    try:
        pf = rpf.PicketFence(io.BytesIO(b"fake"), **(ctor_kwargs or {}))
    finally:
        rimage.retrieve_dicom_file, rimage.pixels = old
    pf.analyze(**(analyze_kwargs or {}))
    rd = pf.results_data()
    out = {}
    out["shape"] = tuple(pf.image.shape)
    out["dpmm"] = float(pf.image.dpmm)
    out["orientation"] = 0 if pf.orientation == rpf.Orientation.UP_DOWN else 1
    out["n_meas"] = len(pf.mlc_meas)
    out["meas_leaf"] = np.array([m.leaf_num for m in pf.mlc_meas], dtype=np.int64)
    out["meas_picket"] = np.array([m.picket_num for m in pf.mlc_meas], dtype=np.int64)
    out["meas_position"] = np.array([list(m.position) for m in pf.mlc_meas], dtype=np.float64)
    out["meas_error"] = np.array([list(m.error) for m in pf.mlc_meas], dtype=np.float64)
    out["meas_width_mm"] = np.array([m.profile.field_width_mm for m in pf.mlc_meas], dtype=np.float64)
    out["picket_idx"] = np.array(sorted({int(m._approximate_idx) for m in pf.mlc_meas}), dtype=np.int64)
    out["picket_spacing"] = float(pf.mlc_meas[0]._spacing)
    out["fits"] = np.array([np.asarray(p.fit.coefficients, dtype=float) for p in pf.pickets])
    out["percent_passing"] = rd.percent_leaves_passing
    out["max_error"] = rd.max_error_mm
    out["abs_median_error"] = rd.absolute_median_error_mm
    out["max_error_picket"] = int(rd.max_error_picket)
    out["max_error_leaf"] = rd.max_error_leaf
    out["passed"] = bool(rd.passed)
    out["failed_leaves"] = list(rd.failed_leaves)
    out["offsets_from_cax_mm"] = np.array(rd.offsets_from_cax_mm)
    out["mean_picket_spacing"] = rd.mean_picket_spacing_mm
    out["mlc_skew"] = rd.mlc_skew
    out["number_of_pickets"] = rd.number_of_pickets
    out["picket_widths"] = np.array([[rd.picket_widths[f"picket_{k}"][s] for s in ("max", "mean", "median", "min")]
                                     for k in range(rd.number_of_pickets)])
    out["mlc_positions_by_leaf"] = rd.mlc_positions_by_leaf
    out["mlc_errors_by_leaf"] = rd.mlc_errors_by_leaf
    return out


def reference_starshot(frame, pixel_spacing_mm, sid, analyze_kwargs=None):
    """Run the UNMODIFIED reference Starshot on an ndarray -> the dict layout of oracle.starshot_oracle.starshot_analyze."""
    import_reference()
    from pylinac import starshot as rs

    s = rs.Starshot(np.array(frame), dpi=25.4 / pixel_spacing_mm, sid=sid)
    # instrument the loop count of _get_reasonable_wobble without touching the reference: count StarProfile constructions
    count = {"n": 0}
    orig = rs.StarProfile

    class Counting(orig):
        def __init__(self, *a, **k):
            count["n"] += 1
            super().__init__(*a, **k)

    rs.StarProfile = Counting
    try:
        auto, local_max = None, None
        s.analyze(**(analyze_kwargs or {}))
    finally:
        rs.StarProfile = orig
    cp = s.circle_profile
    out = {
        "iterations": count["n"],
        "radius_px": float(cp.radius),
        "profile_len": len(cp.values),
        "peak_idx": np.array([int(p.idx) for p in cp.peaks], dtype=np.int64),
        "peak_xy": np.array([[float(p.x), float(p.y)] for p in cp.peaks]),
        "wobble_center": np.array([s.wobble.center.x, s.wobble.center.y], dtype=float),
        "wobble_radius_px": float(s.wobble.radius),
        "wobble_radius_mm": float(s.wobble.radius_mm),
        "angles": np.array(s.angles, dtype=float),
        "n_lines": len(s.lines),
        "passed": bool(s.passed),
    }
    return out


def reference_field(frame, pixel_spacing_mm, sid, analyze_kwargs=None):
    """Run the UNMODIFIED reference FieldAnalysis on an ndarray -> flat dict of its _results / _extra_results, plus the central ROI
    statistics (skimage.draw.polygon is served by oracle/skimage_shim.py: restated, unpinned at that boundary)."""
    from oracle import skimage_shim

    skimage_shim.install()
    from pylinac import field_analysis as fa

    kw = dict(analyze_kwargs or {})
    if "protocol" in kw:
        kw["protocol"] = fa.Protocol[kw["protocol"]]
    f = fa.FieldAnalysis(np.array(frame), image_kwargs=dict(dpi=25.4 / pixel_spacing_mm, sid=sid))
    f.analyze(**kw)
    out = {}
    for k, v in f._results.items():
        out[k] = np.asarray(v, dtype=float)
    for k, v in f._extra_results.items():
        out[k] = float(v)
    for k in ("mean", "max", "min", "std"):
        out[f"central_roi_{k}"] = float(getattr(f.central_roi, k))
    out["strip_rows"] = np.array([f._upper_h_index, f._lower_h_index])
    out["strip_cols"] = np.array([f._left_v_index, f._right_v_index])
    out["profile_len"] = np.array([len(f.horiz_profile.values), len(f.vert_profile.values)])
    return out


def reference_wl2d(frame, pixel_spacing_mm, sid, gantry, coll, couch, analyze_kwargs=None):
    """Run the UNMODIFIED reference WinstonLutz2D (skimage calls served by oracle/skimage_shim.py) on an ndarray."""
    from oracle import skimage_shim
    from oracle.refstub import reference_image_from_array

    skimage_shim.install()
    from pylinac import winston_lutz as wl

    img = reference_image_from_array(wl.WinstonLutz2D, np.array(frame), pixel_spacing_mm, sid=sid, gantry=gantry, coll=coll, couch=couch)
    img.analyze(**(analyze_kwargs or {}))
    rd = img.results_data()
    return {
        "shape": np.array(img.shape), "field_cax": np.array([img.field_cax.x, img.field_cax.y], dtype=float),
        "bb": np.array([img.bb.x, img.bb.y], dtype=float), "epid": np.array([img.epid.x, img.epid.y], dtype=float),
        "cax2bb_vector": np.array([rd.cax2bb_vector.x, rd.cax2bb_vector.y], dtype=float), "cax2bb_distance": float(rd.cax2bb_distance),
        "cax2epid_vector": np.array([rd.cax2epid_vector.x, rd.cax2epid_vector.y], dtype=float),
        "cax2epid_distance": float(rd.cax2epid_distance), "variable_axis": np.array(str(rd.variable_axis)),
    }


def reference_vmat(klass, image1, image2, pixel_spacing_mm, sid, ctor_kwargs=None, analyze_kwargs=None):
    """Run the UNMODIFIED reference DRGS / DRMLC / DRCS (pylinac/vmat.py) on two ndarrays (``image.load`` of an array gives an
    ArrayImage; skimage.draw.polygon / EuclideanTransform behind ``RectangleROI.pixels_flat`` and the DRCS geometry are served by
    oracle/skimage_shim.py: restated, unpinned at that boundary)."""
    import warnings

    from oracle import skimage_shim

    skimage_shim.install()
    from pylinac import vmat as rv

    cls = getattr(rv, klass)
    with warnings.catch_warnings(record=True) as wlist:
        warnings.simplefilter("always")
        # observe (not alter) which loaded image the reference takes as the open field
        seen = {}
        orig_identify = cls._identify_images

        def spy(self, i1, i2):
            orig_identify(self, i1, i2)
            seen["open_is_first"] = int(self.open_image is i1)

        cls._identify_images = spy
        try:
            v = cls(image_paths=(np.array(image1), np.array(image2)), dpi=25.4 / pixel_spacing_mm, sid=sid, **(ctor_kwargs or {}))
        finally:
            cls._identify_images = orig_identify
        v.analyze(**(analyze_kwargs or {}))
        out = {
            "open_is_first": seen["open_is_first"],
            "r_corr": np.array([s.r_corr for s in v.segments], dtype=float),
            "r_dev": np.array([s.r_dev for s in v.segments], dtype=float),
            "stdev": np.array([s.stdev for s in v.segments], dtype=float),
            "passed_seg": np.array([bool(s.passed) for s in v.segments]),
            "center_x": np.array([s.center.x for s in v.segments], dtype=float),
            "center_y": np.array([s.center.y for s in v.segments], dtype=float),
            "max_r_deviation": float(v.max_r_deviation), "avg_abs_r_deviation": float(v.avg_abs_r_deviation),
            "avg_r_deviation": float(v.avg_r_deviation), "passed": bool(v.passed),
            "open_sum": float(np.asarray(v.open_image.array, dtype=np.float64).sum()),
            "dmlc_sum": float(np.asarray(v.dmlc_image.array, dtype=np.float64).sum()),
        }
        if klass == "DRCS":
            out["coll_angle_deviation"] = np.array([cd.angle_deviation for cd in v.collimator_deviations], dtype=float)
            out["coll_angle_measured"] = np.array([cd.angle_measured for cd in v.collimator_deviations], dtype=float)
            out["rotation_offset_deg"] = float(v.rotation_offset_deg)
            out["rotation"] = np.array([s.rotation for s in v.segments], dtype=float)
        rd = v.results_data()
        out["rd_max_deviation_percent"] = float(rd.max_deviation_percent)
        out["rd_abs_mean_deviation"] = float(rd.abs_mean_deviation)
    out["n_user_warnings"] = sum(1 for w in wlist if issubclass(w.category, UserWarning) and "VMAT field center" in str(w.message))
    return out


def reference_dlg(frame, pixel_spacing_mm, sid, gaps, mlc_name, y_field_size=100, profile_width=10):
    """Run the UNMODIFIED reference DLG (pylinac/dlg.py) on an ndarray (LinacDicomImage served by a fake dataset)."""
    import_reference()
    from pylinac import dlg as rdlg
    from pylinac.core import image as rimage
    from pylinac.picketfence import MLC

    ds = FakeDicomDataset(frame, pixel_spacing_mm, sid=sid)
    old = rimage.retrieve_dicom_file, rimage.pixels
    rimage.retrieve_dicom_file = lambda path: ds
    rimage.pixels = types.SimpleNamespace(apply_rescale=lambda arr, md: arr)
    try:
        d = rdlg.DLG(io.BytesIO(b"fake"))
    finally:
        rimage.retrieve_dicom_file, rimage.pixels = old
    d.analyze(gaps=gaps, mlc=MLC[mlc_name], y_field_size=y_field_size, profile_width=profile_width)
    return {"measured_dlg": float(d.measured_dlg), "measured_dlg_per_leaf": np.array(d.measured_dlg_per_leaf, dtype=float),
            "planned_dlg_per_leaf": np.array(d.planned_dlg_per_leaf, dtype=float), "slope": float(d._lin_fit.slope),
            "intercept": float(d._lin_fit.intercept), "dpmm": float(d.image.dpmm), "dtype": np.array(str(d.image.array.dtype))}


def reference_locator(frame, pixel_spacing_mm, sid, spec):
    """Run the UNMODIFIED reference's whole-frame locator metric (metrics/image.py) on an ArrayImage; skimage label / clear_border /
    regionprops are served by oracle/skimage_shim.py (restated, unpinned at that boundary).  -> points [k, 2] (x, y) or raises."""
    from oracle import skimage_shim

    skimage_shim.install()
    from pylinac.core import image as rimage
    from pylinac.metrics import image as rmi

    img = rimage.ArrayImage(np.array(frame), dpi=25.4 / pixel_spacing_mm, sid=sid)
    cls = getattr(rmi, spec["cls"])
    metric = cls.from_physical(**spec["kw"]) if spec.get("physical") else cls(**spec["kw"])
    pts = img.compute(metrics=metric)
    return np.array([[p.x, p.y] for p in pts], dtype=float).reshape(-1, 2)
