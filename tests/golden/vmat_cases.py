"""Seeded synthetic VMAT (DRGS / DRMLC / DRCS) and DLG cases shared by the golden generator, the oracle tests and the GPU parity
tests.  The frames are built with oracle/synth.py (the restated image generator).  ``drmlc_contrived`` is the reference's own
synthetic test (tests_basic/test_vmat.py:708-752: AS1200 open 110 x 110 mm FFF field, DMLC image of four 150 x 20 mm strips at
+-15 / +-45 mm, expected R_corr 100, R_dev 0, segment centres (506, 640) / (685, 640) +- 5 px)."""
from __future__ import annotations

import numpy as np

from oracle import synth

VMAT_CASES = ["drmlc_contrived", "drgs_7", "drgs_inverted_swapped", "drmlc_offset_field", "drgs_custom_roi", "drmlc_1024",
              "drgs_no_ground", "drgs_failing", "drmlc_center_fallback", "drmlc_nan"]
DRCS_CASES = ["drcs_basic", "drcs_swapped"]
DLG_CASES = ["dlg_millennium", "dlg_hd"]


def _strips(fr, offsets_mm, width_mm, alphas, length_mm=150, field="fff"):
    add = {"fff": fr.add_fff_field, "filtered": fr.add_filtered_field, "perfect": fr.add_perfect_field}[field]
    for off, a in zip(offsets_mm, alphas):
        add((length_mm, width_mm), cax_offset_mm=(0, off), alpha=a)


def vmat_case(name):
    """-> (klass name, image1 uint16, image2 uint16, pixel_spacing_mm, sid, ctor kwargs, analyze kwargs)"""
    if name == "drmlc_contrived":
        o = synth.as1200(1000.0)
        o.add_fff_field((110, 110))
        o.gaussian(2.0)
        d = synth.as1200(1000.0)
        _strips(d, (45, 15, -15, -45), 20, (1.0,) * 4)
        d.gaussian(2.0)
        d.noise(0.005, seed=101)
        return "DRMLC", o.image, d.image, o.pixel_size, 1000.0, {}, {}
    if name in ("drgs_7", "drgs_inverted_swapped", "drgs_no_ground", "drgs_failing"):
        o = synth.as1200(1000.0)
        o.add_filtered_field((150, 150), alpha=0.6)
        o.gaussian(2.0)
        o.noise(0.002, seed=111)
        d = synth.as1200(1000.0)
        alphas = (0.30, 0.302, 0.299, 0.30, 0.301, 0.298, 0.30) if name != "drgs_failing" else (0.30, 0.33, 0.299, 0.30, 0.27, 0.298, 0.30)
        _strips(d, (-60, -40, -20, 0, 20, 40, 60), 18, alphas, field="filtered")
        d.gaussian(1.5)
        d.noise(0.002, seed=112)
        if name == "drgs_inverted_swapped":
            return "DRGS", d.inverted(), o.inverted(), o.pixel_size, 1000.0, {}, {}
        if name == "drgs_no_ground":
            o.constant(900)
            d.constant(700)
            return "DRGS", o.image, d.image, o.pixel_size, 1000.0, {"ground": False, "check_inversion": False}, {"tolerance": 3}
        return "DRGS", o.image, d.image, o.pixel_size, 1000.0, {}, {}
    if name == "drmlc_offset_field":
        o = synth.as1200(1500.0)
        o.add_filtered_field((120, 130), cax_offset_mm=(0, 8.0), alpha=0.7)
        o.gaussian(2.0)
        o.noise(0.002, seed=121)
        d = synth.as1200(1500.0)
        for off, a in zip((-45, -15, 15, 45), (0.35, 0.352, 0.349, 0.351)):
            d.add_filtered_field((120, 26), cax_offset_mm=(0, 8.0 + off), alpha=a)
        d.gaussian(2.0)
        d.noise(0.002, seed=122)
        return "DRMLC", o.image, d.image, o.pixel_size, 1500.0, {}, {"tolerance": 2.0}
    if name == "drgs_custom_roi":
        o = synth.as1000(1500.0)
        o.add_filtered_field((140, 140), alpha=0.6)
        o.gaussian(2.0)
        o.noise(0.002, seed=131)
        d = synth.as1000(1500.0)
        _strips(d, (-50, -25, 0, 25, 50), 22, (0.3, 0.31, 0.3, 0.29, 0.3), field="filtered")
        d.gaussian(2.0)
        d.noise(0.002, seed=132)
        roi = {"A": {"offset_mm": -50}, "B": {"offset_mm": -25}, "C": {"offset_mm": 0}, "D": {"offset_mm": 25}, "E": {"offset_mm": 50}}
        return "DRGS", o.image, d.image, o.pixel_size, 1500.0, {}, {"roi_config": roi, "segment_size_mm": (8, 80), "tolerance": 5}
    if name == "drmlc_1024":
        o = synth.epid1024()
        o.add_filtered_field((130, 130), alpha=0.6)
        o.gaussian(2.0)
        o.noise(0.002, seed=141)
        d = synth.epid1024()
        _strips(d, (-45, -15, 15, 45), 28, (0.3, 0.3005, 0.2995, 0.3), field="filtered")
        d.gaussian(2.0)
        d.noise(0.002, seed=142)
        return "DRMLC", d.image, o.image, o.pixel_size, 1000.0, {}, {}
    if name in ("drmlc_center_fallback", "drmlc_nan"):
        # field centre far off the central third: the reference warns and uses the image centre; in "drmlc_nan" the ROIs then lie
        # on the zero background of both images (0 / 0 -> nan, x / 0 -> inf like numpy)
        o = synth.as1200(1000.0)
        o.add_filtered_field((150, 60), cax_offset_mm=(0, -100.0), alpha=0.6)
        o.gaussian(2.0)
        o.noise(0.002, seed=151)
        d = synth.as1200(1000.0)
        d.add_filtered_field((150, 60), cax_offset_mm=(0, -100.0), alpha=0.3)
        d.gaussian(2.0)
        d.noise(0.003, seed=152)
        d.image[:, ::7] = (d.image[:, ::7].astype(np.int64) * 9 // 10).astype(np.uint16)   # rougher: identified as the DMLC image
        if name == "drmlc_center_fallback":
            o.constant(3000)
            d.constant(1500)
            o.noise(0.002, seed=153)
            d.noise(0.002, seed=154)
        return "DRMLC", o.image, d.image, o.pixel_size, 1000.0, {}, {"tolerance": 4}
    raise KeyError(name)


def _pie(fr, alpha_by_sector, r_out_mm=90.0, spoke_boost=0.04, spoke_deg=1.2, missing=(60.0, 120.0), rot_deg=0.0):
    """A disc of radius r_out with an empty pie slice (image angles ``missing``, degrees, measured like CircleProfile does:
    atan2(-dy, dx) ... the exact convention does not matter, both implementations see the same pixels), sectors of slightly
    different intensity and brighter spokes at the sector boundaries (what a DRCS DMLC image looks like)."""
    h, w = fr.shape
    yy, xx = np.mgrid[0:h, 0:w]
    cy, cx = (h - 1) / 2.0, (w - 1) / 2.0
    r = np.hypot(yy - cy, xx - cx) * fr.pixel_size / fr.mag
    ang = (np.degrees(np.arctan2(yy - cy, xx - cx)) - rot_deg) % 360.0
    img = np.zeros(fr.shape)
    inside = r < r_out_mm
    n = len(alpha_by_sector)
    lo, hi = missing
    span = (360.0 - (hi - lo)) / n
    for k, a in enumerate(alpha_by_sector):
        a0 = (hi + k * span) % 360.0
        sel = inside & (((ang - a0) % 360.0) < span)
        img[sel] = a * synth.U16_MAX
    if spoke_boost:
        for k in range(n + 1):
            a0 = (hi + k * span) % 360.0
            d = np.abs(((ang - a0 + 180.0) % 360.0) - 180.0)
            img[inside & (d < spoke_deg)] += spoke_boost * synth.U16_MAX
    fr.image = np.clip(fr.image.astype(float) + img, 0, synth.U16_MAX).astype(np.uint16)


def drcs_case(name):
    o = synth.as1200(1000.0)
    _pie(o, (0.5,), spoke_boost=0.0, missing=(90.0, 90.0))
    o.gaussian(1.5)
    o.noise(0.001, seed=201)
    d = synth.as1200(1000.0)
    _pie(d, (0.25, 0.2505, 0.2495, 0.25, 0.2502), rot_deg=0.2)
    d.gaussian(1.0)
    d.noise(0.001, seed=202)
    if name == "drcs_basic":
        return "DRCS", o.image, d.image, o.pixel_size, 1000.0, {}, {}
    if name == "drcs_swapped":
        return "DRCS", d.image, o.image, o.pixel_size, 1000.0, {}, {"tolerance": 2.5, "collimator_config": {"A": 150, "C": 30, "E": 270}}
    raise KeyError(name)


def dlg_case(name):
    """-> (image uint16, pixel_spacing_mm, sid, gaps, mlc name, y_field_size, profile_width)"""
    mlc = {"dlg_millennium": "MILLENNIUM", "dlg_hd": "HD_MILLENNIUM"}[name]
    fr = synth.as1200(1000.0)
    h, w = fr.shape
    gaps = (-0.9, -1.1, -1.3, -1.5, -1.7, -1.9)
    y_field = 100.0
    base = np.full(fr.shape, 0.35 * synth.U16_MAX)
    band = y_field / len(gaps)
    yy = (np.arange(h) - h / 2) / fr.dpmm          # mm from the centre line
    xx = (np.arange(w) - w / 2) / fr.dpmm
    g_sorted = sorted(gaps)
    for k, g in enumerate(g_sorted):
        rows = (yy < y_field / 2 - k * band) & (yy >= y_field / 2 - (k + 1) * band)
        true_dlg = 1.45 if name == "dlg_millennium" else 1.2
        height = (g + true_dlg) * 0.08 * synth.U16_MAX          # dip for overlap, bump for a gap
        prof = height * np.exp(-0.5 * (xx / 0.8) ** 2)
        base[rows, :] += prof[None, :]
    fr.image = np.clip(base, 0, synth.U16_MAX).astype(np.uint16)
    fr.gaussian(0.6)
    fr.noise(0.0008, seed=301 if name == "dlg_millennium" else 302)
    return fr.image, fr.pixel_size, 1000.0, gaps, mlc, y_field, 10
