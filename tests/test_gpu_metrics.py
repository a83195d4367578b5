"""GPU parity of the image-metric plug-ins (rows a12, a28, a36): ``BaseImage.compute`` + ``SizedDiskLocator`` / ``SizedDiskRegion``
(metrics/image.py:402-667), ``WeightedCentroid`` (:959-983) and ``RectangleROI`` statistics (core/roi.py:481-706).

* the disk locator is checked (a) against the reference's own synthetic expectations (tests_basic/core/test_image_metrics.py:277-514:
  a 5 mm BB on a 50 x 50 mm field of an AS1000 panel must be found within 1 px of (511.5, 383.5), wrong areas / sizes must raise
  ValueError) and (b) point by point against the CPU oracle of find_features (oracle/wl_oracle.find_bbs) on the same window;
* WeightedCentroid and RectangleROI are checked against numpy on the same pixels (integer frames: exact)."""
import math

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def create_bb_image(field_size=(50, 50), bb_size=5, offset=(0, 0), seed=0):
    """tests_basic/core/test_image_metrics.py:30-46 with the restated generator (skimage is absent): AS1000 @ SID 1000, filtered
    field, gaussian 2 mm, PerfectBBLayer (alpha -0.5), RandomNoiseLayer (seeded here)."""
    from oracle import synth

    fr = synth.as1000(1000.0)
    fr.add_filtered_field(field_size, (0, 0))
    fr.gaussian(2.0)
    fr.add_bb(bb_size, offset, alpha=-0.5)
    fr.noise(0.001, seed=seed)
    return fr


def _img(fr):
    from pylinac_b200.core import image

    return image.ArrayImage(fr.image, dpi=25.4 / fr.pixel_size, sid=fr.sid)


def _oracle_points(img, left, right, top, bottom, radius_mm, tol_mm, invert=True, max_number=1, min_sep_mm=5):
    from oracle import wl_oracle

    sample = img.array[top:bottom, left:right]
    if invert:
        sample = -sample + sample.max() + sample.min()        # uint16 modular, like array_utils.invert
    pts, _ = wl_oracle.find_bbs(sample, top, left, img.dpmm, radius_mm, tol_mm, max_number=max_number, min_separation_mm=min_sep_mm)
    return pts


def test_disk_locator_pixels_perfect_image():
    from pylinac_b200.metrics.image import SizedDiskLocator

    img = _img(create_bb_image(bb_size=5))
    pos = img.compute(metrics=[SizedDiskLocator(expected_position=(511.5, 383.5), search_window=(50, 50), radius=6, radius_tolerance=1,
                                                max_number=1)])
    assert len(pos) == 1
    assert abs(pos[0].x - 511.5) < 1 and abs(pos[0].y - 383.5) < 1
    # the oracle on the same window (radius / tolerance go through the reference's px -> mm conversion)
    left, right = math.floor(511.5 - 25), math.ceil(511.5 + 25)
    top, bottom = math.floor(383.5 - 25), math.ceil(383.5 + 25)
    o = _oracle_points(img, left, right, top, bottom, 6 / img.dpmm, 1 / img.dpmm, min_sep_mm=5 / img.dpmm)
    assert len(o) == 1
    assert abs(pos[0].x - o[0][0]) < 1e-9 and abs(pos[0].y - o[0][1]) < 1e-9
    assert "Disk Region" in img.metric_values and img.metrics[0].x_offset == left and img.metrics[0].y_offset == top


@pytest.mark.parametrize("kw", [
    dict(img=dict(offset=(20, 20)), m=dict(expected_position=(511.5, 383.5), search_window=(10, 10), radius=20, radius_tolerance=1)),
    dict(img=dict(bb_size=1), m=dict(expected_position=(511.5, 383.5), search_window=(10, 10), radius=20, radius_tolerance=2)),
])
def test_disk_locator_pixels_raises(kw):
    from pylinac_b200.metrics.image import SizedDiskLocator

    img = _img(create_bb_image(**kw["img"]))
    with pytest.raises(ValueError):
        img.compute(metrics=[SizedDiskLocator(**kw["m"])])


def test_disk_locator_physical_variants():
    """The four constructors (metrics/image.py:416-562) on the reference's own cases (tests_basic/core/test_image_metrics.py:329-500)."""
    from pylinac_b200.metrics.image import SizedDiskLocator, SizedDiskRegion

    img = _img(create_bb_image(bb_size=5))
    ph = img.compute(SizedDiskLocator.from_physical(expected_position_mm=(200, 150), search_window_mm=(10, 10), radius_mm=2.5,
                                                    radius_tolerance_mm=1, name="phys"))
    assert abs(ph[0].x - 511.5) < 1 and abs(ph[0].y - 383.5) < 1
    with pytest.raises(ValueError):   # expected radius 5 mm for a 5 mm BB (test_bb_too_small of the physical variant)
        _img(create_bb_image(bb_size=5)).compute(SizedDiskLocator.from_physical((200, 150), (10, 10), radius_mm=5, radius_tolerance_mm=1))
    pos = img.compute(SizedDiskLocator.from_center_physical(expected_position_mm=(0, 0), search_window_mm=(10, 10), radius_mm=2.5,
                                                            radius_tolerance_mm=1))
    assert abs(pos[0].x - 511.5) < 1 and abs(pos[0].y - 383.5) < 1
    pos2 = img.compute(SizedDiskLocator.from_center(expected_position=(0, 0), search_window=(50, 50), radius=6, radius_tolerance=1))
    assert abs(pos2[0].x - pos[0].x) < 0.2 and abs(pos2[0].y - pos[0].y) < 0.2
    assert set(img.metric_values) == {"phys", "Disk Region", "Disk Region-1"}          # uniquify (core/utilities.py:368-377)
    with pytest.raises(ValueError):   # radius 2 mm too big with 1 mm tolerance (test_barely_too_small)
        _img(create_bb_image(bb_size=10)).compute(SizedDiskLocator.from_center_physical((0, 0), (20, 20), radius_mm=7, radius_tolerance_mm=1))
    with pytest.raises(ValueError):   # radius 2 mm too small (test_barely_too_big)
        _img(create_bb_image(bb_size=10)).compute(SizedDiskLocator.from_center_physical((0, 0), (20, 20), radius_mm=3, radius_tolerance_mm=1))
    # shifted BB and shifted expectation
    sh = _img(create_bb_image(bb_size=5, offset=(20, 20), field_size=(75, 75)))
    p = sh.compute(SizedDiskLocator.from_center_physical(expected_position_mm=(20, 20), search_window_mm=(10, 10), radius_mm=2.5,
                                                         radius_tolerance_mm=1))
    assert len(p) == 1 and abs(p[0].x - (511.5 + 20 * sh.dpmm)) < 1.5 and abs(p[0].y - (383.5 + 20 * sh.dpmm)) < 1.5
    # regions: SizedDiskRegion returns region properties of the accepted regions (sample coordinates)
    regs = img.compute(SizedDiskRegion.from_center_physical((0, 0), (10, 10), radius_mm=2.5, radius_tolerance_mm=1))
    assert len(regs) == 1
    r = regs[0]
    assert r.area_filled >= r.area > 50 and 0.9 < r.solidity <= 1.0 and r.perimeter > 20
    assert r.bbox[2] - r.bbox[0] == pytest.approx(r.bbox[3] - r.bbox[1], abs=2)
    assert abs(r.centroid_weighted[0] - r.centroid[0]) < 1 and abs(r.centroid_weighted[1] - r.centroid[1]) < 1


def test_disk_locator_shifted_like_the_oracle():
    from pylinac_b200.metrics.image import SizedDiskLocator

    for seed, off in [(1, (3.0, -2.0)), (2, (-4.5, 1.25)), (3, (0.4, 0.7))]:
        fr = create_bb_image(bb_size=5, offset=off, seed=seed)
        img = _img(fr)
        m = SizedDiskLocator.from_center_physical((0, 0), (45, 45), radius_mm=2.5, radius_tolerance_mm=2.0)
        pos = img.compute(m)
        h, w = img.shape
        win = 45 * img.dpmm
        left, right = max(math.floor(w / 2 - win / 2), 0), math.ceil(w / 2 + win / 2)
        top, bottom = max(math.floor(h / 2 - win / 2), 0), math.ceil(h / 2 + win / 2)
        o = _oracle_points(img, left, right, top, bottom, 2.5, 2.0)
        assert len(pos) == len(o) == 1
        assert abs(pos[0].x - o[0][0]) < 1e-9 and abs(pos[0].y - o[0][1]) < 1e-9, (seed, pos, o)


def test_metric_must_not_modify_the_image():
    from pylinac_b200.metrics.image import MetricBase

    class Bad(MetricBase):
        name = "bad"

        def calculate(self):
            self.image.array[0, 0] += 1

    img = _img(create_bb_image())
    with pytest.raises(RuntimeError):
        img.compute(Bad())


def test_weighted_centroid_matches_numpy():
    from pylinac_b200.core import image
    from pylinac_b200.metrics.image import WeightedCentroid

    rng = np.random.default_rng(4)
    a = rng.integers(0, 65535, (300, 421)).astype(np.uint16)
    a[100:180, 250:330] = 60000
    p = image.ArrayImage(a).compute(WeightedCentroid())
    yi, xi = np.indices(a.shape)
    assert p.x == np.sum(xi * a) / np.sum(a) and p.y == np.sum(yi * a) / np.sum(a)        # exact integer sums, one fp64 division
    f = a.astype(np.float64) / 7.0
    pf = image.ArrayImage(f).compute(WeightedCentroid())
    assert pf.x == pytest.approx(np.sum(xi * f) / np.sum(f), rel=1e-12) and pf.y == pytest.approx(np.sum(yi * f) / np.sum(f), rel=1e-12)
    with pytest.raises(ValueError):
        image.ArrayImage(np.zeros((8, 8), np.uint16)).compute(WeightedCentroid())


def test_rectangle_roi_statistics():
    from pylinac_b200.core.geometry import Point
    from pylinac_b200.core.roi import RectangleROI

    rng = np.random.default_rng(8)
    a = rng.integers(100, 60000, (200, 260)).astype(np.uint16)
    # the FieldAnalysis central ROI construction (field_analysis.py:755-766): integer corners -> rows [top, top + h), cols [left, left + w)
    left, top, w, h = 40, 30, 37, 22
    roi = RectangleROI(a, width=w, height=h, center=Point(w / 2 + left, h / 2 + top))
    px = a[top:top + h, left:left + w]
    assert np.array_equal(roi.pixel_array, px)
    assert roi.mean == float(np.mean(px)) and roi.min == float(px.min()) and roi.max == float(px.max())
    assert roi.std == pytest.approx(float(np.std(px)), rel=1e-13)
    assert roi.pixel_value == roi.mean
    # float image, half-pixel centre: pixel centres inside or on the boundary of the corner polygon
    f = a.astype(np.float64) / 3.0
    roi2 = RectangleROI(f, width=10, height=6, center=Point(100.5, 50.0))
    bl_y, tl_y, tl_x, br_x = 50.0 + 3, 50.0 - 3, 100.5 - 5, 100.5 + 5
    rows = [r for r in range(200) if tl_y <= r <= bl_y - 1]
    cols = [c for c in range(260) if tl_x <= c <= br_x - 1]
    sel = f[np.ix_(rows, cols)]
    assert roi2.mean == pytest.approx(float(sel.mean()), rel=1e-13) and roi2.std == pytest.approx(float(sel.std()), rel=1e-10)
    assert roi2.min == float(sel.min()) and roi2.max == float(sel.max())
    # rotated ROI: selection by the same point-in-polygon rule evaluated in numpy
    roi3 = RectangleROI(a, width=40, height=20, center=Point(130, 100), rotation=30)
    poly = roi3._polygon_xy()
    yy, xx = np.mgrid[0:200, 0:260]
    inside = np.ones(a.shape, bool)
    area2 = sum(poly[k][0] * poly[(k + 1) % 4][1] - poly[(k + 1) % 4][0] * poly[k][1] for k in range(4))
    sgn = 1.0 if area2 > 0 else -1.0
    for k in range(4):
        ax, ay = poly[k]
        bx, by = poly[(k + 1) % 4]
        inside &= sgn * ((bx - ax) * (yy - ay) - (by - ay) * (xx - ax)) >= 0
    sel = a[inside]
    assert roi3.mean == pytest.approx(float(sel.mean()), rel=1e-13) and roi3.max == float(sel.max())
    with pytest.raises(ValueError):
        RectangleROI(a, width=1, height=5, center=Point(5, 5))
    with pytest.raises(ValueError):
        _ = roi3.pixel_array
