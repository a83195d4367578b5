"""Host-side API glue that needs no GPU: warning capture (core/warnings.py), geometry helpers, dtype helpers."""
import warnings

import numpy as np
import pytest

from pylinac_b200.core.utilities import ResultBase, ResultsDataMixin
from pylinac_b200.core.warnings import WarningCollectorMixin, capture_warnings


class _Res(ResultBase):
    x: int


@capture_warnings
class _Tool(ResultsDataMixin[_Res]):
    def inner(self):
        warnings.warn("inner happened")
        return 1

    def analyze(self):
        warnings.warn("outer happened", RuntimeWarning)
        return self.inner() + 1          # nested decorated call: captured once, by the outermost

    def _generate_results_data(self):
        return _Res(x=3)


def test_capture_warnings_contract():
    """pylinac/core/warnings.py:42-112 -- captured, deduplicated, reported through results_data(), still shown."""
    t = _Tool()
    with warnings.catch_warnings(record=True) as shown:
        warnings.simplefilter("always")
        assert t.analyze() == 2
        t.analyze()
    msgs = [(w["message"], w["category"]) for w in t.get_captured_warnings()]
    assert msgs == [("outer happened", "RuntimeWarning"), ("inner happened", "UserWarning")]   # de-duplicated across the two calls
    assert set(t.get_captured_warnings()[0]) == {"message", "category", "filename", "lineno", "line"}
    assert len(shown) == 4                                                                     # re-emitted every time
    rd = t.results_data()
    assert [w["message"] for w in rd.warnings] == ["outer happened", "inner happened"]
    assert t.results_data(as_dict=True)["warnings"][0]["category"] == "RuntimeWarning"
    t.clear_captured_warnings()
    assert t.results_data().warnings == []
    assert isinstance(t, WarningCollectorMixin)


def test_analysis_classes_collect_warnings():
    from pylinac_b200 import field_profile_analysis, picketfence, starshot, winston_lutz

    for cls in (picketfence.PicketFence, starshot.Starshot, winston_lutz.WinstonLutz2D, winston_lutz.WinstonLutz,
                field_profile_analysis.FieldProfileAnalysis):
        assert issubclass(cls, WarningCollectorMixin)
        assert hasattr(cls.analyze, "__wrapped__"), cls


def test_geometry_circle_rectangle():
    """core/geometry.py:213-405, 632-723 (data parts)."""
    import math

    from pylinac_b200.core.geometry import Circle, Point, Rectangle

    c = Circle((3, 4), radius=2)
    assert (c.center.x, c.center.y, c.diameter) == (3, 4, 4) and c.area == pytest.approx(math.pi * 4)
    assert c.as_dict() == {"center_x": 3, "center_y": 4, "diameter": 4}
    with pytest.raises(TypeError):
        Circle(5)
    r = Rectangle(width=4, height=2, center=(10, 20))
    assert r.area == 8
    assert [(p.x, p.y) for p in r.vertices] == [(8, 19), (12, 19), (12, 21), (8, 21)]
    assert (r.tl_corner.x, r.br_corner.y) == (8, 21)
    r90 = Rectangle(width=4, height=2, center=(10, 20), rotation=90)      # TL -> top right on screen (clockwise)
    v = r90.vertices
    assert (v[0].x, v[0].y) == pytest.approx((11, 18)) and (v[2].x, v[2].y) == pytest.approx((9, 22))
    with pytest.raises(ValueError):
        Rectangle(width=0, height=1, center=Point(0, 0))


def test_convert_to_dtype_formula():
    """core/array_utils.py:172-198 (integer input: no device work involved)."""
    from pylinac_b200.core import array_utils as au

    a = np.array([0, 100, 255], dtype=np.uint8)
    info = np.iinfo(np.uint16)
    exp = np.array(a.astype(float) / 255 * (info.max - info.min) - info.max - 1, dtype=np.uint16)
    np.testing.assert_array_equal(au.convert_to_dtype(a, np.uint16), exp)
    with pytest.raises(ValueError):
        au.convert_to_dtype(np.array([], dtype=np.uint8), np.uint16)


# ---------------------------------------------------------------------------------------------- VMAT / DLG / locators (host logic)
def test_vmat_host_side_validation_and_geometry():
    from pylinac_b200 import vmat

    with pytest.raises(ValueError, match="Exactly 2 images"):
        vmat.DRGS(image_paths=(np.zeros((64, 64), np.uint16),))
    with pytest.raises(ValueError, match="at most"):
        vmat._make_params(2.5, 1.5, (5, 100), list(range(20)), True, True, False)
    p = vmat._make_params(2.5, 1.5, (5, 100), [-45, -15, 15, 45], True, False, True)
    assert (p.nseg, p.ground, p.check_inversion, p.invert_image_order) == (4, 1, 0, 1) and p.offset_mm[3] == 45
    assert vmat.wrap180(190) == -170 and vmat.wrap180(-180) == -180 and vmat.wrap180(180) == -180
    # IEC angle of a line from image points (vmat.py:208-215)
    from pylinac_b200.core.geometry import Point

    assert vmat.CollimatorDeviation.calculate_angle_measured(Point(0, 0), Point(0, -10)) == pytest.approx(0.0)
    assert vmat.CollimatorDeviation.calculate_angle_measured(Point(0, 0), Point(10, 0)) == pytest.approx(270.0)
    assert list(vmat.DRGS._default_roi_config()) == [f"ROI {k}" for k in range(1, 8)]


def test_dlg_windows_follow_the_reference_geometry():
    from pylinac_b200 import dlg
    from pylinac_b200.picketfence import MLC

    bottoms, tops, c0, c1, planned = dlg._windows((1280, 1280), 1 / 0.336, (-0.9, -1.1, -1.3, -1.5, -1.7, -1.9), MLC.MILLENNIUM, 100, 10)
    assert len(bottoms) == len(tops) == len(planned) == 20 and (c0, c1) == (610, 670)
    assert all(t > b for b, t in zip(bottoms, tops))
    assert sorted(set(planned)) == [-1.9, -1.7, -1.5, -1.3, -1.1, -0.9]
    assert dlg._get_dlg_offset(100, 0.0, [-1.9, -1.7]) is None      # a centre exactly on a band edge, like the reference
    with pytest.raises(TypeError):
        dlg._windows((1280, 1280), 1 / 0.336, (-1.0, -1.2, -1.4, -1.6), MLC.MLCI, 100, 10)      # MLCi leaf centres (+-25 mm) sit on band edges


def test_locator_point_bookkeeping_matches_the_reference_rules():
    from pylinac_b200.core.geometry import Point
    from pylinac_b200.metrics import image as mi

    total = [Point(10, 10)]
    mi._dedupe(total, [Point(12, 10), Point(30, 30), Point(31, 30)], 5)      # the second new point also blocks the third
    assert [(p.x, p.y) for p in total] == [(10, 10), (30, 30)]
    regs = np.zeros(5, mi.nat.REGION_DTYPE)
    regs["threshold_index"] = [0, 0, 2, 2, 5]
    assert [len(g) for g in mi._by_threshold(regs)] == [2, 2, 1]
    assert list(mi._by_threshold(regs[:0])) == []
    with pytest.raises(NotImplementedError):
        mi.GlobalFieldLocator.from_physical(1, 2, 3)
