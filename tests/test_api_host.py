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
