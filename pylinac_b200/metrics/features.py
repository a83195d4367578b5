"""Detection conditions of the disk / BB finder -- the names of ``pylinac.metrics.features`` (metrics/features.py:7-68).

In the reference these are callables applied to scikit-image ``RegionProperties``; here the region analysis runs inside the CUDA
BB finder (csrc/wl.cu), so each condition is a marker object that selects the corresponding device-side test.  They are still
callable on the region records the locator returns (``DiskRegion``), with the reference's formulas, for user code that filters
regions itself.
"""
from __future__ import annotations

import math


class _Condition:
    def __init__(self, name: str, bit: int, fn):
        self.__name__ = name
        self.bit = bit
        self._fn = fn

    def __call__(self, region, *args, **kwargs) -> bool:
        return self._fn(region, **kwargs)

    def __repr__(self):
        return f"<detection condition {self.__name__}>"


def _is_symmetric(region, **kwargs) -> bool:  # :7-14
    ymin, xmin, ymax, xmax = region.bbox
    y, x = abs(ymax - ymin), abs(xmax - xmin)
    return not (x > max(y * 1.05, y + 3) or x < min(y * 0.95, y - 3))


def _is_right_size_bb(region, **kwargs) -> bool:  # :34-45
    bb_area = region.area_filled / (kwargs["dpmm"] ** 2)
    bb_size, tolerance = kwargs["bb_size"], kwargs["tolerance"]
    larger = math.pi * (bb_size + tolerance) ** 2
    smaller = max((math.pi * (bb_size - tolerance) ** 2, 2))
    return smaller < bb_area < larger


def _is_solid(region, **kwargs) -> bool:  # :48-52
    return region.solidity > 0.9


def _is_round(region, **kwargs) -> bool:  # :55-59
    expected = math.pi / 4
    actual = region.filled_area / region.bbox_area
    return expected * 1.2 > actual > expected * 0.8


def _is_right_circumference(region, **kwargs) -> bool:  # :62-68
    upper = 2 * math.pi * (kwargs["bb_size"] + kwargs["tolerance"])
    lower = 2 * math.pi * (kwargs["bb_size"] - kwargs["tolerance"])
    return upper > region.perimeter / kwargs["dpmm"] > lower


is_right_size_bb = _Condition("is_right_size_bb", 1, _is_right_size_bb)
is_round = _Condition("is_round", 2, _is_round)
is_right_circumference = _Condition("is_right_circumference", 4, _is_right_circumference)
is_symmetric = _Condition("is_symmetric", 8, _is_symmetric)
is_solid = _Condition("is_solid", 16, _is_solid)



def _is_square(region, **kwargs) -> bool:  # :85-88
    return region.filled_area / region.bbox_area > 0.8


def _is_right_square_perimeter(region, **kwargs) -> bool:  # :71-82
    actual = region.perimeter / kwargs["dpmm"]
    upper = 1.20 * 2 * (kwargs["field_width_mm"] + kwargs["field_tolerance_mm"]) + 2 * (kwargs["field_height_mm"] + kwargs["field_tolerance_mm"])
    lower = 2 * (kwargs["field_width_mm"] - kwargs["field_tolerance_mm"]) + 2 * (kwargs["field_height_mm"] - kwargs["field_tolerance_mm"])
    return upper > actual > lower


def _is_right_area_square(region, **kwargs) -> bool:  # :91-101
    field_area = region.area_filled / (kwargs["dpmm"] ** 2)
    low = (kwargs["field_width_mm"] - kwargs["field_tolerance_mm"]) * (kwargs["field_height_mm"] - kwargs["field_tolerance_mm"])
    high = (kwargs["field_width_mm"] + kwargs["field_tolerance_mm"]) * (kwargs["field_height_mm"] + kwargs["field_tolerance_mm"])
    return low < field_area < high


def _is_modest_size(region, **kwargs) -> bool:  # winston_lutz.py:598-606
    bb_area = region.area_filled / (kwargs["dpmm"] ** 2)
    bb_size = kwargs["bb_size"]
    return max((math.pi * ((bb_size - 2) / 2) ** 2, 2)) < bb_area < math.pi * ((bb_size + 2) / 2) ** 2


def _is_right_square_size(region, **kwargs) -> bool:  # winston_lutz.py:615-621
    field_area = region.area_filled / (kwargs["dpmm"] ** 2)
    rad_size = max((kwargs["rad_size"], 5))
    return (rad_size - 5) ** 2 < field_area < (rad_size + 5) ** 2


is_modest_size = _Condition("is_modest_size", 32, _is_modest_size)
is_square = _Condition("is_square", 64, _is_square)
is_right_square_size = _Condition("is_right_square_size", 128, _is_right_square_size)
is_right_square_perimeter = _Condition("is_right_square_perimeter", 256, _is_right_square_perimeter)
is_right_area_square = _Condition("is_right_area_square", 512, _is_right_area_square)

DEFAULT_CONDITIONS = (is_right_size_bb, is_round, is_right_circumference, is_symmetric, is_solid)


def conditions_mask(conditions) -> int:
    mask = 0
    for c in conditions:
        bit = getattr(c, "bit", None)
        if bit is None:
            raise NotImplementedError(f"detection condition {c!r} has no device-side implementation; use the conditions of "
                                      "pylinac_b200.metrics.features")
        mask |= bit
    return mask
