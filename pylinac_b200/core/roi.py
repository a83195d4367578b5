"""``RectangleROI`` (core/roi.py:481-706): a rectangle on an image array with mean / std / min / max of its pixels.

The pixel selection is skimage.draw.polygon's in the reference (``pixels_flat``); here the statistics are device reductions over the
same pixel set (``epid_roi_stats``, csrc/roi.cu: integer pixel coordinates inside or on the boundary of the corner polygon the
reference builds, clipped to the image).  ``pixel_array`` (non-rotated ROIs) is a numpy view like the reference's."""
from __future__ import annotations

import numpy as np

from .. import _native as nat
from .geometry import Point, Rectangle


class RectangleROI(Rectangle):
    def __init__(self, array: np.ndarray, width: float, height: float, center, rotation: float = 0.0):
        if width < 2:
            raise ValueError(f"The width must be >= 2. Given {width}")
        if height < 2:
            raise ValueError(f"The height must be >= 2. Given {height}")
        super().__init__(width, height, center, rotation=rotation)
        self._array = array
        self._stats = None

    @classmethod
    def from_phantom_center(cls, array, width: float, height: float, angle: float, dist_from_center: float, phantom_center: Point,
                            rotation: float = 0.0):
        """core/roi.py:484-531"""
        y_shift = np.sin(np.deg2rad(angle)) * dist_from_center
        x_shift = np.cos(np.deg2rad(angle)) * dist_from_center
        return cls(array=array, width=width, height=height, center=Point(phantom_center.x + x_shift, phantom_center.y + y_shift),
                   rotation=rotation)

    def _polygon_xy(self) -> np.ndarray:
        """The corner list ``pixels_flat`` hands to skimage.draw.polygon (core/roi.py:646-656), as (x, y) pairs."""
        bl, br, tr, tl = self.bl_corner, self.br_corner, self.tr_corner, self.tl_corner
        return np.array([(bl.x, bl.y - 1), (br.x - 1, br.y - 1), (tr.x - 1, tr.y), (tl.x, tl.y)], dtype=np.float64)

    def _compute(self) -> dict:
        if self._stats is None:
            a = np.asarray(getattr(self._array, "array", self._array))
            if a.dtype not in nat._NP2DT:
                a = a.astype(np.float64)
            out = nat.roi_stats(nat.Context.default(), a, self._polygon_xy()[None])
            self._stats = {k: float(v[0, 0]) for k, v in out.items()}
        return self._stats

    @property
    def pixel_array(self) -> np.ndarray:
        if self.rotation != 0:
            raise ValueError("The pixel array cannot be reshaped into a 2D array when the rotation is not 0.")
        a = getattr(self._array, "array", self._array)
        return a[int(np.round(self.tl_corner.y)): int(np.round(self.bl_corner.y)), int(np.round(self.bl_corner.x)): int(np.round(self.br_corner.x))]

    @property
    def pixel_value(self) -> float:
        return self._compute()["mean"]

    @property
    def mean(self) -> float:
        return self._compute()["mean"]

    @property
    def std(self) -> float:
        return self._compute()["std"]

    @property
    def min(self) -> float:
        return self._compute()["min"]

    @property
    def max(self) -> float:
        return self._compute()["max"]

    def __repr__(self):
        return f"Rectangle ROI @ {self.center}; mean pixel: {self.pixel_value}"
