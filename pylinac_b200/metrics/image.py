"""2-D image metric plug-ins -- mirror of ``pylinac.metrics.image`` (metrics/image.py:38-76, 402-667, 959-983): ``MetricBase``,
``SizedDiskRegion`` / ``SizedDiskLocator`` (the BB finder) and ``WeightedCentroid``, computed through ``image.compute(metric)``.

The pixel work runs on the device: the disk locator is the threshold sweep / labelling / region-property kernel of the
Winston-Lutz pipeline exposed on its own (``epid_disk_locate``, csrc/wl.cu), the weighted centroid is a device reduction
(``epid_weighted_centroid``, csrc/roi.cu).  Plotting hooks are out of scope.
"""
from __future__ import annotations

import math
import weakref
from abc import ABC, abstractmethod
from typing import Any

import numpy as np

from .. import _native as nat
from ..core.geometry import Point
from .features import DEFAULT_CONDITIONS, conditions_mask


class MetricBase(ABC):
    """metrics/image.py:38-76"""

    unit: str = ""
    image_compatibility = None
    name: str

    def inject_image(self, image) -> None:
        if self.image_compatibility is not None and not isinstance(image, tuple(self.image_compatibility)):
            raise TypeError(f"Image must be one of {self.image_compatibility}")
        self.image = weakref.proxy(image)

    def context_calculate(self) -> Any:
        """Calculate the metric; a metric must not modify the image (hash check, metrics/image.py:61-71)."""
        img_hash = hash(self.image.array.tobytes())
        calculation = self.calculate()
        if hash(self.image.array.tobytes()) != img_hash:
            raise RuntimeError("A metric modified an image. This is not allowed as this could affect other, downstream metrics. "
                               "Change the calculate method to not modify the underlying image.")
        return calculation

    @abstractmethod
    def calculate(self) -> Any:
        ...

    def plot(self, axis, **kwargs) -> None:        # presentation: out of scope
        pass

    def plotly(self, fig, **kwargs) -> None:
        pass

    def additional_plots(self):
        pass


class DiskRegion:
    """The region properties the reference reads from scikit-image's ``RegionProperties`` for a detected disk (sample-window
    coordinates like ``regionprops``): area, area_filled, bbox, area_bbox, centroid, centroid_weighted, perimeter, area_convex,
    solidity -- with the older skimage aliases."""

    def __init__(self, row, k: int):
        self.area = float(row["r_area"][k])
        self.area_filled = self.filled_area = float(row["r_filled_area"][k])
        self.perimeter = float(row["r_perimeter"][k])
        self.area_convex = self.convex_area = float(row["r_convex_area"][k])
        self.solidity = self.area / self.area_convex if self.area_convex else float("nan")
        self.bbox = tuple(int(v) for v in row["r_bbox"][k])
        self.area_bbox = self.bbox_area = float((self.bbox[2] - self.bbox[0]) * (self.bbox[3] - self.bbox[1]))
        self.centroid = (float(row["r_centroid_y"][k]), float(row["r_centroid_x"][k]))
        self.centroid_weighted = self.weighted_centroid = (float(row["r_wcentroid_y"][k]), float(row["r_wcentroid_x"][k]))

    def __repr__(self):
        return f"DiskRegion(area={self.area:.0f}, bbox={self.bbox}, centroid_weighted=({self.centroid_weighted[0]:.2f}, {self.centroid_weighted[1]:.2f}))"


class SizedDiskRegion(MetricBase):
    """metrics/image.py:402-612.  Same constructors (pixels / ``from_physical`` / ``from_center`` / ``from_center_physical``) and the
    same unit bookkeeping as the reference, including that every conversion is applied IN PLACE on each ``calculate()`` call
    (:571-587; ``Point.__mul__`` mutates the expected position, core/geometry.py:190-196)."""

    def __init__(self, expected_position, search_window, radius: float, radius_tolerance: float,
                 detection_conditions=DEFAULT_CONDITIONS, invert: bool = True, name: str = "Disk Region", max_number: int = 1,
                 min_number: int = 1, min_separation_pixels: float = 5):
        self.expected_position = Point(expected_position)
        self.radius = radius
        self.radius_tolerance = radius_tolerance
        self.search_window = search_window
        self.detection_conditions = detection_conditions
        self.name = name
        self.invert = invert
        self.is_from_center = False
        self.is_from_physical = False
        self.max_number = max_number
        self.min_number = min_number
        self.min_separation = min_separation_pixels

    @classmethod
    def from_physical(cls, expected_position_mm, search_window_mm, radius_mm: float, radius_tolerance_mm: float,
                      detection_conditions=DEFAULT_CONDITIONS, invert: bool = True, name="Disk Region", max_number: int = 1,
                      min_number: int = 1, min_separation_mm: float = 5):
        inst = cls(expected_position_mm, search_window_mm, radius_mm, radius_tolerance_mm, detection_conditions, invert, name,
                   max_number, min_number, min_separation_mm)
        inst.is_from_physical = True
        return inst

    @classmethod
    def from_center(cls, expected_position, search_window, radius: float, radius_tolerance: float,
                    detection_conditions=DEFAULT_CONDITIONS, invert: bool = True, name="Disk Region", max_number: int = 1,
                    min_number: int = 1, min_separation_pixels: float = 5):
        inst = cls(expected_position, search_window, radius, radius_tolerance, detection_conditions, invert, name, max_number,
                   min_number, min_separation_pixels)
        inst.is_from_center = True
        return inst

    @classmethod
    def from_center_physical(cls, expected_position_mm, search_window_mm, radius_mm: float, radius_tolerance_mm: float = 0.25,
                             detection_conditions=DEFAULT_CONDITIONS, invert: bool = True, name="Disk Region", max_number: int = 1,
                             min_number: int = 1, min_separation_mm: float = 5):
        inst = cls(expected_position_mm, search_window_mm, radius_mm, radius_tolerance_mm, detection_conditions, invert, name,
                   max_number, min_number, min_separation_mm)
        inst.is_from_physical = True
        inst.is_from_center = True
        return inst

    def _locate(self) -> np.ndarray:
        from ..core import image as _image

        dpmm = self.image.dpmm
        if self.is_from_physical:
            self.expected_position * dpmm                                       # in place (:573)
            self.search_window = np.asarray(self.search_window) * dpmm
        else:
            self.min_separation /= dpmm
            self.radius /= dpmm
            self.radius_tolerance /= dpmm
        if self.is_from_center:
            self.expected_position.x += self.image.shape[1] / 2
            self.expected_position.y += self.image.shape[0] / 2
        p = nat.DiskParams()
        p.dpmm = float(dpmm)
        p.expected_x, p.expected_y = float(self.expected_position.x), float(self.expected_position.y)
        p.window_w, p.window_h = float(self.search_window[0]), float(self.search_window[1])
        p.radius_mm, p.tolerance_mm = float(self.radius), float(self.radius_tolerance)
        p.min_separation_px = float(self.min_separation * dpmm)
        p.invert = 1 if self.invert else 0
        p.max_number = int(self.max_number)
        p.conditions = conditions_mask(self.detection_conditions)
        frame = _image.frame_u16(self.image, "disk locator")
        row = nat.disk_locate(nat.Context.default(), frame, p)[0]
        status = int(row["status"])
        if status == 4:
            raise MemoryError("disk locator: the search window or a candidate region exceeds the device tile (EPID_WL_CAPACITY)")
        n = int(row["n_points"]) if status == 0 else 0
        if n < self.min_number:        # metrics/utils.py:181-185
            raise ValueError(f"Couldn't find the minimum number of disks in the image. Found {n}; required: {self.min_number}")
        return row

    def calculate(self) -> list[DiskRegion]:
        row = self._locate()
        self.x_offset = int(row["left"])
        self.y_offset = int(row["top"])
        self.points = [Point(float(row["x"][k]), float(row["y"][k])) for k in range(int(row["n_points"]))]
        self.boundaries = []           # plot-only outlines (metrics/utils.py:40-63): not produced
        return [DiskRegion(row, k) for k in range(int(row["n_regions"]))]


class SizedDiskLocator(SizedDiskRegion):
    """metrics/image.py:661-667: the weighted centroids of the detected disks as Points (image coordinates)."""

    def calculate(self) -> list[Point]:
        super().calculate()
        return self.points


class WeightedCentroid(MetricBase):
    """metrics/image.py:959-983"""

    def __init__(self, name: str = "Weighted Centroid"):
        self.name = name

    def calculate(self) -> Point:
        arr = np.asarray(self.image.array)
        cx, cy, total = nat.weighted_centroid(nat.Context.default(), arr if arr.dtype in nat._NP2DT else arr.astype(np.float64))
        if total[0] == 0:
            raise ValueError("Image is blank; cannot calculate weighted centroid")
        return Point(float(cx[0]), float(cy[0]))
