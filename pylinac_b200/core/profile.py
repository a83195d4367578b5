"""1-D profile operators -- mirror of the parts of ``pylinac/core/profile.py`` the hot path uses.

All arithmetic on profile values runs in CUDA through the C-ABI (``epid_find_peaks`` and the element-wise / stencil operators of
``core.array_utils``); index bookkeeping stays in Python.  There is no numpy / scipy compute fallback.

    enums Interpolation / Normalization / Edge / Centering        core/profile.py:160-192
    find_peaks                                                    core/profile.py:2545-2649
    MultiProfile.find_peaks / find_valleys / find_fwxm_peaks      core/profile.py:2021-2176
    ProfileMixin (invert / normalize / stretch / ground / filter) core/profile.py:86-159
    FWXMProfile.field_edge_idx / center_idx / field_width_px      core/profile.py:322-344, 582-611
"""
from __future__ import annotations

import enum

import numpy as np

from .. import _native as nat
from . import array_utils as utils
from .geometry import Point


class Interpolation(enum.Enum):
    """core/profile.py:160-165"""

    NONE = None
    LINEAR = "Linear"
    SPLINE = "Spline"


class Normalization(enum.Enum):
    """core/profile.py:168-174"""

    NONE = None
    GEOMETRIC_CENTER = "Geometric center"
    BEAM_CENTER = "Beam center"
    MAX = "Max"


class Edge(enum.Enum):
    """core/profile.py:177-182"""

    FWHM = "FWHM"
    INFLECTION_DERIVATIVE = "Inflection Derivative"
    INFLECTION_HILL = "Inflection Hill"


class Centering(enum.Enum):
    """core/profile.py:185-190"""

    MANUAL = "Manual"
    BEAM_CENTER = "Beam center"
    GEOMETRIC_CENTER = "Geometric center"


def find_peaks(values, threshold=-np.inf, peak_separation=0, max_number=None, fwxm_height=0.5, min_width=0,
               search_region=(0.0, 1.0), peak_sort="prominences", required_prominence=None):
    """core/profile.py:2545-2623 -- scipy.signal.find_peaks semantics, executed by the CUDA kernel behind epid_find_peaks."""
    return nat.find_peaks(nat.Context.default(), np.asarray(values, dtype=np.float64), threshold=threshold,
                          peak_separation=peak_separation, max_number=max_number, fwxm_height=fwxm_height, min_width=min_width,
                          search_region=search_region, peak_sort=peak_sort, required_prominence=required_prominence)


class ProfileMixin:
    """core/profile.py:86-159: in-place value operators of 1-D profiles (device arithmetic via core.array_utils)."""

    values: np.ndarray

    def invert(self) -> None:
        self.values = utils.invert(self.values)

    def normalize(self, norm_val="max") -> None:
        self.values = utils.normalize(self.values, value=None if norm_val == "max" else norm_val)

    def stretch(self, min: float = 0, max: float = 1) -> None:
        self.values = utils.stretch(self.values, min=min, max=max)

    def ground(self) -> float:
        mn = float(np.asarray(self.values).min())
        self.values = utils.ground(self.values)
        return mn

    def filter(self, size: float = 0.05, kind: str = "median") -> None:
        self.values = utils.filter(self.values, size=size, kind=kind)


class MultiProfile(ProfileMixin):
    """core/profile.py:2002-2176"""

    def __init__(self, values):
        self.values = np.asarray(values)
        self.peaks: list[Point] = []
        self.valleys: list[Point] = []

    def find_peaks(self, threshold=0.3, min_distance=0.05, max_number=None, search_region=(0.0, 1.0), peak_sort="prominences"):
        idx, props = find_peaks(self.values, threshold=threshold, peak_separation=min_distance, max_number=max_number,
                                search_region=search_region, peak_sort=peak_sort)
        self.peaks = [Point(value=v, idx=i) for i, v in zip(idx, props["peak_heights"])]
        return idx, props["peak_heights"]

    def find_valleys(self, threshold=0.3, min_distance=0.05, max_number=None, search_region=(0.0, 1.0)):
        idx, props = find_peaks(utils_negate(self.values), threshold=threshold, peak_separation=min_distance, max_number=max_number,
                                search_region=search_region)
        vals = np.asarray(self.values)[idx]
        self.valleys = [Point(value=v, idx=i) for i, v in zip(idx, vals)]
        return idx, vals

    def find_fwxm_peaks(self, threshold=0.3, min_distance=0.05, max_number=None, search_region=(0.0, 1.0),
                        peak_sort="prominences", required_prominence=None):
        _, props = find_peaks(self.values, threshold=threshold, peak_separation=min_distance, max_number=max_number,
                              search_region=search_region, peak_sort=peak_sort, required_prominence=required_prominence)
        idxs = [int(round(lt + (rt - lt) / 2)) for lt, rt in zip(props["left_ips"], props["right_ips"])]
        vals = [np.asarray(self.values)[i] for i in idxs]
        self.peaks = [Point(value=v, idx=i) for i, v in zip(idxs, vals)]
        return np.array(idxs), np.array(vals)


class CircleProfile(MultiProfile):
    """core/profile.py:2179-2402: a profile sampled on a circle (nearest neighbour), peaks mapped back to image coordinates."""

    def __init__(self, center, radius: float, image_array: np.ndarray, start_angle: float = 0, ccw: bool = True,
                 sampling_ratio: float = 1.0):
        self.center = Point(center)
        self.radius = radius
        self.image_array = image_array
        self.start_angle = start_angle
        self.ccw = ccw
        self.sampling_ratio = sampling_ratio
        prof, self.x_locations, self.y_locations = self._sample()
        super().__init__(prof)

    def _sample(self):
        return nat.circle_profile(nat.Context.default(), self.image_array, (self.center.x, self.center.y), self.radius,
                                  start_angle=self.start_angle, ccw=self.ccw, sampling_ratio=self.sampling_ratio)

    @property
    def size(self) -> float:
        return np.pi * self.radius * 2 * self.sampling_ratio

    def _map_peaks(self) -> None:
        for peak in self.peaks:
            peak.x = self.x_locations[int(peak.idx)]
            peak.y = self.y_locations[int(peak.idx)]

    def find_peaks(self, threshold=0.3, min_distance=0.05, max_number=None, search_region=(0.0, 1.0)):
        out = super().find_peaks(threshold, min_distance, max_number, search_region)
        self._map_peaks()
        return out

    def find_valleys(self, threshold=0.3, min_distance=0.05, max_number=None, search_region=(0.0, 1.0)):
        out = super().find_valleys(threshold, min_distance, max_number, search_region)
        self._map_peaks()
        return out

    def find_fwxm_peaks(self, threshold=0.3, min_distance=0.05, max_number=None, search_region=(0.0, 1.0)):
        out = super().find_fwxm_peaks(threshold, min_distance, max_number, search_region=search_region)
        self._map_peaks()
        return out

    def roll(self, amount: int) -> None:
        self.values = np.roll(self.values, -amount)
        self.x_locations = np.roll(self.x_locations, -amount)
        self.y_locations = np.roll(self.y_locations, -amount)


class CollapsedCircleProfile(CircleProfile):
    """core/profile.py:2405-2483: mean of `num_profiles` circle profiles in a band of relative width `width_ratio`."""

    def __init__(self, center, radius: float, image_array: np.ndarray, start_angle: float = 0, ccw: bool = True,
                 sampling_ratio: float = 1.0, width_ratio: float = 0.1, num_profiles: int = 20):
        if not 0 <= width_ratio <= 1:
            raise ValueError("width_ratio must be between 0 and 1")
        self.width_ratio = width_ratio
        self.num_profiles = num_profiles
        super().__init__(center, radius, image_array, start_angle, ccw, sampling_ratio)

    def _sample(self):
        return nat.circle_profile(nat.Context.default(), self.image_array, (self.center.x, self.center.y), self.radius,
                                  start_angle=self.start_angle, ccw=self.ccw, sampling_ratio=self.sampling_ratio, collapsed=True,
                                  width_ratio=self.width_ratio, num_profiles=self.num_profiles)

    @property
    def size(self) -> float:
        return np.pi * self.radius * (1 + self.width_ratio) * 2 * self.sampling_ratio


def utils_negate(values) -> np.ndarray:
    """-values for find_valleys: a sign flip of the stored samples (no arithmetic on magnitudes)."""
    return np.negative(np.asarray(values, dtype=np.float64))


class FWXMProfile(ProfileMixin):
    """core/profile.py:195-344, 578-611: FWXM field edges of a single-peak profile (x_values = sample indices)."""

    def __init__(self, values, x_values=None, ground: bool = False, normalization=Normalization.NONE, fwxm_height: float = 50):
        self.values = np.asarray(values, dtype=np.float64)
        if x_values is not None and not np.array_equal(np.asarray(x_values), np.arange(len(self.values))):
            raise NotImplementedError("custom x_values are outside the accelerated hot path")
        self.fwxm_height = fwxm_height
        if ground:
            self.ground()
        norm = Normalization(normalization) if not isinstance(normalization, Normalization) else normalization
        if norm == Normalization.MAX:
            self.normalize("max")
        elif norm != Normalization.NONE:
            raise NotImplementedError("only Normalization.NONE / MAX are available on FWXMProfile")

    def field_edge_idx(self, side: str) -> float:
        _, props = find_peaks(self.values, fwxm_height=self.fwxm_height / 100, max_number=1)
        return float(props["left_ips"][0] if side == "left" else props["right_ips"][0])

    @property
    def center_idx(self) -> float:
        left, right = self.field_edge_idx("left"), self.field_edge_idx("right")
        return abs(right - left) / 2 + left

    @property
    def field_width_px(self) -> float:
        return self.field_edge_idx("right") - self.field_edge_idx("left")
