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


_NORM_CODE = {Normalization.NONE: 0, Normalization.GEOMETRIC_CENTER: 1, Normalization.BEAM_CENTER: 2, Normalization.MAX: 3}


class SingleProfile(ProfileMixin):
    """core/profile.py:1119-1937 -- same constructor and query methods; the numerics (interpolation, grounding, normalisation,
    FWXM / inflection edges, penumbra, field data) are the device engine of csrc/field.cu behind ``epid_single_profile``.
    Supported: interpolation NONE / LINEAR, edge detection FWHM / INFLECTION_DERIVATIVE, x_values = range(len(values))."""

    def __init__(self, values, dpmm: float | None = None, interpolation=Interpolation.LINEAR, ground: bool = True,
                 interpolation_resolution_mm: float = 0.1, interpolation_factor: float = 10,
                 normalization_method=Normalization.BEAM_CENTER, edge_detection_method=Edge.FWHM,
                 edge_smoothing_ratio: float = 0.003, hill_window_ratio: float = 0.1, x_values=None, centering=Centering.BEAM_CENTER):
        self._interp_method = interpolation if isinstance(interpolation, Interpolation) else Interpolation(interpolation)
        self._norm_method = normalization_method if isinstance(normalization_method, Normalization) else Normalization(normalization_method)
        self._edge_method = edge_detection_method if isinstance(edge_detection_method, Edge) else Edge(edge_detection_method)
        self._centering = centering if isinstance(centering, Centering) else Centering(centering)
        if self._interp_method == Interpolation.SPLINE:
            raise NotImplementedError("Interpolation.SPLINE (cubic interp1d) is outside the accelerated hot path")
        if self._edge_method == Edge.INFLECTION_HILL:
            raise NotImplementedError("Edge.INFLECTION_HILL (Hill-function fits) is outside the accelerated hot path")
        raw = np.asarray(values, dtype=np.float64)
        if raw.ndim != 1:
            raise ValueError("Profile values must be 1-D")
        if x_values is not None and not np.array_equal(np.asarray(x_values), np.arange(len(raw))):
            raise NotImplementedError("custom x_values are outside the accelerated hot path")
        self._raw = raw
        self.dpmm = dpmm
        self._interpolation_res = interpolation_resolution_mm
        self._interpolation_factor = interpolation_factor
        self._ground = ground
        self._edge_smoothing_ratio = edge_smoothing_ratio
        self._hill_window_ratio = hill_window_ratio
        sp = nat.SpParams()
        sp.dpmm = float(dpmm) if dpmm else 0.0
        sp.interpolation = 0 if self._interp_method == Interpolation.NONE else 1
        sp.interpolation_resolution_mm = float(interpolation_resolution_mm)
        sp.interpolation_factor = float(interpolation_factor)
        sp.ground = 1 if ground else 0
        sp.normalization = _NORM_CODE[self._norm_method]
        sp.edge = 0 if self._edge_method == Edge.FWHM else 1
        sp.centering = 2 if self._centering == Centering.GEOMETRIC_CENTER else 1
        sp.edge_smoothing_ratio = float(edge_smoothing_ratio)
        self._params = sp
        self._cache = {}
        r, vals, _ = self._query()
        if int(r["status"]) != 0:
            raise IndexError("no peak was found in the profile")       # what find_peaks(...)[0] raises in the reference
        self.values = vals
        self.x_indices = np.linspace(float(r["x_start"]), float(r["x_stop"]), num=int(r["n"]))

    def _query(self, fwxm_x=50.0, penumbra=(20.0, 80.0), in_field_ratio=0.8, slope_exclusion_ratio=0.2):
        key = (float(fwxm_x), float(penumbra[0]), float(penumbra[1]), float(in_field_ratio), float(slope_exclusion_ratio))
        if key not in self._cache:
            self._cache[key] = nat.single_profile(nat.Context.default(), self._raw, self._params, fwxm_x=fwxm_x, penumbra=penumbra,
                                                  in_field_ratio=in_field_ratio, slope_exclusion_ratio=slope_exclusion_ratio)
        return self._cache[key]

    # -- core/profile.py:1373-1409
    def geometric_center(self) -> dict:
        r, _, _ = self._query()
        return {"index (exact)": float(r["geometric_center_index"]), "value (exact)": float(r["geometric_center_value"])}

    def beam_center(self) -> dict:
        r, _, _ = self._query()
        if not r["beam_ok"]:
            raise IndexError("no field edges were found")
        idx = float(r["beam_center_index"])
        return {"index (rounded)": int(round(idx)), "index (exact)": idx, "value (@rounded)": float(r["beam_center_value_at_rounded"])}

    # -- core/profile.py:1411-1461
    def fwxm_data(self, x: float = 50) -> dict:
        if not 0 <= x <= 100:
            raise ValueError("x must be between 0 and 100")
        r, _, _ = self._query(fwxm_x=x)
        if not r["fwxm_ok"]:
            raise IndexError("no peak was found in the profile")
        left, right = float(r["fwxm_left"]), float(r["fwxm_right"])
        width = right - left
        center = (right - left) / 2 + left
        data = {"width (exact)": width, "width (rounded)": int(round(width)), "center index (rounded)": int(round(center)),
                "center index (exact)": center, "center value (@rounded)": float(r["fwxm_center_value_at_rounded"]),
                "left index (exact)": left, "left index (rounded)": int(round(left)),
                "left value (@rounded)": float(r["fwxm_left_value_at_rounded"]), "right index (exact)": right,
                "right index (rounded)": int(round(right)), "right value (@rounded)": float(r["fwxm_right_value_at_rounded"])}
        if self.dpmm:
            data["width (exact) mm"] = width / self.dpmm
            data["left distance (exact) mm"] = abs(center - left) / self.dpmm
            data["right distance (exact) mm"] = abs(right - center) / self.dpmm
        return data

    # -- core/profile.py:1635-1721
    def inflection_data(self) -> dict:
        if self._edge_method == Edge.FWHM:
            raise ValueError("FWHM edge method does not have inflection points. Use a different edge detection method")
        r, _, _ = self._query()
        if not r["infl_ok"]:
            raise IndexError("no inflection points were found")
        left, right = float(r["infl_left"]), float(r["infl_right"])
        return {"left index (rounded)": int(round(left)), "left index (exact)": left, "right index (rounded)": int(round(right)),
                "right index (exact)": right, "left value (@rounded)": float(r["infl_left_value_rounded"]),
                "left value (@exact)": float(r["infl_left_value_exact"]), "right value (@rounded)": float(r["infl_right_value_rounded"]),
                "right value (@exact)": float(r["infl_right_value_exact"])}

    # -- core/profile.py:1723-1907
    def penumbra(self, lower: int = 20, upper: int = 80) -> dict:
        if lower > upper:
            raise ValueError("Upper penumbra value must be larger than the lower penumbra value")
        r, _, _ = self._query(penumbra=(lower, upper))
        if not r["pen_ok"]:
            raise IndexError("no field edges were found")
        data = {f"left {lower}% index (exact)": float(r["pen_left_lower"]), f"left {upper}% index (exact)": float(r["pen_left_upper"]),
                f"right {lower}% index (exact)": float(r["pen_right_lower"]), f"right {upper}% index (exact)": float(r["pen_right_upper"]),
                "left penumbra width (exact)": abs(float(r["pen_left_upper"]) - float(r["pen_left_lower"])),
                "right penumbra width (exact)": abs(float(r["pen_right_upper"]) - float(r["pen_right_lower"]))}
        if self.dpmm:
            data["left penumbra width (exact) mm"] = data["left penumbra width (exact)"] / self.dpmm
            data["right penumbra width (exact) mm"] = data["right penumbra width (exact)"] / self.dpmm
        return data

    # -- core/profile.py:1463-1633
    def field_data(self, in_field_ratio: float = 0.8, slope_exclusion_ratio: float = 0.2) -> dict:
        if slope_exclusion_ratio >= in_field_ratio:
            raise ValueError("The exclusion region must be smaller than the field ratio")
        r, _, fv = self._query(in_field_ratio=in_field_ratio, slope_exclusion_ratio=slope_exclusion_ratio)
        if not r["fd_ok"]:
            raise IndexError("no field edges were found")
        g = lambda k: float(r[k])
        data = {"width (exact)": g("fd_width"), "beam center index (exact)": g("fd_beam_center"),
                "beam center index (rounded)": int(round(g("fd_beam_center"))), "beam center value (@rounded)": g("fd_beam_center_value"),
                "cax index (exact)": g("fd_cax"), "cax index (rounded)": int(round(g("fd_cax"))), "cax value (@rounded)": g("fd_cax_value"),
                "left index (exact)": g("fd_left"), "left index (rounded)": int(round(g("fd_left"))), "left value (@rounded)": g("fd_left_value"),
                "left slope": g("fd_left_slope"), "left intercept": g("fd_left_intercept"), "right slope": g("fd_right_slope"),
                "right intercept": g("fd_right_intercept"), "left inner index (exact)": g("fd_inner_left"),
                "left inner index (rounded)": int(round(g("fd_inner_left"))), "right inner index (exact)": g("fd_inner_right"),
                "right inner index (rounded)": int(round(g("fd_inner_right"))), '"top" index (exact)': g("fd_top_index"),
                '"top" index (rounded)': int(round(g("fd_top_index"))), '"top" value (@exact)': g("fd_top_value"),
                "top params": np.array(r["fd_top_params"], dtype=float), "right index (exact)": g("fd_right"),
                "right index (rounded)": int(round(g("fd_right"))), "right value (@rounded)": g("fd_right_value"), "field values": fv}
        if self.dpmm:
            d = self.dpmm
            data["width (exact) mm"] = data["width (exact)"] / d
            data["left slope (%/mm)"] = data["left slope"] * d * 100
            data["right slope (%/mm)"] = data["right slope"] * d * 100
            data["left distance->beam center (exact) mm"] = abs(data["beam center index (exact)"] - data["left index (exact)"]) / d
            data["right distance->beam center (exact) mm"] = abs(data["right index (exact)"] - data["beam center index (exact)"]) / d
            data["left distance->CAX (exact) mm"] = abs(data["cax index (exact)"] - data["left index (exact)"]) / d
            data["right distance->CAX (exact) mm"] = abs(data["cax index (exact)"] - data["right index (exact)"]) / d
            data["left distance->top (exact) mm"] = abs(data['"top" index (exact)'] - data["left index (exact)"]) / d
            data["right distance->top (exact) mm"] = abs(data['"top" index (exact)'] - data["right index (exact)"]) / d
            data['"top"->beam center (exact) mm'] = (data['"top" index (exact)'] - data["beam center index (exact)"]) / d
            data['"top"->CAX (exact) mm'] = abs(data['"top" index (exact)'] - data["cax index (exact)"]) / d
        return data

    # -- core/profile.py:1909-1937
    def field_calculation(self, in_field_ratio: float = 0.8, calculation: str = "mean", slope_exclusion_ratio: float = 0.2):
        if calculation not in ("mean", "median", "max", "min", "area"):
            raise ValueError("calculation must be one of mean, median, max, min, area")
        fv = self.field_data(in_field_ratio, slope_exclusion_ratio=slope_exclusion_ratio)["field values"]
        if calculation == "max":
            return float(fv.max())
        if calculation == "min":
            return float(fv.min())
        srt = np.sort(fv)                     # a few hundred field values: selection / pairwise sum on the host-resident result
        if calculation == "median":
            m = len(srt)
            return float(srt[m // 2] if m % 2 else (srt[m // 2 - 1] + srt[m // 2]) / 2)
        if calculation == "mean":
            return float(fv.mean())
        return None


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
