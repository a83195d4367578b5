"""Profile metric plug-ins -- mirror of ``pylinac.metrics.profile`` (metrics/profile.py:22-335): the same class names, constructor
arguments, ``name`` / ``unit`` / ``full_name`` and ``calculate()`` semantics, injected into a profile by ``profile.compute(...)``.
Plotting hooks are out of scope.  The metrics are scalar formulas on a few hundred in-field samples of a 1-D profile whose
edges come from the device peak search (``core/profile.py``)."""
from __future__ import annotations

import math
from abc import ABC, abstractmethod
from typing import Any

import numpy as np

LEFT, RIGHT = "left", "right"


class ProfileMetric(ABC):
    """metrics/profile.py:22-58"""

    name: str
    unit: str = ""
    profile = None

    def __init__(self, color: str | None = None, linestyle: str | None = None):
        self.color = color
        self.linestyle = linestyle

    @property
    def full_name(self) -> str:
        return f"{self.name} ({self.unit})" if self.unit else self.name

    def inject_profile(self, profile) -> None:
        self.profile = profile

    def plot(self, axis) -> None:      # presentation: out of scope
        pass

    @abstractmethod
    def calculate(self) -> Any:
        ...


class FlatnessDifferenceMetric(ProfileMetric):
    """metrics/profile.py:61-104: 100 * (max - min) / (max + min) of the in-field values."""

    name = "Flatness (Difference)"
    unit = "%"

    def __init__(self, in_field_ratio: float = 0.8, color="g", linestyle="-."):
        self.in_field_ratio = in_field_ratio
        super().__init__(color=color, linestyle=linestyle)

    def calculate(self) -> float:
        fv = self.profile.field_values()      # the reference calls field_values() with its default ratio here
        return 100 * (fv.max() - fv.min()) / (fv.max() + fv.min())


class FlatnessRatioMetric(FlatnessDifferenceMetric):
    """metrics/profile.py:107-116"""

    name = "Flatness (Ratio)"

    def calculate(self) -> float:
        fv = self.profile.field_values()
        return 100 * fv.max() / fv.min()


class SymmetryPointDifferenceMetric(ProfileMetric):
    """metrics/profile.py:119-189"""

    unit = "%"
    name = "Point Difference Symmetry"

    def __init__(self, in_field_ratio: float = 0.8, color="magenta", linestyle="--", max_sym_range: float = 2,
                 min_sym_range: float = -2):
        self.in_field_ratio = in_field_ratio
        self.max_sym = max_sym_range
        self.min_sym = min_sym_range
        super().__init__(color=color, linestyle=linestyle)

    @staticmethod
    def _calc_point(lt: float, rt: float, cax: float) -> float:
        return 100 * (lt - rt) / cax

    @property
    def symmetry_values(self) -> list[float]:
        fv = self.profile.field_values(in_field_ratio=self.in_field_ratio)
        cax_value = self.profile.y_at_x(self.profile.center_idx)
        return [self._calc_point(lt, rt, cax_value) for lt, rt in zip(fv, fv[::-1])]

    def calculate(self) -> float:
        sv = self.symmetry_values
        return sv[int(np.argmax(np.abs(sv)))]


class SymmetryPointDifferenceQuotientMetric(SymmetryPointDifferenceMetric):
    """metrics/profile.py:192-213"""

    name = "Point Difference Quotient Symmetry"

    def __init__(self, in_field_ratio: float = 0.8, color="magenta", linestyle="--", max_sym_range: float = 105,
                 min_sym_range: float = 100):
        super().__init__(in_field_ratio, color, linestyle, max_sym_range, min_sym_range)

    @staticmethod
    def _calc_point(lt: float, rt: float, cax: float) -> float:
        return 100 * max((lt / rt), (rt / lt))


class SymmetryAreaMetric(ProfileMetric):
    """metrics/profile.py:216-246"""

    name = "Symmetry (Area)"

    def __init__(self, in_field_ratio: float = 0.8):
        self.in_field_ratio = in_field_ratio
        super().__init__()

    def calculate(self) -> float:
        _, _, width = self.profile.field_indices(in_field_ratio=self.in_field_ratio)
        fv = self.profile.field_values(self.in_field_ratio)
        area_left = np.sum(fv[: math.floor(width / 2) + 1])
        area_right = np.sum(fv[math.ceil(width / 2):])
        return 100 * (area_left - area_right) / (area_left + area_right)


class PenumbraLeftMetric(ProfileMetric):
    """metrics/profile.py:249-289: distance between the lower / upper fractions of twice the edge value on one side."""

    unit = "mm"
    name = "Left Penumbra"
    side = LEFT

    def __init__(self, lower: float = 20, upper: float = 80, color="pink", ls="-."):
        self.lower = lower
        self.upper = upper
        super().__init__(color=color, linestyle=ls)

    def calculate(self) -> float:
        edge = self.profile.field_edge_idx(side=self.side)
        edge_value = self.profile.y_at_x(edge)
        self.lower_index = self.profile.x_at_y(y=edge_value * 2 * self.lower / 100, side=self.side)
        self.upper_index = self.profile.x_at_y(y=edge_value * 2 * self.upper / 100, side=self.side)
        return abs(self.upper_index - self.lower_index) / self.profile.dpmm


class PenumbraRightMetric(PenumbraLeftMetric):
    side = RIGHT
    name = "Right Penumbra"


class CAXToLeftEdgeMetric(ProfileMetric):
    """metrics/profile.py:292-314"""

    name = "CAX to Left Beam Edge"
    unit = "mm"

    def __init__(self, color: str | None = "cyan", linestyle: str | None = "--"):
        super().__init__(color=color, linestyle=linestyle)

    def calculate(self) -> float:
        return (self.profile.cax_index - self.profile.field_edge_idx(side=LEFT)) / self.profile.dpmm


class CAXToRightEdgeMetric(CAXToLeftEdgeMetric):
    """metrics/profile.py:317-324"""

    name = "CAX to Right Beam Edge"

    def calculate(self) -> float:
        return (self.profile.field_edge_idx(side=RIGHT) - self.profile.cax_index) / self.profile.dpmm
