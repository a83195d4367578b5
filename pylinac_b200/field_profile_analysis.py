"""Field analysis via profiles and metric plug-ins -- drop-in for ``pylinac.field_profile_analysis.FieldProfileAnalysis``
(field_profile_analysis.py:90-220, 307-368) on top of the GPU primitives.

What runs on the GPU: the histogram inversion check, the row / column sums of the frame (field centre), the strip sums that
become the x / y profiles (exact integer sums of the uint16 pixels, one ``epid_frame_stats`` call per strip), the FWXM peak
search and the gaussian smoothing of the edge search.  The metric formulas (flatness, symmetry, penumbra, CAX distances:
``pylinac_b200/metrics/profile.py``) are scalar host arithmetic on the few hundred in-field samples of the two profiles.

Supported edge types: FWHM and INFLECTION_DERIVATIVE (the reference's default; its scipy minimisation of the cubic interpolant is
replaced by the closed-form stationary point of the same spline).  INFLECTION_HILL (Hill-function curve fit) and the central ROI
statistics (scikit-image polygon rasterisation) are outside the accelerated path.
"""
from __future__ import annotations

import copy
from collections.abc import Sequence

import numpy as np

from . import _native as nat
from .core import image
from .core.profile import Centering, Edge, FWXMProfilePhysical, InflectionDerivativeProfilePhysical, Normalization
from .core.warnings import capture_warnings
from .core.utilities import ResultBase, ResultsDataMixin, convert_to_enum
from .metrics.profile import (CAXToLeftEdgeMetric, CAXToRightEdgeMetric, FlatnessDifferenceMetric, PenumbraLeftMetric,
                              PenumbraRightMetric, ProfileMetric, SymmetryPointDifferenceMetric)


class FieldProfileResult(ResultBase):
    """field_profile_analysis.py:33-71 (without the central ROI statistics)."""

    x_metrics: dict
    y_metrics: dict
    normalization: str
    edge_type: str
    centering: str


def default_metrics() -> tuple[ProfileMetric, ...]:
    """field_profile_analysis.py:73-80 (fresh instances: a metric object keeps the profile it was injected into)."""
    return (FlatnessDifferenceMetric(), SymmetryPointDifferenceMetric(), PenumbraRightMetric(), PenumbraLeftMetric(),
            CAXToLeftEdgeMetric(), CAXToRightEdgeMetric())


PROFILES = {Edge.FWHM: FWXMProfilePhysical, Edge.INFLECTION_DERIVATIVE: InflectionDerivativeProfilePhysical}


class NotAnalyzed(Exception):
    pass


@capture_warnings
class FieldProfileAnalysis(ResultsDataMixin[FieldProfileResult]):
    """field_profile_analysis.py:90-368 -- same constructor / analyze() keywords."""

    def __init__(self, path, **kwargs):
        if isinstance(path, np.ndarray):
            self.image = image.ArrayImage(path, **kwargs)
        elif isinstance(path, image.BaseImage):
            self.image = path
        else:
            self.image = image.load(path, **kwargs)
        self._is_analyzed = False
        self.image.check_inversion_by_histogram()

    # ---- device helpers: exact integer sums of frame strips
    def _frame_u16(self) -> np.ndarray:
        return image.frame_u16(self.image, "GPU field-profile")

    def analyze(self, centering=Centering.BEAM_CENTER, position: tuple[float, float] = (0.5, 0.5), x_width: float = 0.0,
                y_width: float = 0.0, normalization=Normalization.NONE, edge_type=Edge.INFLECTION_DERIVATIVE, invert: bool = False,
                ground: bool = True, metrics: Sequence[ProfileMetric] | None = None, **kwargs) -> None:
        """field_profile_analysis.py:123-194"""
        if invert:
            self.image.invert()
        self._normalization = convert_to_enum(normalization, Normalization)
        self._edge_type = convert_to_enum(edge_type, Edge)
        self._centering = convert_to_enum(centering, Centering)
        if self._edge_type not in PROFILES:
            raise NotImplementedError("Edge.INFLECTION_HILL (Hill-function curve fit) is outside the accelerated path")
        metrics = default_metrics() if metrics is None else metrics
        ctx = nat.Context.default()
        batch = nat.Batch.upload(ctx, self._frame_u16()[None])
        try:
            x_values, y_values = self._get_profile_values(ctx, batch, position, x_width, y_width)
        finally:
            batch.free()
        cls = PROFILES[self._edge_type]
        self.x_profile = cls(values=x_values, dpmm=self.image.dpmm, normalization=normalization, ground=ground, **kwargs)
        self.x_profile.compute(metrics=metrics)
        self.y_profile = cls(values=y_values, dpmm=self.image.dpmm, normalization=normalization, ground=ground, **kwargs)
        self.y_profile.compute(metrics=copy.deepcopy(metrics))
        self._is_analyzed = True

    def _get_profile_values(self, ctx, batch, position, x_width, y_width):
        """field_profile_analysis.py:307-341: strips around (x, y); mean over the strip = exact integer sum / row count"""
        h, w = self.image.shape
        x, y = self._get_x_y_position(ctx, batch, position)
        if x_width > 1 or x_width < 0 or y_width > 1 or y_width < 0:
            raise ValueError("Width must be between 0 and 1")
        top = round(y - h * x_width / 2 - 1)
        bottom = round(max(y + h * x_width / 2, top + 2))
        left = round(x - w * y_width / 2 - 1)
        right = round(max(x + w * y_width / 2, left + 2))
        t, b = max(top, 0), min(bottom, h)        # numpy slicing clips (negative starts would wrap: not reached for centred fields)
        l, r = max(left, 0), min(right, w)
        self._strip_rows, self._strip_cols = (t, b), (l, r)
        xs = nat.frame_stats(ctx, batch, view=(t, 0, b - t, w))["colsum"][0] / (b - t)
        ys = nat.frame_stats(ctx, batch, view=(0, l, h, r - l))["rowsum"][0] / (r - l)
        return xs, ys

    def _get_x_y_position(self, ctx, batch, position):
        """field_profile_analysis.py:343-368"""
        if self._centering != Centering.MANUAL:
            st = nat.frame_stats(ctx, batch)
            cls = PROFILES[self._edge_type]
            v_p = cls(values=st["colsum"][0], dpmm=self.image.dpmm)
            h_p = cls(values=st["rowsum"][0], dpmm=self.image.dpmm)
            if self._centering == Centering.BEAM_CENTER:
                return v_p.center_idx, h_p.center_idx
            return v_p.cax_index, h_p.cax_index
        if len(position) != 2:
            raise ValueError("Position must be a tuple of two values")
        if any(pos < 0 or pos > 1 for pos in position):
            raise ValueError("Position values must be between 0 and 1")
        return self.image.shape[1] * position[1], self.image.shape[0] * position[0]

    def _generate_results_data(self) -> FieldProfileResult:
        if not self._is_analyzed:
            raise NotAnalyzed("Image is not analyzed yet. Use analyze() first.")

        def pack(p):
            return {k: float(v) for k, v in p.metric_values.items()} | {"Field Width (mm)": p.field_width_mm, "values": p.values.tolist()}

        return FieldProfileResult(edge_type=self._edge_type.value, normalization=str(self._normalization.value),
                                  centering=self._centering.value, x_metrics=pack(self.x_profile), y_metrics=pack(self.y_profile))

    def results(self) -> str:
        """field_profile_analysis.py:221-233"""
        d = self.results_data(as_dict=True)
        s = ""
        for key, value in d.items():
            if isinstance(value, dict):
                s += f"{key}:\n"
                for k, v in value.items():
                    if not isinstance(v, list):
                        s += f"{k}: {v}\n"
            else:
                s += f"{key}: {value}\n"
        return s
