"""Dosimetric leaf gap from an EPID image with varying MLC overlaps -- drop-in for ``pylinac.dlg.DLG`` (dlg.py:15-127).

``DLG(path).analyze(gaps, mlc, y_field_size=100, profile_width=10)`` keeps the reference's signature and attributes
(``measured_dlg``, ``measured_dlg_per_leaf``, ``planned_dlg_per_leaf``).  The per-leaf window means, the inversion rule and peak
prominence of ``_determine_measured_gap`` and the line fit are two CUDA launches (``epid_dlg_analyze``, csrc/vmat.cu);
``analyze_batch(frames, dpmm, ...)`` runs n frames at once.  Not here: ``plot_dlg``.
"""
from __future__ import annotations

from collections.abc import Sequence
from math import ceil, floor
from types import SimpleNamespace

import numpy as np

from . import _native as nat
from .core import image
from .core.image import frame_u16
from .picketfence import MLC


def _get_dlg_offset(field_size: float, leaf_center: float, dlgs: Sequence) -> float | None:
    """dlg.py:101-110: the planned overlap of the gap band a leaf centre falls in (None exactly on a band edge, like the reference)"""
    roi_size = field_size / len(dlgs)
    y_bounds = [field_size / 2 - idx * roi_size for idx in range(len(dlgs) + 1)]
    for idx, gap in enumerate(dlgs):
        if y_bounds[idx + 1] < leaf_center < y_bounds[idx]:
            return gap
    return None


def _windows(shape, dpmm: float, gaps: Sequence, mlc, y_field_size: float, profile_width: int):
    """dlg.py:53-82: per analysed leaf the row window [bottom, top), the shared column window and the planned gap"""
    arrangement = mlc.value["arrangement"] if isinstance(mlc, MLC) else mlc
    g = sorted(gaps)
    profile_width_px = round(dpmm * profile_width)
    mid_width, mid_height = shape[1] / 2, shape[0] / 2
    bottoms, tops, planned = [], [], []
    for idx, center in enumerate(arrangement.centers):
        if -y_field_size / 2 < center < y_field_size / 2:
            center_px = center * dpmm
            width_px = arrangement.widths[idx] / 4 * dpmm
            tops.append(ceil(mid_height + center_px + width_px))
            bottoms.append(floor(mid_height + center_px - width_px))
            planned.append(_get_dlg_offset(y_field_size, center, g))
    if any(p is None for p in planned):
        # np.asarray([.., None]) makes linregress fail in the reference as well
        raise TypeError("a leaf centre lies exactly on a gap-band edge: the reference's _get_dlg_offset returns None for it")
    c0, c1 = int(mid_width - profile_width_px), int(mid_width + profile_width_px)
    return bottoms, tops, c0, c1, planned


def analyze_batch(frames, dpmm: float, gaps: Sequence, mlc, y_field_size: float = 100, profile_width: int = 10, device: int | None = None):
    """n uint16 frames [n, H, W] -> dict(measured_dlg [n], measured_dlg_per_leaf [n, leaves], planned_dlg_per_leaf, slope, intercept)"""
    a = np.asarray(frames)
    shape = a.shape[-2:]
    bottoms, tops, c0, c1, planned = _windows(shape, dpmm, gaps, mlc, y_field_size, profile_width)
    meas, slope, icpt, dlg = nat.dlg_analyze(nat.Context.default(device), frames, bottoms, tops, c0, c1, planned)
    return {"measured_dlg": dlg, "measured_dlg_per_leaf": meas, "planned_dlg_per_leaf": list(planned), "slope": slope, "intercept": icpt}


class DLG:
    """dlg.py:15-127"""

    def __init__(self, path):
        self.image = image.LinacDicomImage(path) if not isinstance(path, image.BaseImage) else path
        self.measured_dlg: float = -np.inf
        self.measured_dlg_per_leaf: list = []
        self.planned_dlg_per_leaf: list = []
        self._lin_fit = None

    def analyze(self, gaps: Sequence, mlc: MLC, y_field_size: float = 100, profile_width: int = 10):
        frame = np.ascontiguousarray(frame_u16(self.image, "DLG"))
        out = analyze_batch(frame[None], self.image.dpmm, gaps, mlc, y_field_size, profile_width)
        self._lin_fit = SimpleNamespace(slope=float(out["slope"][0]), intercept=float(out["intercept"][0]))
        self.measured_dlg = float(out["measured_dlg"][0])
        self.planned_dlg_per_leaf = list(out["planned_dlg_per_leaf"])
        self.measured_dlg_per_leaf = [float(v) for v in out["measured_dlg_per_leaf"][0]]
