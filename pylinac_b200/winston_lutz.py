"""Winston-Lutz per-image (2-D) analysis -- drop-in for ``pylinac.winston_lutz.WinstonLutz2D`` (reference file cited per item).

``WinstonLutz2D(image, **kw).analyze(**kw)`` keeps the reference's signature and accessors; the per-frame pipeline (histogram
inversion check, edge clean-up, ground / normalize, field mask centre of mass, BB search over <= 50 thresholds with labelling,
region properties and the five detection predicates, field / BB matching, CAX->BB and CAX->EPID vectors) runs in CUDA
(pylinac_b200/csrc/wl.cu).  ``analyze_batch(frames, dpmm, ...)`` is the batched entry point (one result per frame).

Scope: BB arrangement ISO, one field and one BB per image.  Out of scope here: the
set-level 3-D solve of ``WinstonLutz`` (a32: host-side scipy minimisation over N small vectors), multi-target arrangements,
``shift_vector`` virtual shifts, plotting / PDF / QuAAC.

The reference delegates labelling and region properties to scikit-image, which is absent from the build container: the CUDA
kernels follow the published algorithms (see oracle/skimage_shim.py); that boundary is unpinned against skimage itself.
"""
from __future__ import annotations

import enum
from collections.abc import Sequence

import numpy as np

from . import _native as nat
from .core import image
from .core.geometry import Point, Vector
from .core.utilities import ResultBase, ResultsDataMixin

BB_ERROR_MESSAGE = (
    "Unable to locate the BB. Make sure the field edges do not obscure the BB, that there are no artifacts in the images, that "
    "the 'bb_size' parameter is close to reality, and that the BB is near the center (within 2cm). If this is a large-field "
    "image or kV image try setting 'low_density_bb' to True."
)  # winston_lutz.py:50-54

_STATUS_ERRORS = {
    nat.WL_NO_BB: BB_ERROR_MESSAGE,
    nat.WL_MISMATCH: "The number of detected fields and BBs do not match",                     # winston_lutz.py:743-746
    nat.WL_NO_FIELD: "No fields were detected",                                                # winston_lutz.py:747-748
    nat.WL_CAPACITY: "The BB search window, field or a candidate region exceeds the capacity of the GPU pipeline",
    nat.WL_FLAT_IMAGE: "The image is flat (max == min)",
}


class Axis(enum.Enum):
    """winston_lutz.py:422-429"""

    GANTRY = "Gantry"
    COLLIMATOR = "Collimator"
    COUCH = "Couch"
    GB_COMBO = "GB Combo"
    GBP_COMBO = "GBP Combo"
    EPID = "Epid"
    REFERENCE = "Reference"


class WinstonLutz2DResult(ResultBase):
    """winston_lutz.py:432-453 (points / vectors serialised as dicts)"""

    variable_axis: str
    bb_location: dict
    cax2epid_vector: dict
    cax2epid_distance: float
    cax2bb_vector: dict
    cax2bb_distance: float
    field_cax: dict


def wrap360(value: float) -> float:
    """core/utilities.py wrap360"""
    return value % 360


def is_close_degrees(angle1: float, angle2: float, delta: float = 1) -> bool:
    """core/utilities.py:170-188"""
    if delta < 0:
        raise ValueError("Delta must be positive")
    a1, a2 = wrap360(angle1), wrap360(angle2)
    simple = abs(a1 - a2)
    return min(simple, 360 - simple) <= delta


def variable_axis(gantry: float, coll: float, couch: float, *, snap_tolerance: float = 3, gantry_reference: float = 0,
                  collimator_reference: float = 0, couch_reference: float = 0) -> Axis:
    """winston_lutz.py:1073-1107"""
    G0 = is_close_degrees(gantry, gantry_reference, delta=snap_tolerance)
    B0 = is_close_degrees(coll, collimator_reference, delta=snap_tolerance)
    P0 = is_close_degrees(couch, couch_reference, delta=snap_tolerance)
    if G0 and B0 and not P0:
        return Axis.COUCH
    if G0 and P0 and not B0:
        return Axis.COLLIMATOR
    if P0 and B0 and not G0:
        return Axis.GANTRY
    if P0 and B0 and G0:
        return Axis.REFERENCE
    if P0:
        return Axis.GB_COMBO
    return Axis.GBP_COMBO


def make_params(dpmm: float, *, bb_size_mm: float = 5, low_density_bb: bool = False, open_field: bool = False,
                bb_proximity_mm: float = 20) -> nat.WlParams:
    """WinstonLutz2D.analyze() arguments (winston_lutz.py:1152-1164) -> the C-ABI struct."""
    if not dpmm > 0:
        raise ValueError("dpmm must be positive")
    if not bb_size_mm > 0:
        raise ValueError("bb_size_mm must be positive")
    p = nat.WlParams()
    p.dpmm = float(dpmm)
    p.bb_size_mm = float(bb_size_mm)
    p.low_density_bb = 1 if low_density_bb else 0
    p.open_field = 1 if open_field else 0
    p.bb_proximity_mm = float(bb_proximity_mm)
    return p


class WLFrameResult:
    """One frame's results (a row of the struct-of-arrays the GPU returns)."""

    def __init__(self, row):
        self.r = row

    @property
    def status(self) -> int:
        return int(self.r["status"])

    def raise_for_status(self):
        if self.status != nat.WL_OK:
            raise ValueError(_STATUS_ERRORS.get(self.status, f"Winston-Lutz status {self.status}"))

    @property
    def shape(self) -> tuple[int, int]:
        return int(self.r["height"]), int(self.r["width"])

    @property
    def bb(self) -> Point:
        return Point(float(self.r["bb_x"]), float(self.r["bb_y"]))

    @property
    def field_cax(self) -> Point:
        return Point(float(self.r["field_x"]), float(self.r["field_y"]))

    @property
    def epid(self) -> Point:
        return Point(float(self.r["epid_x"]), float(self.r["epid_y"]))

    @property
    def cax2bb_vector(self) -> Vector:
        return Vector(float(self.r["cax2bb_x"]), float(self.r["cax2bb_y"]), 0.0)

    @property
    def cax2bb_distance(self) -> float:
        return float(self.r["cax2bb_distance"])

    @property
    def cax2epid_vector(self) -> Vector:
        return Vector(float(self.r["cax2epid_x"]), float(self.r["cax2epid_y"]), 0.0)

    @property
    def cax2epid_distance(self) -> float:
        return float(self.r["cax2epid_distance"])


class WLBatchResult(Sequence):
    def __init__(self, rows: np.ndarray):
        self.rows = rows

    def __len__(self):
        return len(self.rows)

    def __getitem__(self, i) -> WLFrameResult:
        return WLFrameResult(self.rows[i])


def analyze_batch(frames, dpmm: float, *, device: int | None = None, bb_size_mm: float = 5, low_density_bb: bool = False,
                  open_field: bool = False, bb_proximity_mm: float = 20) -> WLBatchResult:
    """WinstonLutz2D(...).analyze(**kw) for every frame of ``frames`` (uint16 [n,h,w] ndarray or a device-resident Batch)."""
    ctx = nat.Context.default(device)
    params = make_params(dpmm, bb_size_mm=bb_size_mm, low_density_bb=low_density_bb, open_field=open_field,
                         bb_proximity_mm=bb_proximity_mm)
    return WLBatchResult(nat.wl2d_analyze(ctx, frames, params))


class WinstonLutz2D(ResultsDataMixin[WinstonLutz2DResult]):
    """winston_lutz.py:629-1231 -- same constructor keywords / analyze() signature for the single-image case."""

    def __init__(self, file, use_filenames: bool = False, **kwargs):
        if use_filenames:
            raise NotImplementedError("axis values from file names are an ingest feature outside the accelerated hot path")
        if isinstance(file, np.ndarray):
            self.image = image.ArrayImage(file, **{k: v for k, v in kwargs.items() if k in ("dpi", "sid", "dtype")})
        elif isinstance(file, image.BaseImage):
            self.image = file
        else:
            self.image = image.load(file, **kwargs)
        self.gantry_angle = float(kwargs.get("gantry", getattr(self.image, "gantry_angle", 0.0) or 0.0))
        self.collimator_angle = float(kwargs.get("coll", getattr(self.image, "collimator_angle", 0.0) or 0.0))
        self.couch_angle = float(kwargs.get("couch", getattr(self.image, "couch_angle", 0.0) or 0.0))
        if self.image.dpmm is None:
            raise ValueError("DPI was not a tag in the image nor was it passed in. Please pass a DPI value")
        self._is_analyzed = False
        self._result: WLFrameResult | None = None
        self._snap_tolerance = 3.0
        self._gantry_reference = self._collimator_reference = self._couch_reference = 0.0

    def _frame_u16(self) -> np.ndarray:
        a = np.asarray(self.image.array)
        if a.dtype == np.uint16:
            return a
        if a.dtype == np.uint8:
            return a.astype(np.uint16)
        if a.dtype.kind in "fiu" and a.min() >= 0 and a.max() <= 65535 and np.array_equal(a, np.floor(a)):
            return a.astype(np.uint16)
        raise NotImplementedError("the GPU Winston-Lutz path takes integer-valued pixel data in [0, 65535]")

    @property
    def dpmm(self) -> float:
        return self.image.dpmm

    def analyze(self, bb_size_mm: float = 5, low_density_bb: bool = False, open_field: bool = False, shift_vector=None,
                snap_tolerance: float = 3, gantry_reference: float = 0, collimator_reference: float = 0,
                couch_reference: float = 0, bb_proximity_mm: float = 20, machine_scale=None) -> None:
        """winston_lutz.py:1152-1183"""
        if shift_vector is not None:
            raise NotImplementedError("virtual BB shifts (shift_vector) are outside the accelerated per-image path")
        self._snap_tolerance = snap_tolerance
        self._gantry_reference = gantry_reference
        self._collimator_reference = collimator_reference
        self._couch_reference = couch_reference
        res = analyze_batch(self._frame_u16(), self.dpmm, bb_size_mm=bb_size_mm, low_density_bb=low_density_bb,
                            open_field=open_field, bb_proximity_mm=bb_proximity_mm)[0]
        res.raise_for_status()
        self._result = res
        self._is_analyzed = True
        self.bb = res.bb
        self.field_cax = res.field_cax
        self.shape = res.shape

    def __repr__(self):
        return f"WLImage(gantry={self.gantry_angle:.1f}, coll={self.collimator_angle:.1f}, couch={self.couch_angle:.1f})"

    def _need(self) -> WLFrameResult:
        if self._result is None:
            raise ValueError("The image is not analyzed. Use .analyze() first.")
        return self._result

    @property
    def epid(self) -> Point:
        return self._need().epid

    @property
    def cax2bb_vector(self) -> Vector:
        return self._need().cax2bb_vector

    @property
    def cax2bb_distance(self) -> float:
        return self._need().cax2bb_distance

    @property
    def cax2epid_vector(self) -> Vector:
        return self._need().cax2epid_vector

    @property
    def cax2epid_distance(self) -> float:
        return self._need().cax2epid_distance

    @property
    def variable_axis(self) -> Axis:
        return variable_axis(self.gantry_angle, self.collimator_angle, self.couch_angle, snap_tolerance=self._snap_tolerance,
                             gantry_reference=self._gantry_reference, collimator_reference=self._collimator_reference,
                             couch_reference=self._couch_reference)

    def _generate_results_data(self) -> WinstonLutz2DResult:
        """winston_lutz.py:1215-1231"""
        r = self._need()

        def ser(p):
            return {"x": p.x, "y": p.y, "z": p.z}

        return WinstonLutz2DResult(variable_axis=self.variable_axis.value, cax2bb_vector=ser(r.cax2bb_vector),
                                   cax2epid_vector=ser(r.cax2epid_vector), cax2bb_distance=r.cax2bb_distance,
                                   cax2epid_distance=r.cax2epid_distance, bb_location=ser(r.bb), field_cax=ser(r.field_cax))
