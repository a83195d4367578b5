"""Starshot analysis -- drop-in for the hot path of ``pylinac.starshot`` (reference file cited per item).

``Starshot(image, **kw).analyze(**kw)`` keeps the reference's signature and result accessors; underneath, the whole
per-frame pipeline (histogram inversion check, ground, FW80M start point, collapsed circle profile, gaussian filter, FWXM
peaks, line matching, Nelder-Mead wobble circle, recursive parameter search) runs in CUDA (pylinac_b200/csrc/starshot.cu).
``analyze_batch(frames, dpmm, ...)`` is the batched entry point (one result per frame).

Out of scope: plotting, PDF / QuAAC export.
"""
from __future__ import annotations

from collections.abc import Sequence

import numpy as np

from . import _native as nat
from .core import image
from .core.geometry import Line, Point
from .core.warnings import capture_warnings
from .core.utilities import ResultBase, ResultsDataMixin

_STATUS_ERRORS = {
    nat_code: msg for nat_code, msg in {
        1: "The algorithm was unable to determine a reasonable wobble. Try setting recursive to False and manually adjusting "
           "algorithm parameters",
        2: "The algorithm was unable to properly detect the radiation lines. Try setting recursive to True or lower the minimum "
           "peak height",
        3: "No full-width-80%-maximum peak was found in the central third of the image",
        4: "The circle profile or its peak count exceeds the capacity of the GPU pipeline",
        5: "The image is flat (max == min)",
    }.items()
}


class StarshotResults(ResultBase):
    """starshot.py:48-74"""

    tolerance_mm: float
    circle_diameter_mm: float
    circle_radius_mm: float
    circle_center_x_y: tuple[float, float]
    angles: list[float]
    passed: bool


class Wobble:
    """starshot.py:683-698 (Circle with radius_mm)."""

    def __init__(self, center_point=None, radius=None):
        self.center = Point(center_point) if center_point is not None else Point()
        self.radius = radius
        self.radius_mm = 0

    @property
    def diameter(self):
        return self.radius * 2

    @property
    def diameter_mm(self) -> float:
        return self.radius_mm * 2


def make_params(dpmm: float, *, radius: float = 0.85, min_peak_height: float = 0.25, max_wobble_diameter: float = 2.0,
                tolerance: float = 1.0, start_point=None, fwhm: bool = True, recursive: bool = True,
                invert: bool = False) -> nat.StarParams:
    """analyze() arguments (starshot.py:230-240) -> the C-ABI struct; bounds as in @argue.bounds (starshot.py:229)."""
    if not 0.2 <= radius <= 0.95:
        raise ValueError("radius must be between 0.2 and 0.95")
    if not 0.05 <= min_peak_height <= 0.95:
        raise ValueError("min_peak_height must be between 0.05 and 0.95")
    p = nat.StarParams()
    p.dpmm = float(dpmm)
    p.radius = float(radius)
    p.min_peak_height = float(min_peak_height)
    p.max_wobble_diameter = float(max_wobble_diameter)
    p.tolerance = float(tolerance)
    if start_point is not None:
        sp = Point(start_point)
        p.has_start_point, p.start_x, p.start_y = 1, float(sp.x), float(sp.y)
    p.fwhm = 1 if fwhm else 0
    p.recursive = 1 if recursive else 0
    p.invert = 1 if invert else 0
    return p


class StarFrameResult:
    """One frame's results (a row of the struct-of-arrays the GPU returns)."""

    def __init__(self, row, tolerance: float):
        self.r = row
        self.tolerance = tolerance

    @property
    def status(self) -> int:
        return int(self.r["status"])

    def raise_for_status(self):
        if self.status != 0:
            raise RuntimeError(_STATUS_ERRORS.get(self.status, f"starshot status {self.status}"))

    @property
    def wobble(self) -> Wobble:
        w = Wobble(Point(float(self.r["wobble_x"]), float(self.r["wobble_y"])), float(self.r["wobble_radius_px"]))
        w.radius_mm = float(self.r["wobble_radius_mm"])
        return w

    @property
    def peaks(self) -> list[Point]:
        n = int(self.r["n_peaks"])
        return [Point(float(self.r["peak_x"][k]), float(self.r["peak_y"][k]), idx=int(self.r["peak_idx"][k])) for k in range(n)]

    @property
    def lines(self) -> list[Line]:
        pk = self.peaks
        n = int(self.r["n_lines"])
        return [Line(pk[k], pk[k + n]) for k in range(n)]

    @property
    def angles(self) -> list[float]:
        return [float(v) for v in self.r["angles"][: int(self.r["n_lines"])]]

    @property
    def passed(self) -> bool:
        return bool(self.r["passed"])

    def results_data(self) -> StarshotResults:
        w = self.wobble
        return StarshotResults(tolerance_mm=self.tolerance, circle_diameter_mm=w.radius_mm * 2, circle_radius_mm=w.radius_mm,
                               circle_center_x_y=(w.center.x, w.center.y), angles=self.angles, passed=self.passed)


class StarBatchResult(Sequence):
    def __init__(self, rows: np.ndarray, tolerance: float):
        self.rows = rows
        self.tolerance = tolerance

    def __len__(self):
        return len(self.rows)

    def __getitem__(self, i) -> StarFrameResult:
        return StarFrameResult(self.rows[i], self.tolerance)


def analyze_batch(frames, dpmm: float, *, device: int | None = None, radius: float = 0.85, min_peak_height: float = 0.25,
                  max_wobble_diameter: float = 2.0, tolerance: float = 1.0, start_point=None, fwhm: bool = True,
                  recursive: bool = True, invert: bool = False) -> StarBatchResult:
    """Starshot(...).analyze(**kw) for every frame of ``frames`` (uint16 [n,h,w] ndarray or a device-resident Batch)."""
    ctx = nat.Context.default(device)
    params = make_params(dpmm, radius=radius, min_peak_height=min_peak_height, max_wobble_diameter=max_wobble_diameter,
                         tolerance=tolerance, start_point=start_point, fwhm=fwhm, recursive=recursive, invert=invert)
    rows = nat.starshot_analyze(ctx, frames, params)
    return StarBatchResult(rows, tolerance)


@capture_warnings
class Starshot(ResultsDataMixin[StarshotResults]):
    """starshot.py:77-125, 230-304, 403-447 -- same constructor / analyze() signature."""

    def __init__(self, filepath, **kwargs):
        if isinstance(filepath, np.ndarray):
            self.image = image.ArrayImage(filepath, **kwargs)
        elif isinstance(filepath, image.BaseImage):
            self.image = filepath
        else:
            self.image = image.load(filepath, **kwargs)
        self.wobble = Wobble()
        self.tolerance = 1
        if self.image.dpmm is None:
            raise ValueError("DPI was not a tag in the image nor was it passed in. Please pass a DPI value")
        if self.image.sid is None:
            raise ValueError("Source-to-Image distance was not an image tag and was not passed in. Please pass an SID value.")
        self._result: StarFrameResult | None = None

    @classmethod
    def from_zip(cls, zip_file, **kwargs):
        """starshot.py:176-195: one image, or several images of one test sequence that are superimposed."""
        with image.TemporaryZipDirectory(zip_file) as tmp:
            files = image.retrieve_image_files(tmp)
            if not files:
                raise IndexError(f"No valid starshot images were found in {zip_file}")
            return cls.from_multiple_images(files, **kwargs) if len(files) > 1 else cls(files[0], **kwargs)

    @classmethod
    def from_multiple_images(cls, filepath_list, stretch_each: bool = True, method: str = "sum", **kwargs):
        """starshot.py:148-174: superimpose the images of the individual spokes (``load_multiples``), then construct from the
        composite as the reference does after its in-memory DICOM write / read (``image._resaved``: full-range re-quantisation
        to the stored dtype, then the file's own rescale tags)."""
        combined = image.load_multiples(filepath_list, stretch_each=stretch_each, method=method, **kwargs)
        return cls(image._resaved(combined), **kwargs)

    def _frame_u16(self) -> np.ndarray:
        return image.frame_u16(self.image, "GPU starshot")

    def analyze(self, radius: float = 0.85, min_peak_height: float = 0.25, max_wobble_diameter: float = 2.0,
                tolerance: float = 1.0, start_point=None, fwhm: bool = True, recursive: bool = True, invert: bool = False):
        """starshot.py:230-304"""
        self.tolerance = tolerance
        res = analyze_batch(self._frame_u16(), self.image.dpmm, radius=radius, min_peak_height=min_peak_height,
                            max_wobble_diameter=max_wobble_diameter, tolerance=tolerance, start_point=start_point, fwhm=fwhm,
                            recursive=recursive, invert=invert)[0]
        res.raise_for_status()
        self._result = res
        self.wobble = res.wobble
        self.lines = res.lines
        self.angles = res.angles

    @property
    def passed(self) -> bool:
        return bool(self.wobble.radius_mm * 2 < self.tolerance)

    @property
    def _passfail_str(self) -> str:
        return "PASS" if self.passed else "FAIL"

    def results(self, as_list: bool = False):
        """starshot.py:413-431"""
        results = [
            " - Starshot Results - ",
            f"Result: {self._passfail_str}",
            f"The minimum circle that touches all the star lines has a diameter of {self.wobble.radius_mm * 2:2.3f} mm.",
            f"The center of the minimum circle is at {self.wobble.center.x:3.1f}, {self.wobble.center.y:3.1f}",
        ]
        if not as_list:
            results = "\n".join(results)
        return results

    def _generate_results_data(self) -> StarshotResults:
        return StarshotResults(tolerance_mm=self.tolerance, circle_diameter_mm=self.wobble.radius_mm * 2,
                               circle_radius_mm=self.wobble.radius_mm,
                               circle_center_x_y=(self.wobble.center.x, self.wobble.center.y), angles=self.angles, passed=self.passed)
