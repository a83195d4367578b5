"""Multi-target, multi-field Winston-Lutz -- drop-in for the analysis path of ``WinstonLutzMultiTargetMultiField`` /
``WinstonLutzMultiTargetMultiFieldImage`` (winston_lutz.py:91-362, 624-845, 2728-3230, 3401-3650).

Per image, on the device: the histogram inversion check + edge clean-up of the Winston-Lutz front end (``epid_wl2d_analyze`` reports
the decisions), the radiation fields by the whole-frame field locator (``epid_global_locate``: 8-connectivity labelling at every
threshold of the sweep, csrc/locate.cu) and one windowed disk search per BB of the arrangement (``epid_disk_locate``, csrc/wl.cu).
Matching detected points to the arrangement, the 3-D positions (Low et al.) and the 6-degree-of-freedom alignment are scalar work
on a handful of points, written out with numpy like the reference does (no scipy).  Not here: plotting, PDF / QuAAC export.
"""
from __future__ import annotations

import math
from collections.abc import Sequence
from dataclasses import dataclass

import numpy as np
from pydantic import BaseModel, Field

from . import _native as nat
from .core import image
from .core.geometry import Point, Vector
from .core.utilities import ResultBase, ResultsDataMixin
from .core.warnings import capture_warnings
from .metrics.features import is_modest_size, is_right_square_size, is_round, is_square, is_symmetric
from .metrics.image import GlobalSizedFieldLocator, SizedDiskLocator
from .winston_lutz import (BB_ERROR_MESSAGE, Axis, MachineScale, _virtual_shift_px, bb_projection_with_rotation,
                           solve_3d_position_from_2d_planes, straight_ray, variable_axis)


class BBConfig(BaseModel):
    """winston_lutz.py:91-104"""

    name: str
    offset_left_mm: float
    offset_up_mm: float
    offset_in_mm: float
    bb_size_mm: float
    rad_size_mm: float

    def to_human(self) -> str:
        lr = "Left" if self.offset_left_mm >= 0 else "Right"
        ud = "Up" if self.offset_up_mm >= 0 else "Down"
        io = "In" if self.offset_in_mm >= 0 else "Out"
        return f"{lr} {abs(self.offset_left_mm)}mm, {ud} {abs(self.offset_up_mm)}mm, {io} {abs(self.offset_in_mm)}mm"


def _cfgs(rows):
    return tuple(BBConfig(name=n, offset_left_mm=l, offset_up_mm=u, offset_in_mm=i, bb_size_mm=5, rad_size_mm=20) for n, l, u, i in rows)


class BBArrangement:
    """winston_lutz.py:107-250: presets for multi-target phantoms"""

    ISO = _cfgs([("Iso", 0, 0, 0)])
    SNC_MULTIMET = _cfgs([("Iso", 0, 0, 0), ("1", 0, 0, 30), ("2", -30, 0, 15), ("3", 0, 0, -30), ("4", 30, 0, -50), ("5", 0, 0, -70)])
    DEMO = SNC_MULTIMET


@dataclass
class BBFieldMatch:
    """winston_lutz.py:252-310: a BB and a field matched to one arrangement position"""

    epid: Point
    field: Point
    bb: Point
    dpmm: float
    gantry_angle: float
    couch_angle: float
    sad: float

    def _coord(self, a: Point, b: Point) -> Vector:
        v = (a - b) / self.dpmm
        return Vector(v.x, -v.y, v.z)      # image y grows downwards, coordinate-space y upwards

    @property
    def field_epid_vector_mm(self) -> Vector:
        return self._coord(self.field, self.epid)

    @property
    def bb_field_vector_mm(self) -> Vector:
        return self._coord(self.bb, self.field)

    @property
    def bb_epid_vector_mm(self) -> Vector:
        return self._coord(self.bb, self.epid)

    @property
    def bb_field_distance_mm(self) -> float:
        return self.field.distance_to(self.bb) / self.dpmm

    @property
    def bb_epid_distance_mm(self) -> float:
        return self.epid.distance_to(self.bb) / self.dpmm

    @property
    def field_epid_distance_mm(self) -> float:
        return self.epid.distance_to(self.field) / self.dpmm

    @property
    def bb_to_field_projection(self):
        return straight_ray(self.bb_field_vector_mm, self.gantry_angle)


class BB3D:
    """winston_lutz.py:313-362"""

    def __init__(self, bb_config: BBConfig, bb_matches: Sequence[BBFieldMatch], scale: MachineScale):
        self.bb_config = bb_config
        self.matches = bb_matches
        self.scale = scale

    def _solve(self, which: str) -> Point:
        vs = [getattr(m, which) for m in self.matches]
        v = solve_3d_position_from_2d_planes([t.x for t in vs], [t.y for t in vs], [m.gantry_angle for m in self.matches],
                                             [m.couch_angle for m in self.matches], self.scale)
        return Point(v.x, v.y, v.z)

    @property
    def measured_bb_position(self) -> Point:
        return self._solve("bb_epid_vector_mm")

    @property
    def measured_field_position(self) -> Point:
        return self._solve("field_epid_vector_mm")

    @property
    def nominal_bb_position(self) -> Point:
        return Point(-self.bb_config.offset_left_mm, self.bb_config.offset_in_mm, self.bb_config.offset_up_mm)


def _rot(axis: str, a: float) -> np.ndarray:
    c, s = math.cos(a), math.sin(a)
    return {"x": np.array([[1, 0, 0], [0, c, -s], [0, s, c]]), "y": np.array([[c, 0, s], [0, 1, 0], [-s, 0, c]]),
            "z": np.array([[c, -s, 0], [s, c, 0], [0, 0, 1]])}[axis]


def align_points(measured_points: Sequence[Point], ideal_points: Sequence[Point], axes_order: str = "roll,pitch,yaw"):
    """winston_lutz.py:3608-3660: rigid alignment (Kabsch) of the measured onto the ideal points -> (translation, yaw, pitch, roll).
    The reference reads the angles with scipy's Rotation.as_euler (extrinsic, roll = y, pitch = x, yaw = z); for the default order
    R = Rz(yaw) Rx(pitch) Ry(roll), whose angles are read off the matrix directly."""
    order = [a.strip() for a in axes_order.split(",")]
    euler = "".join({"pitch": "x", "yaw": "z", "roll": "y"}[a] for a in order)        # conventional_to_euler_notation (:3592-3605)
    measured = np.array([[p.x, p.y, p.z] for p in measured_points], dtype=float)
    ideal = np.array([[p.x, p.y, p.z] for p in ideal_points], dtype=float)
    mc, ic = np.mean(measured, axis=0), np.mean(ideal, axis=0)
    H = (measured - mc).T @ (ideal - ic)
    U, _, Vt = np.linalg.svd(H)
    R = Vt.T @ U.T
    if np.linalg.det(R) < 0:
        Vt[2, :] *= -1
        R = Vt.T @ U.T
    if order == ["roll", "pitch", "yaw"]:
        # R = Rz(c) Rx(b) Ry(a):  R[2] = (-cos b sin a, sin b, cos b cos a),  R[0][1] = -sin c cos b,  R[1][1] = cos c cos b
        pitch = math.degrees(math.asin(max(-1.0, min(1.0, R[2, 1]))))
        roll = math.degrees(math.atan2(-R[2, 0], R[2, 2]))
        yaw = math.degrees(math.atan2(-R[0, 1], R[1, 1]))
    else:
        # any other order: exactly the reference's call, including its positional unpacking of the three angles (:3656-3658);
        # scalar host work on a 3 x 3 matrix, like the set-level scipy.optimize call of WinstonLutz
        from scipy.spatial.transform import Rotation

        roll, pitch, yaw = (float(v) for v in Rotation.from_matrix(R).as_euler(euler, degrees=True))
    translation = ic - R @ mc
    return Vector(*translation), yaw, pitch, roll


class WinstonLutzMultiTargetMultiFieldResult(ResultBase):
    """winston_lutz.py:544-587"""

    num_total_images: int = Field(description="The total number of images analyzed.")
    max_2d_field_to_bb_mm: float = Field(description="The maximum 2D distance from any BB to its field center.")
    mean_2d_field_to_bb_mm: float = Field(description="The mean 2D distance from any BB to its field center.")
    median_2d_field_to_bb_mm: float = Field(description="The median 2D distance from any BB to its field center.")
    bb_arrangement: tuple[BBConfig, ...] = Field(description="A list of expected arrangements of the BBs")
    bb_maxes: dict[str, float] = Field(description="The maximum 2D distances of each BB to its field center, keyed by BB name.")
    bb_shift_vector: dict = Field(description="The vector (in 3D cartesian space) to move the phantom to align with the isocenter in mm.")
    bb_shift_yaw: float = Field(description="The yaw rotation in degrees needed to align the phantom with the radiation isocenter.")
    bb_shift_pitch: float = Field(description="The pitch rotation needed in degrees to align the phantom with the radiation isocenter.")
    bb_shift_roll: float = Field(description="The roll rotation needed in degrees to align the phantom with the radiation isocenter.")


class WinstonLutzMultiTargetMultiFieldImage:
    """winston_lutz.py:624-845, 2728-2801: one image with N fields and M BBs."""

    detection_conditions = [is_round, is_symmetric, is_modest_size]
    field_conditions = [is_square, is_right_square_size]      # kept for API parity; the reference's field search uses the locator defaults

    def __init__(self, file, *, gantry: float | None = None, coll: float | None = None, couch: float | None = None, sad: float | None = None,
                 detection_conditions=None, **kwargs):
        if detection_conditions:
            self.detection_conditions = detection_conditions
        if isinstance(file, np.ndarray):
            self.image = image.ArrayImage(file, **{k: v for k, v in kwargs.items() if k in ("dpi", "sid", "dtype")})
        elif isinstance(file, image.BaseImage):
            self.image = file
        else:
            self.image = image.LinacDicomImage(file, **kwargs)
        pick = lambda v, attr: float(v if v is not None else (getattr(self.image, attr, 0.0) or 0.0))      # noqa: E731
        self.gantry_angle, self.collimator_angle, self.couch_angle = pick(gantry, "gantry_angle"), pick(coll, "collimator_angle"), pick(couch, "couch_angle")
        self.sad = float(sad if sad is not None else (getattr(self.image, "sad", 1000.0) or 1000.0))
        if self.image.dpmm is None:
            raise ValueError("DPI was not a tag in the image nor was it passed in. Please pass a DPI value")
        self.dpmm = self.image.dpmm
        self._is_analyzed = False
        self.arrangement_matches: dict[str, BBFieldMatch] = {}

    def __repr__(self):
        return f"WLMTMFImage(gantry={self.gantry_angle:.1f}, coll={self.collimator_angle:.1f}, couch={self.couch_angle:.1f})"

    @property
    def variable_axis(self) -> Axis:
        return variable_axis(self.gantry_angle, self.collimator_angle, self.couch_angle, snap_tolerance=self._snap_tolerance,
                             gantry_reference=self._gantry_reference, collimator_reference=self._collimator_reference,
                             couch_reference=self._couch_reference)

    # -- geometry helpers (winston_lutz.py:1041-1060)
    def nominal_bb_position(self, bb_config: BBConfig) -> Point:
        sx, sy = bb_projection_with_rotation(offset_left=bb_config.offset_left_mm, offset_up=bb_config.offset_up_mm,
                                             offset_in=bb_config.offset_in_mm, sad=self.sad, gantry=self.gantry_angle,
                                             couch=self.couch_angle, machine_scale=self.machine_scale)
        return Point(self.epid.x + sx * self.dpmm, self.epid.y - sy * self.dpmm)

    def _find_matches(self, detected_points: list[Point], bb_proximity_mm: float) -> dict[str, Point]:
        """winston_lutz.py:826-845 (ValueError from min() of an empty list when nothing was detected, like the reference)"""
        out = {}
        for cfg in self.bb_arrangement:
            nominal = self.nominal_bb_position(cfg)
            distances = [nominal.distance_to(p) for p in detected_points]
            dmin = min(distances)
            if dmin < bb_proximity_mm * self.dpmm:
                out[cfg.name] = detected_points[distances.index(dmin)]
        return out

    def analyze(self, bb_arrangement: Sequence[BBConfig], is_open_field: bool = False, is_low_density: bool = False, shift_vector=None,
                snap_tolerance: float = 3, gantry_reference: float = 0, collimator_reference: float = 0, couch_reference: float = 0,
                bb_proximity_mm: float = 20, machine_scale: MachineScale = MachineScale.IEC61217):
        """winston_lutz.py:660-762"""
        if snap_tolerance < 0:
            raise ValueError("Snap tolerance must be >= 0")
        self._snap_tolerance, self._gantry_reference = snap_tolerance, gantry_reference
        self._collimator_reference, self._couch_reference = collimator_reference, couch_reference
        self.machine_scale = machine_scale
        self.bb_arrangement = tuple(bb_arrangement)
        ctx = nat.Context.default()
        frame = np.ascontiguousarray(image.frame_u16(self.image, "GPU Winston-Lutz"))
        # check_inversion_by_histogram((0.01, 50, 99.99)) + _clean_edges() on the device: the front end of the single-BB pipeline
        # reports both decisions (its own BB / field search is not used here)
        from .winston_lutz import make_params

        front = nat.wl2d_analyze(ctx, frame, make_params(self.dpmm, bb_size_mm=self.bb_arrangement[0].bb_size_mm, low_density_bb=is_low_density,
                                                        open_field=True, bb_proximity_mm=1e9))[0]
        if int(front["status"]) == nat.WL_FLAT_IMAGE:
            raise ValueError("The image is flat (max == min)")
        crop, inverted = int(front["crop_px"]), bool(front["inverted"])
        work = image.ArrayImage(frame, dpi=self.dpmm * 25.4)      # sid-free copy at the same dots-per-mm
        if inverted:
            work.invert()
        if crop:
            work.crop(pixels=crop)
        work._locator_sample_kind = 1      # the reference's image is ground()-ed and normalize()-d at this point (float g / D)
        self._work = work
        self.shape = work.shape
        self.epid = Point(work.shape[1] / 2 - 0.5, work.shape[0] / 2 - 0.5)      # image.cax of the cropped frame (core/image.py:526-533)
        # -- fields (winston_lutz.py:2734-2763)
        if is_open_field:
            field_caxs = [self.epid]
        else:
            sizes = [c.rad_size_mm for c in self.bb_arrangement]
            mean_size = (max(sizes) + min(sizes)) / 2
            tol = max((max(sizes) - min(sizes)) * 1.2, 0.1 * mean_size)
            field_caxs = work.compute(metrics=GlobalSizedFieldLocator.from_physical(max_number=len(self.bb_arrangement), field_height_mm=mean_size,
                                                                                    field_width_mm=mean_size, field_tolerance_mm=tol))
        field_matches = self._find_matches(field_caxs, bb_proximity_mm)
        # -- BBs (winston_lutz.py:2765-2801): one windowed search per arrangement entry
        detected = []
        for cfg in self.bb_arrangement:
            d = cfg.bb_size_mm
            tol_mm = float(np.interp(d, (1.5, 30), (2, 4)))      # _calculate_bb_tolerance (winston_lutz.py:1062-1067)
            left, sup = bb_projection_with_rotation(offset_left=cfg.offset_left_mm, offset_up=cfg.offset_up_mm, offset_in=cfg.offset_in_mm,
                                                    gantry=self.gantry_angle, couch=self.couch_angle, sad=self.sad)
            try:
                detected.extend(work.compute(metrics=SizedDiskLocator.from_center_physical(
                    expected_position_mm=Point(x=left, y=-sup), search_window_mm=(40 + d, 40 + d), radius_mm=d / 2,
                    radius_tolerance_mm=tol_mm / 2, invert=not is_low_density, detection_conditions=self.detection_conditions)))
            except ValueError:
                pass
        if shift_vector:
            dx, dy = _virtual_shift_px(shift_vector, self.dpmm, self.gantry_angle, self.couch_angle, self.sad, machine_scale)
            for p in detected:
                p.x += dx
                p.y += dy
        bb_matches = self._find_matches(detected, bb_proximity_mm)
        if len(bb_matches) != len(field_matches):
            raise ValueError("The number of detected fields and BBs do not match")
        if not field_matches:
            raise ValueError("No fields were detected")
        if not bb_matches:
            raise ValueError(BB_ERROR_MESSAGE)
        self.field_caxs, self.bb_positions = field_caxs, detected
        self.arrangement_matches = {name: BBFieldMatch(epid=self.epid, field=field_matches[name], bb=bb, dpmm=self.dpmm,
                                                       gantry_angle=self.gantry_angle, couch_angle=self.couch_angle, sad=self.sad)
                                    for name, bb in bb_matches.items()}
        self._is_analyzed = True


@capture_warnings
class WinstonLutzMultiTargetMultiField(ResultsDataMixin[WinstonLutzMultiTargetMultiFieldResult]):
    """winston_lutz.py:2804-3230"""

    image_type = WinstonLutzMultiTargetMultiFieldImage

    def __init__(self, images: Sequence[WinstonLutzMultiTargetMultiFieldImage]):
        super().__init__()
        self.images = list(images)
        self._is_analyzed = False

    @classmethod
    def from_arrays(cls, frames: np.ndarray, axes, *, dpmm: float, sad: float = 1000.0):
        """n frames [n, H, W] with one (gantry, collimator, couch) triple each"""
        return cls([WinstonLutzMultiTargetMultiFieldImage(f, gantry=g, coll=c, couch=p, sad=sad, dpi=dpmm * 25.4) for f, (g, c, p) in zip(frames, axes)])

    def analyze(self, bb_arrangement: Sequence[BBConfig], is_open_field: bool = False, is_low_density: bool = False,
                machine_scale: MachineScale = MachineScale.IEC61217, bb_proximity_mm: float = 10):
        self.machine_scale = machine_scale
        self.bb_arrangement = tuple(bb_arrangement)
        for img in self.images:
            img.analyze(bb_arrangement=self.bb_arrangement, is_open_field=is_open_field, is_low_density=is_low_density,
                        bb_proximity_mm=bb_proximity_mm, machine_scale=machine_scale)
        self.bbs = []
        for cfg in self.bb_arrangement:
            matches = [img.arrangement_matches[cfg.name] for img in self.images if cfg.name in img.arrangement_matches]
            self.bbs.append(BB3D(bb_config=cfg, bb_matches=matches, scale=machine_scale))
        self._is_analyzed = True

    def cax2bb_distance(self, metric: str = "max") -> float:
        """winston_lutz.py:1776-1796 over every match of every image"""
        d = [m.bb_field_distance_mm for img in self.images for m in img.arrangement_matches.values()]
        return float({"max": np.max, "median": np.median, "mean": np.mean}[metric](d))

    @property
    def max_bb_deviation_2d(self) -> float:
        return self.cax2bb_distance("max")

    @property
    def mean_bb_deviation_2d(self) -> float:
        return self.cax2bb_distance("mean")

    @property
    def median_bb_deviation_2d(self) -> float:
        return self.cax2bb_distance("median")

    @property
    def bb_shift_vector(self):
        """winston_lutz.py:2937-2962 -> (translation Vector, yaw, pitch, roll)"""
        return align_points(measured_points=[bb.measured_bb_position for bb in self.bbs],
                            ideal_points=[bb.measured_field_position for bb in self.bbs])

    def bb_shift_instructions(self) -> str:
        t, yaw, pitch, roll = self.bb_shift_vector
        x_dir = "LEFT" if t.x < 0 else "RIGHT"
        y_dir = "IN" if t.y > 0 else "OUT"
        z_dir = "UP" if t.z > 0 else "DOWN"
        return (f"{x_dir} {abs(t.x):2.2f}mm; {y_dir} {abs(t.y):2.2f}mm; {z_dir} {abs(t.z):2.2f}mm; Rotation {yaw:2.2f}°; "
                f"Pitch {pitch:2.2f}°; Roll {roll:2.2f}°")

    def _generate_results_data(self) -> WinstonLutzMultiTargetMultiFieldResult:
        if not self._is_analyzed:
            raise ValueError("The set is not analyzed. Use .analyze() first.")
        bb_maxes = {}
        for cfg in self.bb_arrangement:
            max_d = 0.0
            for img in self.images:
                if cfg.name in img.arrangement_matches:
                    max_d = max(max_d, img.arrangement_matches[cfg.name].bb_field_distance_mm)
            bb_maxes[cfg.name] = max_d
        t, yaw, pitch, roll = self.bb_shift_vector
        return WinstonLutzMultiTargetMultiFieldResult(
            num_total_images=len(self.images), max_2d_field_to_bb_mm=self.max_bb_deviation_2d, mean_2d_field_to_bb_mm=self.mean_bb_deviation_2d,
            median_2d_field_to_bb_mm=self.median_bb_deviation_2d, bb_maxes=bb_maxes, bb_arrangement=self.bb_arrangement,
            bb_shift_vector={"x": t.x, "y": t.y, "z": t.z}, bb_shift_yaw=yaw, bb_shift_pitch=pitch, bb_shift_roll=roll)
