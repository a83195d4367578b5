"""TEST / BENCH INPUT GENERATOR (test infrastructure -- not part of the product path).

Numpy+scipy restatement of the reference's synthetic EPID image generator
(``pylinac/core/image_generator``), which needs scikit-image and therefore
cannot be imported in this image.  The product never imports this module; it is
used by tests, ``bench.py`` (synthetic batches) and ``__graft_entry__.smoke``.

Follows (reference file:line):
  * Simulator / AS500 / AS1000 / AS1200 .......... simulators.py:23-121
  * clip_add / clip_multiply / even_round ........ layers.py:12-33
  * PerfectFieldLayer._create_perfect_field ...... layers.py:216-235 (+ draw_rotated_rectangle :507-554)
  * FilteredFieldLayer.apply (horn gaussian) ..... layers.py:284-303,  gaussian2d :36-56
  * FilterFreeFieldLayer.apply ................... layers.py:342-362
  * PerfectConeLayer / PerfectBBLayer ............ layers.py:113-134, 365-381 (skimage.draw.disk: (r-cr)^2+(c-cc)^2 < R^2)
  * GaussianFilterLayer .......................... layers.py:389-393 (skimage.filters.gaussian(preserve_range=True)
                                                   == ndimage.gaussian_filter(float64, sigma, mode='nearest', truncate=4) -> astype)
  * RandomNoiseLayer ............................. layers.py:403-407 (reference is UNSEEDED; a seed is mandatory here)
  * ConstantLayer / SlopeLayer ................... layers.py:410-445
  * generate_picketfence ......................... utils.py:78-136
  * generate_winstonlutz ......................... utils.py:139-263 (+ winston_lutz.py:3401-3460 bb_projection_with_rotation)

The skimage rasterisation rules (disk strict '<', polygon edge rule) are recalled,
not verified against a skimage install; this does not affect GPU-vs-oracle parity
because both consume the same generated frames.
"""
from __future__ import annotations

import math

import numpy as np
from scipy import ndimage

U16_MAX = 65535


def even_round(num: float) -> int:
    n = int(round(num))
    return n + n % 2


def _clip_to_u16(x):
    return np.clip(x, 0, U16_MAX).astype(np.uint16)


class Frame:
    """A synthetic EPID panel (shape, pixel pitch at the panel, SID)."""

    def __init__(self, shape, pixel_size, sid=1500.0):
        self.shape = tuple(shape)
        self.pixel_size = float(pixel_size)
        self.sid = float(sid)
        self.mag = self.sid / 1000.0
        self.image = np.zeros(self.shape, np.uint16)

    # dots-per-mm at isocentre as DicomImage.dpmm computes it (core/image.py:1534-1547)
    @property
    def dpmm(self):
        return (1.0 / self.pixel_size) * self.sid / 1000.0

    # ------------------------------------------------------------------ fields
    def _rect_mask(self, field_size_mm, cax_offset_mm, rotation=0.0):
        h, w = self.shape
        ext = [even_round(f * self.mag / self.pixel_size) for f in field_size_mm]
        off = [v * self.mag / self.pixel_size for v in cax_offset_mm]
        cy = off[0] + h / 2 - 0.5
        cx = off[1] + w / 2 - 0.5
        x0, x1 = cx - ext[1] / 2, cx + ext[1] / 2
        y0, y1 = cy - ext[0] / 2, cy + ext[0] / 2
        if rotation == 0:
            r_lo, r_hi = max(int(math.ceil(y0)), 0), min(int(math.floor(y1)), h - 1)
            c_lo, c_hi = max(int(math.ceil(x0)), 0), min(int(math.floor(x1)), w - 1)
            mask = np.zeros(self.shape, bool)
            if r_hi >= r_lo and c_hi >= c_lo:
                mask[r_lo : r_hi + 1, c_lo : c_hi + 1] = True
            return mask
        # rotated rectangle: rotate the corner points about the rectangle centre, then
        # half-plane test of every pixel centre against the 4 edges.
        corners = np.array([[x0, y0], [x1, y0], [x1, y1], [x0, y1]], float)
        th = np.radians(rotation)
        c, s = np.cos(th), np.sin(th)
        rot = np.array([[c, -s], [s, c]])
        centre = np.array([cy, cx])  # the reference subtracts `center` given as (row, col) from (x, y) pairs
        pts = np.dot(corners - centre, rot) + centre
        yy, xx = np.mgrid[0:h, 0:w]
        inside = np.ones(self.shape, bool)
        sign = None
        for i in range(4):
            ax, ay = pts[i]
            bx, by = pts[(i + 1) % 4]
            cross = (bx - ax) * (yy - ay) - (by - ay) * (xx - ax)
            if sign is None:
                # orientation from the polygon's signed area
                area = sum(pts[k][0] * pts[(k + 1) % 4][1] - pts[(k + 1) % 4][0] * pts[k][1] for k in range(4))
                sign = 1.0 if area > 0 else -1.0
            inside &= (cross * sign) >= 0
        return inside

    def add_perfect_field(self, field_size_mm, cax_offset_mm=(0, 0), alpha=1.0, rotation=0.0):
        mask = self._rect_mask(field_size_mm, cax_offset_mm, rotation)
        tmp = np.zeros(self.shape)
        tmp[mask] = int(U16_MAX * alpha)
        self.image = _clip_to_u16(self.image.astype(float) + tmp)
        return mask

    def _gauss_about_centre(self, mask, height, sigma_px, constant=0.0):
        rr, cc = np.nonzero(mask)
        cr = (self.shape[0] - 1) / 2.0
        ccx = (self.shape[1] - 1) / 2.0
        g = height * np.exp(-(((cr - rr) / sigma_px) ** 2 + ((ccx - cc) / sigma_px) ** 2) / 2) + constant
        return rr, cc, g

    def add_filtered_field(self, field_size_mm, cax_offset_mm=(0, 0), alpha=1.0, gaussian_height=0.03,
                           gaussian_sigma_mm=32.0, rotation=0.0):
        mask = self.add_perfect_field(field_size_mm, cax_offset_mm, alpha, rotation)
        rr, cc, horns = self._gauss_about_centre(mask, -gaussian_height * U16_MAX, gaussian_sigma_mm / self.pixel_size)
        # ``image[rr, cc] += horns.astype(uint16)``: float->uint16 cast of a negative value wraps
        # modulo 2**16 (C cast through int), and the uint16 add wraps again -> net subtraction.
        self.image[rr, cc] = (self.image[rr, cc].astype(np.int64) + np.trunc(horns).astype(np.int64)) % 65536
        return mask

    def add_fff_field(self, field_size_mm, cax_offset_mm=(0, 0), alpha=1.0, gaussian_height=0.4,
                      gaussian_sigma_mm=80.0, rotation=0.0):
        mask = self.add_perfect_field(field_size_mm, cax_offset_mm, alpha, rotation)
        rr, cc, n = self._gauss_about_centre(mask, gaussian_height * U16_MAX, gaussian_sigma_mm / self.pixel_size,
                                             constant=-gaussian_height * U16_MAX)
        self.image[rr, cc] = (self.image[rr, cc].astype(np.int64) + np.trunc(n).astype(np.int64)) % 65536
        return mask

    # ------------------------------------------------------------------- cones / BBs
    def add_cone(self, cone_size_mm, cax_offset_mm=(0, 0), alpha=1.0, rotation=0.0):
        h, w = self.shape
        radius = self.mag * (cone_size_mm / 2) / self.pixel_size
        x = cax_offset_mm[0] * self.mag / self.pixel_size
        y = cax_offset_mm[1] * self.mag / self.pixel_size
        th = np.radians(rotation)
        oy = x * np.cos(th) - y * np.sin(th)
        ox = x * np.sin(th) + y * np.cos(th)
        cy, cx = oy + (h / 2 - 0.5), ox + (w / 2 - 0.5)
        yy, xx = np.mgrid[0:h, 0:w]
        mask = (yy - cy) ** 2 + (xx - cx) ** 2 < radius**2
        tmp = np.zeros(self.shape)
        tmp[mask] = int(U16_MAX * alpha)
        self.image = _clip_to_u16(self.image.astype(float) + tmp)
        return mask

    def add_bb(self, bb_size_mm=5.0, cax_offset_mm=(0, 0), alpha=-0.5):
        return self.add_cone(bb_size_mm, cax_offset_mm, alpha)

    # ------------------------------------------------------------------- post layers
    def gaussian(self, sigma_mm=2.0):
        sigma_px = sigma_mm / self.pixel_size
        self.image = ndimage.gaussian_filter(self.image.astype(np.float64), sigma_px, mode="nearest", truncate=4.0).astype(np.uint16)

    def noise(self, sigma=0.001, mean=0.0, seed=0):
        rng = np.random.default_rng(seed)
        n = rng.normal(mean, sigma * U16_MAX, size=self.shape)
        self.image = _clip_to_u16(self.image.astype(float) + n)

    def constant(self, value):
        self.image = _clip_to_u16(self.image.astype(float) + value)

    def slope(self, slope_x, slope_y):
        h, w = self.shape
        ys = (1 + slope_y * np.arange(h) / h).reshape(-1, 1)
        xs = (1 + slope_x * np.arange(w) / w).reshape(1, -1)
        tmp = _clip_to_u16(self.image.astype(float) * ys)
        self.image = _clip_to_u16(tmp.astype(float) * xs)

    def inverted(self):
        return (-self.image + self.image.max() + self.image.min()).astype(np.uint16)


# panel presets (simulators.py:103-121) + the 1024x1024 benchmark panel
def as500(sid=1500.0):
    return Frame((384, 512), 0.78125, sid)


def as1000(sid=1500.0):
    return Frame((768, 1024), 0.390625, sid)


def as1200(sid=1500.0):
    return Frame((1280, 1280), 0.336, sid)


def epid1024(sid=1000.0):
    return Frame((1024, 1024), 0.390625, sid)


def picketfence_frame(frame: Frame, *, field="filtered", pickets=10, picket_spacing_mm=20, picket_width_mm=3,
                      picket_height_mm=300, orientation="up_down", picket_offset_error=None, blur_mm=1.0,
                      noise_sigma=0.002, seed=0, leaf_errors=None) -> np.ndarray:
    """utils.py:78-136 restated.  ``leaf_errors``: optional list of (picket, leaf_pos_mm, width_mm, dx_mm)
    extra strips (the docs' 'erroneous leaves' recipe, picketfence.rst)."""
    half = int((pickets - 1) * picket_spacing_mm / 2)
    positions = list(range(-half, half + 1, picket_spacing_mm))
    add = {"filtered": frame.add_filtered_field, "perfect": frame.add_perfect_field, "fff": frame.add_fff_field}[field]
    for idx, pos in enumerate(positions):
        if picket_offset_error is not None:
            pos = pos + picket_offset_error[idx]
        if orientation == "up_down":
            add((picket_height_mm, picket_width_mm), (0, pos))
        else:
            add((picket_width_mm, picket_height_mm), (pos, 0))
    if leaf_errors:
        for (along_mm, leaf_mm, width_mm, height_mm) in leaf_errors:
            if orientation == "up_down":
                frame.add_perfect_field((height_mm, width_mm), (leaf_mm, along_mm))
            else:
                frame.add_perfect_field((width_mm, height_mm), (along_mm, leaf_mm))
    if blur_mm:
        frame.gaussian(blur_mm)
    if noise_sigma:
        frame.noise(noise_sigma, seed=seed)
    return frame.image


def bench_pf_frame(i: int, shape=(1024, 1024)) -> np.ndarray:
    """The BASELINE.json config-2 frame #i: 1024^2, 0.390625 mm pitch, SID 1000
    (dpmm 2.56), 10 pickets x 20 mm, width 3 mm, FilteredFieldLayer, Gaussian 1 mm, noise 0.002 seeded by
    frame index, per-picket offset error U(-0.5, 0.5) mm from default_rng(10_000 + i)."""
    fr = Frame(shape, 0.390625, 1000.0)
    err = np.random.default_rng(10_000 + i).uniform(-0.5, 0.5, size=10)
    return picketfence_frame(fr, pickets=10, picket_spacing_mm=20, picket_width_mm=3, picket_height_mm=300,
                             picket_offset_error=err, blur_mm=1.0, noise_sigma=0.002, seed=i)


def bench_pf_batch(n: int, start: int = 0, shape=(1024, 1024), unique: int | None = None) -> np.ndarray:
    """n frames [n, H, W] uint16.  ``unique`` bounds how many distinct frames are generated
    (the rest repeat cyclically with a per-frame additive offset so they are not identical)."""
    unique = n if unique is None else min(unique, n)
    base = [bench_pf_frame(start + k, shape) for k in range(unique)]
    out = np.empty((n,) + tuple(shape), np.uint16)
    for k in range(n):
        out[k] = base[k % unique]
    return out


# ------------------------------------------------------------------------------ Winston-Lutz
def bb_projection_with_rotation(offset_left, offset_up, offset_in, gantry, couch=0.0, sad=1000.0):
    """winston_lutz.py:3401-3460 restated: isoplane projection of a BB offset (mm) for a gantry / couch position (IEC 61217).
    The reference rotates (up, left, in) with scipy's Rotation.from_euler("xyz", [-couch, 0, gantry]) -- extrinsic rotations about
    x (by -couch) then z (by gantry), i.e. Rz(gantry) @ Rx(-couch) -- and magnifies by sad / (sad - rotated_up).
    tests/test_oracle_synth.py checks this restatement against the reference function itself."""
    a, g = np.radians(-couch), np.radians(gantry)
    rx = np.array([[1, 0, 0], [0, np.cos(a), -np.sin(a)], [0, np.sin(a), np.cos(a)]])
    rz = np.array([[np.cos(g), -np.sin(g), 0], [np.sin(g), np.cos(g), 0], [0, 0, 1]])
    v = rz @ rx @ np.array([offset_up, offset_left, offset_in], float)
    mag = sad / (sad - v[0])
    return -v[1] * mag, v[2] * mag


def winstonlutz_frame(frame: Frame, *, field_size_mm=(20, 20), bb_size_mm=5.0, offset_mm_left=0.0, offset_mm_up=0.0,
                      offset_mm_in=0.0, gantry=0.0, coll=0.0, couch=0.0, field_alpha=1.0, bb_alpha=-0.8, blur_mm=1.5,
                      noise_sigma=0.0, seed=0, field="perfect") -> np.ndarray:
    add = {"filtered": frame.add_filtered_field, "perfect": frame.add_perfect_field, "fff": frame.add_fff_field}[field]
    add(field_size_mm, (0, 0), alpha=field_alpha, rotation=coll)
    gplane, long_off = bb_projection_with_rotation(offset_mm_left, offset_mm_up, offset_mm_in, gantry, couch)
    frame.add_bb(bb_size_mm, (-long_off, gplane), bb_alpha)
    if blur_mm:
        frame.gaussian(blur_mm)
    if noise_sigma:
        frame.noise(noise_sigma, seed=seed)
    return frame.image


# ------------------------------------------------------------------------------ Starshot / open field
def starshot_frame(frame: Frame, *, spokes=6, spoke_mm=(270, 5), alpha=0.5, blur_mm=3.0, offsets_mm=None,
                   noise_sigma=0.0, seed=0) -> np.ndarray:
    """docs/source/starshot_docs.rst:256-263 recipe, with the rotation done on the layer (rotation about the
    image centre) instead of ndimage.rotate of the accumulated image."""
    for k in range(spokes):
        ang = 180.0 * k / spokes
        off = (0.0, 0.0) if offsets_mm is None else offsets_mm[k]
        frame.add_filtered_field(spoke_mm, off, alpha=alpha, rotation=ang)
    if blur_mm:
        frame.gaussian(blur_mm)
    if noise_sigma:
        frame.noise(noise_sigma, seed=seed)
    return frame.image


def openfield_frame(frame: Frame, *, field_size_mm=(150, 150), cax_offset_mm=(0, 0), blur_mm=2.0, noise_sigma=0.001,
                    seed=0, slope=None, field="filtered") -> np.ndarray:
    add = {"filtered": frame.add_filtered_field, "perfect": frame.add_perfect_field, "fff": frame.add_fff_field}[field]
    add(field_size_mm, cax_offset_mm)
    if slope is not None:
        frame.slope(*slope)
    if blur_mm:
        frame.gaussian(blur_mm)
    if noise_sigma:
        frame.noise(noise_sigma, seed=seed)
    return frame.image
