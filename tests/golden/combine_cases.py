"""Seeded inputs of the image-combination / resampling goldens (make_combine_golden.py, tests/test_combine.py,
tests/test_gpu_combine.py)."""
from __future__ import annotations

import numpy as np

from oracle import synth

PS, SID = 0.390625, 1000.0


def pf_parts():
    """Two exposures of one picket fence: the even pickets on the first image, the odd ones on the second (the use case of
    PicketFence.from_multiple_images, picketfence.py:357-400), with different panel gains so that ``stretch_each`` matters."""
    offs = np.random.default_rng(41).uniform(-0.4, 0.4, 10)
    parts = []
    for k, gain in ((0, 1.0), (1, 0.6)):
        fr = synth.epid1024()
        for i in range(k, 10, 2):
            fr.add_filtered_field((300, 3), (0, (i - 4.5) * 15 + offs[i]))
        fr.gaussian(1.0)
        fr.noise(0.002, seed=300 + k)
        parts.append(np.clip(fr.image.astype(np.float64) * gain, 0, 65535).astype(np.uint16))
    return parts


def star_parts():
    """Two half-sets of spokes of one starshot (Starshot.from_multiple_images, starshot.py:148-174): the spokes within 45 degrees
    of the horizontal on one image, the others on the second, again with different gains."""
    from tests.golden.starshot_cases import case_frame

    a = case_frame("offset6")[0].astype(np.float64)
    base = float(np.median(a))
    h, w = a.shape
    yy, xx = np.mgrid[:h, :w]
    ang = np.degrees(np.arctan2(yy - h / 2, xx - w / 2)) % 180
    first = (ang < 45) | (ang >= 135)
    p1 = np.where(first, a, base)
    p2 = np.where(~first, a, base) * 0.7
    return [np.clip(p, 0, 65535).astype(np.uint16) for p in (p1, p2)]


def small_stack():
    r = np.random.default_rng(17)
    return [r.integers(100, 4000, (37, 53)).astype(np.uint16), r.integers(0, 65535, (37, 53)).astype(np.uint16),
            (r.random((37, 53)) * 3.5 + 1.25)]


ZOOM_CASES = {
    # name: (shape, zoom, order, mode)
    "z2d_up_cubic": ((41, 57), 1.37, 3, "constant"),
    "z2d_down_cubic": ((64, 48), 0.61, 3, "constant"),
    "z2d_up_linear": ((33, 29), 2.0, 1, "constant"),
    "z2d_near_cubic": ((40, 40), 1.25, 3, "nearest"),
    "z1d_up_cubic_nearest": ((101,), 10.0, 3, "nearest"),
    "z1d_down_cubic_nearest": ((257,), 0.37, 3, "nearest"),
    "z1d_up_linear_nearest": ((64,), 3.3, 1, "nearest"),
    "z2d_big": ((384, 512), 1.1, 3, "constant"),
}


def zoom_input(name):
    shape = ZOOM_CASES[name][0]
    r = np.random.default_rng(sum(map(ord, name)))
    a = r.random(shape) * 1000
    if len(shape) == 2:
        yy, xx = np.mgrid[:shape[0], :shape[1]]
        a += 500 * np.sin(yy / 7.0) * np.cos(xx / 5.0)
    else:
        a += 500 * np.sin(np.arange(shape[0]) / 9.0)
    return a


def equate_inputs():
    """A 0.25 mm/px 480 x 640 'film' and a 0.4 mm/px 360 x 300 'EPID' of different physical size (core/image.py:169-220)."""
    r = np.random.default_rng(23)
    yy, xx = np.mgrid[:480, :640]
    a = 1000 + 800 * np.exp(-((yy - 240) ** 2 + (xx - 320) ** 2) / (2 * 90.0**2)) + r.normal(0, 3, (480, 640))
    yy, xx = np.mgrid[:360, :300]
    b = 1000 + 800 * np.exp(-((yy - 180) ** 2 + (xx - 150) ** 2) / (2 * 56.0**2)) + r.normal(0, 3, (360, 300))
    return (a, 25.4 / 0.25), (b, 25.4 / 0.4)
