"""Seeded synthetic multi-target multi-field Winston-Lutz sets (winston_lutz.py:2728-2870), frames from oracle/synth.py following
the reference generator's placement rule (image_generator/utils.py:440-500: field and BB at cax_offset (-long, gplane) of the
projected phantom offsets)."""
from __future__ import annotations

import numpy as np

from oracle import synth

# name: (arrangement rows (name, left, up, in, bb_size, rad_size), BB error mm (left, up, in) per BB, axes)
ARR3 = [("Iso", 0, 0, 0, 5, 20), ("A", -30, 0, 32, 5, 20), ("B", 30, 0, -42, 5, 20)]
SETS = {
    "three_bbs": (ARR3, [(0.4, -0.3, 0.5), (-0.6, 0.2, 0.1), (0.2, 0.5, -0.4)], [(0, 0, 0), (90, 0, 0), (180, 0, 0), (270, 0, 0)]),
    "couch_kick": (ARR3, [(0.0, 0.0, 0.0), (0.5, 0.0, 0.3), (-0.4, 0.3, 0.0)], [(0, 0, 0), (0, 0, 30), (0, 0, 330), (90, 0, 0), (270, 0, 0)]),
}


def set_frames(name):
    """-> (frames uint16 [n, h, w], pixel_spacing_mm, sid, axes, arrangement rows)"""
    arr, errs, axes = SETS[name]
    frames, ps = [], None
    for k, (g, c, p) in enumerate(axes):
        fr = synth.epid1024()
        ps = fr.pixel_size
        for (nm, left, up, inn, bb_d, rad) in arr:
            gp, lg = synth.bb_projection_with_rotation(left, up, inn, g, p)
            fr.add_perfect_field((rad, rad), cax_offset_mm=(-lg, gp), alpha=0.7)
        for (nm, left, up, inn, bb_d, rad), e in zip(arr, errs):
            gp, lg = synth.bb_projection_with_rotation(left + e[0], up + e[1], inn + e[2], g, p)
            fr.add_bb(bb_d, cax_offset_mm=(-lg, gp), alpha=-0.35)
        fr.gaussian(0.8)
        fr.noise(0.001, seed=800 + k)
        frames.append(fr.image)
    return np.stack(frames), ps, 1000.0, axes, arr
