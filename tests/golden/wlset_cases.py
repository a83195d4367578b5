"""Seeded synthetic Winston-Lutz SETS (gantry / collimator / couch sweeps around one 3-D BB offset), the inputs of the
set-level solve (winston_lutz.py:1519-1850).  Frames come from oracle/synth.py (docs/source/winston_lutz.rst:911-1020 recipe)."""
from __future__ import annotations

import numpy as np

from oracle import synth

SETS = {
    # name: (BB offset left / up / in [mm], [(gantry, coll, couch), ...])
    "standard": ((0.8, -0.5, 0.6), [(0, 0, 0), (90, 0, 0), (180, 0, 0), (270, 0, 0), (0, 90, 0), (0, 270, 0), (0, 0, 45), (0, 0, 315)]),
    "gantry_only": ((-0.4, 0.9, -0.3), [(0, 0, 0), (45, 0, 0), (135, 0, 0), (225, 0, 0), (315, 0, 0)]),
    "combo": ((0.3, 0.2, -0.7), [(0, 0, 0), (120, 30, 0), (240, 330, 0), (60, 0, 30), (1.5, 0, 358)]),
}

VSHIFT_SETS = ["standard", "combo"]


def set_frames(name):
    """-> (frames uint16 [n,h,w], pixel_spacing_mm, sid, axes)"""
    (left, up, inn), axes = SETS[name]
    frames = []
    ps = None
    for k, (g, c, p) in enumerate(axes):
        fr = synth.epid1024()
        ps = fr.pixel_size
        # the collimator image of a perfect square field is the same field; a small per-image field wobble makes the
        # collimator / couch isocentre sizes non-trivial
        rng = np.random.default_rng(900 + k)
        jitter = rng.uniform(-0.3, 0.3, size=3)
        frames.append(synth.winstonlutz_frame(fr, offset_mm_left=left + jitter[0], offset_mm_up=up + jitter[1],
                                              offset_mm_in=inn + jitter[2], gantry=g, coll=c, couch=p, noise_sigma=0.002,
                                              seed=700 + k))
    return np.stack(frames), ps, 1000.0, axes
