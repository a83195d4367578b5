"""Seeded synthetic Winston-Lutz cases (docs/source/winston_lutz.rst:911-1020 recipe through oracle/synth.py)."""
from __future__ import annotations

from oracle import synth

CASES = ["g0", "g90", "g180", "g270", "couch45", "big_offset", "noisy", "as1200", "bb8", "field30", "inverted", "open_field",
         "fff", "low_density"]


def case_frame(name):
    """-> (frame uint16, pixel_spacing_mm, sid, gantry, coll, couch, analyze_kwargs)"""
    base = dict(noise_sigma=0.002)
    if name.startswith("g") and name[1:].isdigit():
        g = float(name[1:])
        fr = synth.epid1024()
        return synth.winstonlutz_frame(fr, offset_mm_left=0.7, offset_mm_up=-0.4, offset_mm_in=0.3, gantry=g, seed=int(g) + 1, **base), fr.pixel_size, 1000.0, g, 0.0, 0.0, {}
    if name == "couch45":
        fr = synth.epid1024()
        return synth.winstonlutz_frame(fr, offset_mm_left=-0.5, offset_mm_up=0.6, offset_mm_in=-0.8, gantry=0, couch=45, seed=51, **base), fr.pixel_size, 1000.0, 0.0, 0.0, 45.0, {}
    if name == "big_offset":
        fr = synth.epid1024()
        return synth.winstonlutz_frame(fr, offset_mm_left=2.5, offset_mm_up=1.0, offset_mm_in=-2.0, gantry=30, seed=52, **base), fr.pixel_size, 1000.0, 30.0, 0.0, 0.0, {}
    if name == "noisy":
        fr = synth.epid1024()
        return synth.winstonlutz_frame(fr, offset_mm_left=0.3, offset_mm_up=0.3, offset_mm_in=0.3, gantry=120, noise_sigma=0.01, seed=53), fr.pixel_size, 1000.0, 120.0, 0.0, 0.0, {}
    if name == "as1200":
        fr = synth.as1200(1500.0)
        return synth.winstonlutz_frame(fr, offset_mm_left=0.9, offset_mm_up=-0.2, offset_mm_in=0.5, gantry=200, seed=54, **base), fr.pixel_size, 1500.0, 200.0, 0.0, 0.0, {}
    if name == "bb8":
        fr = synth.epid1024()
        return synth.winstonlutz_frame(fr, bb_size_mm=8.0, field_size_mm=(30, 30), offset_mm_left=-0.6, offset_mm_up=0.2, gantry=300, seed=55, **base), fr.pixel_size, 1000.0, 300.0, 0.0, 0.0, {"bb_size_mm": 8}
    if name == "field30":
        fr = synth.epid1024()
        return synth.winstonlutz_frame(fr, field_size_mm=(30, 25), offset_mm_left=0.4, offset_mm_in=0.7, gantry=60, coll=20, seed=56, **base), fr.pixel_size, 1000.0, 60.0, 20.0, 0.0, {}
    if name == "inverted":
        fr = synth.epid1024()
        synth.winstonlutz_frame(fr, offset_mm_left=0.5, offset_mm_up=0.5, gantry=0, seed=57, **base)
        return fr.inverted(), fr.pixel_size, 1000.0, 0.0, 0.0, 0.0, {}
    if name == "open_field":
        fr = synth.epid1024()
        return synth.winstonlutz_frame(fr, field_size_mm=(150, 150), offset_mm_left=1.0, offset_mm_up=-1.0, gantry=0, seed=58, **base), fr.pixel_size, 1000.0, 0.0, 0.0, 0.0, {"open_field": True}
    if name == "fff":
        fr = synth.epid1024()
        return synth.winstonlutz_frame(fr, field="fff", field_size_mm=(40, 40), offset_mm_left=-0.8, gantry=90, seed=59, **base), fr.pixel_size, 1000.0, 90.0, 0.0, 0.0, {}
    if name == "low_density":
        fr = synth.epid1024()
        return synth.winstonlutz_frame(fr, bb_alpha=0.5, offset_mm_left=0.6, offset_mm_up=0.4, gantry=0, seed=60, **base), fr.pixel_size, 1000.0, 0.0, 0.0, 0.0, {"low_density_bb": True}
    raise KeyError(name)
