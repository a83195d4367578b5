"""Seeded synthetic cases for the whole-frame locators (GlobalSizedDiskLocator / GlobalSizedFieldLocator / GlobalFieldLocator,
metrics/image.py:275-354, 727-956), shared by the golden generator and the GPU parity tests.  Frames are small (the restated
skimage of the golden run labels the whole frame ~50 times in Python)."""
from __future__ import annotations

import numpy as np

from oracle import synth

CASES = ["disks4", "disks_bright", "disks_cap2", "fields4", "fields_px", "fields_any", "fields_none"]


def _frame(shape=(480, 600), pixel=0.4):
    return synth.Frame(shape, pixel, sid=1000.0)


def case(name):
    """-> (frame uint16, pixel_spacing_mm, sid, metric spec dict)"""
    if name in ("disks4", "disks_cap2"):
        fr = _frame()
        fr.add_perfect_field((170, 220), alpha=0.5)
        for off in ((-40, -60), (-35, 55), (30, -20), (45, 70)):
            fr.add_bb(6.0, cax_offset_mm=off, alpha=-0.22)
        fr.gaussian(0.8)
        fr.noise(0.0015, seed=501)
        spec = {"cls": "GlobalSizedDiskLocator", "kw": {"radius_mm": 3.0, "radius_tolerance_mm": 0.8}}
        if name == "disks_cap2":
            spec["kw"].update(max_number=2, min_separation_mm=8)
        return fr.image, fr.pixel_size, 1000.0, spec
    if name == "disks_bright":
        fr = _frame((400, 400), 0.5)
        fr.constant(4000)
        for off, d in (((-50, -40), 8.0), ((20, 45), 8.0), ((55, -55), 8.0)):
            fr.add_cone(d, cax_offset_mm=off, alpha=0.4)
        fr.gaussian(1.0)
        fr.noise(0.001, seed=502)
        return fr.image, fr.pixel_size, 1000.0, {"cls": "GlobalSizedDiskLocator", "kw": {"radius_mm": 4.0, "radius_tolerance_mm": 1.0, "invert": False}}
    if name in ("fields4", "fields_px", "fields_none"):
        fr = _frame()
        for off in ((-45, -70), (-40, 60), (35, -30), (50, 75)):
            fr.add_perfect_field((20, 20), cax_offset_mm=off, alpha=0.6)
        fr.gaussian(1.0)
        fr.noise(0.001, seed=503)
        if name == "fields4":
            spec = {"cls": "GlobalSizedFieldLocator", "physical": True, "kw": {"field_width_mm": 20, "field_height_mm": 20, "field_tolerance_mm": 3, "max_number": 4}}
        elif name == "fields_px":
            spec = {"cls": "GlobalSizedFieldLocator", "physical": False, "kw": {"field_width_px": 50, "field_height_px": 50, "field_tolerance_px": 6}}
        else:
            spec = {"cls": "GlobalSizedFieldLocator", "physical": True, "raises": True,
                    "kw": {"field_width_mm": 40, "field_height_mm": 40, "field_tolerance_mm": 2}}
        return fr.image, fr.pixel_size, 1000.0, spec
    if name == "fields_any":
        fr = _frame()
        fr.add_perfect_field((15, 15), cax_offset_mm=(-40, -60), alpha=0.5)
        fr.add_perfect_field((30, 24), cax_offset_mm=(30, 40), alpha=0.7)
        fr.gaussian(1.0)
        fr.noise(0.001, seed=504)
        return fr.image, fr.pixel_size, 1000.0, {"cls": "GlobalFieldLocator", "physical": False, "kw": {"max_number": 2}}
    raise KeyError(name)
