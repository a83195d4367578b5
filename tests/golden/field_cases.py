"""Seeded synthetic FieldAnalysis cases shared by the golden generator, the oracle tests and the GPU parity tests
(oracle/synth.py open-field generator: filtered / flattening-filter-free field, gaussian blur, seeded noise)."""
from __future__ import annotations

from oracle import synth

CASES = ["as1200_150", "as1200_offset", "epid1024_100", "fwhm_edges", "geometric", "manual_strips", "siemens", "elekta",
         "inverted", "no_interp", "fff", "norm_max", "slope", "no_ground"]


# Edge.INFLECTION_HILL / Interpolation.SPLINE: reference goldens only (the numpy oracle restates the LINEAR / derivative engine)
EXT_CASES = ["hill", "hill_fff_siemens", "spline", "hill_spline_inverted"]


def case_frame(name):
    """-> (frame uint16, pixel_spacing_mm, sid, analyze_kwargs)"""
    if name == "as1200_150":
        fr = synth.as1200(1000.0)
        return synth.openfield_frame(fr, seed=31), fr.pixel_size, 1000.0, {}
    if name == "as1200_offset":
        fr = synth.as1200(1000.0)
        return synth.openfield_frame(fr, field_size_mm=(120, 180), cax_offset_mm=(7.3, -4.1), seed=32), fr.pixel_size, 1000.0, {}
    if name == "epid1024_100":
        fr = synth.epid1024()
        return synth.openfield_frame(fr, field_size_mm=(100, 100), seed=33), fr.pixel_size, 1000.0, {}
    if name == "fwhm_edges":
        fr = synth.as1200(1000.0)
        return synth.openfield_frame(fr, seed=34), fr.pixel_size, 1000.0, {"edge_detection_method": "FWHM"}
    if name == "geometric":
        fr = synth.as1200(1000.0)
        return synth.openfield_frame(fr, cax_offset_mm=(5.0, 3.0), seed=35), fr.pixel_size, 1000.0, {"centering": "Geometric center"}
    if name == "manual_strips":
        fr = synth.as1200(1000.0)
        return synth.openfield_frame(fr, seed=36), fr.pixel_size, 1000.0, {"centering": "Manual", "vert_position": 0.45, "horiz_position": 0.55,
                                                                           "vert_width": 0.02, "horiz_width": 0.03}
    if name == "siemens":
        fr = synth.as1200(1000.0)
        return synth.openfield_frame(fr, seed=37), fr.pixel_size, 1000.0, {"protocol": "SIEMENS"}
    if name == "elekta":
        fr = synth.as1200(1000.0)
        return synth.openfield_frame(fr, seed=38), fr.pixel_size, 1000.0, {"protocol": "ELEKTA", "in_field_ratio": 0.7}
    if name == "inverted":
        fr = synth.as1200(1000.0)
        synth.openfield_frame(fr, seed=39)
        return fr.inverted(), fr.pixel_size, 1000.0, {}
    if name == "no_interp":
        fr = synth.as1200(1000.0)
        return synth.openfield_frame(fr, seed=40), fr.pixel_size, 1000.0, {"interpolation": None}
    if name == "fff":
        fr = synth.as1200(1000.0)
        return synth.openfield_frame(fr, field="fff", seed=41), fr.pixel_size, 1000.0, {"is_FFF": True}
    if name == "norm_max":
        fr = synth.as1200(1000.0)
        return synth.openfield_frame(fr, seed=42), fr.pixel_size, 1000.0, {"normalization_method": "Max", "penumbra": (10, 90)}
    if name == "slope":
        fr = synth.as1200(1000.0)
        return synth.openfield_frame(fr, seed=43, slope=(0.1, 0.05)), fr.pixel_size, 1000.0, {"slope_exclusion_ratio": 0.3}
    if name == "no_ground":
        fr = synth.as1200(1000.0)
        return synth.openfield_frame(fr, seed=44), fr.pixel_size, 1000.0, {"ground": False, "interpolation_resolution_mm": 0.25}
    if name == "hill":
        fr = synth.as1200(1000.0)
        return synth.openfield_frame(fr, seed=51), fr.pixel_size, 1000.0, {"edge_detection_method": "Inflection Hill"}
    if name == "hill_fff_siemens":
        fr = synth.as1200(1000.0)
        return synth.openfield_frame(fr, field="fff", field_size_mm=(140, 160), cax_offset_mm=(3.0, -2.0), seed=52), fr.pixel_size, 1000.0, \
            {"edge_detection_method": "Inflection Hill", "is_FFF": True, "protocol": "SIEMENS", "hill_window_ratio": 0.2, "penumbra": (10, 90)}
    if name == "spline":
        fr = synth.epid1024()
        return synth.openfield_frame(fr, field_size_mm=(110, 90), seed=53), fr.pixel_size, 1000.0, {"interpolation": "Spline"}
    if name == "hill_spline_inverted":
        fr = synth.as1200(1000.0)
        synth.openfield_frame(fr, seed=54, field_size_mm=(100, 130))
        return fr.inverted(), fr.pixel_size, 1000.0, {"edge_detection_method": "Inflection Hill", "interpolation": "Spline",
                                                       "protocol": "ELEKTA", "interpolation_resolution_mm": 0.2}
    raise KeyError(name)
