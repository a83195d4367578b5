"""Seeded FieldProfileAnalysis cases (frames of field_cases.py, analyze() keyword variants)."""
from __future__ import annotations

from tests.golden.field_cases import case_frame as field_frame

CASES = {
    # name: (field_cases frame, analyze kwargs with enum VALUES as strings)
    "default_inflection": ("as1200_150", {}),
    "fwhm": ("as1200_150", {"edge_type": "FWHM"}),
    "offset_inflection_wide": ("as1200_offset", {"x_width": 0.02, "y_width": 0.03}),
    "offset_fwhm_wide": ("as1200_offset", {"edge_type": "FWHM", "x_width": 0.02, "y_width": 0.03}),
    "geometric_center": ("geometric", {"centering": "Geometric center", "edge_type": "FWHM"}),
    "manual_position": ("as1200_offset", {"centering": "Manual", "position": (0.45, 0.55), "x_width": 0.01, "y_width": 0.01}),
    "norm_max_noground": ("epid1024_100", {"normalization": "Max", "ground": False, "edge_type": "FWHM"}),
    "norm_beam": ("slope", {"normalization": "Beam center"}),
    "inverted_frame": ("inverted", {}),
    "fff": ("fff", {"edge_type": "FWHM"}),
    "hill": ("as1200_150", {"edge_type": "Inflection Hill"}),
    "hill_fff_offset": ("hill_fff_siemens", {"edge_type": "Inflection Hill", "x_width": 0.02, "y_width": 0.02, "hill_window_ratio": 0.15}),
    # enum MEMBERS (resolved by name in case()): only these normalise in the reference, plain strings do not
    "enum_norm_max": ("epid1024_100", {"normalization": "Normalization.MAX", "edge_type": "FWHM"}),
    "enum_norm_beam": ("as1200_offset", {"normalization": "Normalization.BEAM_CENTER", "edge_type": "FWHM", "ground": False}),
    "enum_norm_geometric": ("as1200_150", {"normalization": "Normalization.GEOMETRIC_CENTER"}),
}


def case(name):
    """-> (frame uint16, pixel_spacing_mm, sid, analyze kwargs)"""
    fname, kw = CASES[name]
    a, ps, sid, _ = field_frame(fname)
    return a, ps, sid, dict(kw)


def resolve_enums(kw, normalization_enum):
    """'Normalization.MAX' -> the enum member of whichever package (reference or pylinac_b200) runs the case"""
    kw = dict(kw)
    v = kw.get("normalization")
    if isinstance(v, str) and v.startswith("Normalization."):
        kw["normalization"] = normalization_enum[v.split(".")[1]]
    return kw
