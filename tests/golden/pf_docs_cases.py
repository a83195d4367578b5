"""The reference's own PicketFence fixtures: the seven generated DICOM files under docs/source/files/ with the analyze() calls of
their recipes (docs/source/picketfence.rst:455-730).  1280 x 1280, AS1200 pitch 0.336 mm, SID 1000.

Frames: four of them are committed (lzma-compressed pixel data) in tests/golden/pf_docs_frames.npz so the GPU tests can run
them on a box without /root/reference; the other three (noise makes them ~2 MB each) are read from /root/reference when it is
present (this container: the CPU oracle tests) and skipped otherwise.
"""
from __future__ import annotations

import lzma
import os

import numpy as np

REF_DIR = "/root/reference/docs/source/files"
FRAMES_NPZ = os.path.join(os.path.dirname(__file__), "pf_docs_frames.npz")
PIXEL_MM, SID, SHAPE = 0.336, 1000.0, (1280, 1280)

DOCS = {
    # name: analyze kwargs of the docs recipe
    "perfect_up_down": {"separate_leaves": False, "nominal_gap_mm": 4},
    "perfect_left_right": {"separate_leaves": False, "nominal_gap_mm": 4},
    "noisy_wide_gap_up_down": {},
    "separated_wide_gap_up_down": {"separate_leaves": True, "nominal_gap_mm": 10},
    "rotated_up_down": {"separate_leaves": False, "nominal_gap_mm": 5},
    "offset_picket": {},
    "erroneous_leaves": {"separate_leaves": True, "nominal_gap_mm": 5},
}
COMMITTED = ["perfect_up_down", "perfect_left_right", "rotated_up_down", "erroneous_leaves"]


def _read_pixel_tail(path):
    """Uncompressed little-endian DICOM: PixelData is the last element."""
    n = SHAPE[0] * SHAPE[1] * 2
    with open(path, "rb") as f:
        f.seek(-n, 2)
        return np.frombuffer(f.read(n), dtype=np.uint16).reshape(SHAPE).copy()


def available(name) -> bool:
    return name in COMMITTED and os.path.exists(FRAMES_NPZ) or os.path.exists(os.path.join(REF_DIR, name + ".dcm"))


def docs_frame(name):
    """-> (frame uint16 [1280,1280], pixel_spacing_mm, sid, analyze_kwargs)"""
    p = os.path.join(REF_DIR, name + ".dcm")
    if name in COMMITTED and os.path.exists(FRAMES_NPZ):
        with np.load(FRAMES_NPZ) as z:
            a = np.frombuffer(lzma.decompress(z[name].tobytes()), dtype=np.uint16).reshape(SHAPE).copy()
    elif os.path.exists(p):
        a = _read_pixel_tail(p)
    else:
        raise FileNotFoundError(name)
    return a, PIXEL_MM, SID, dict(DOCS[name])
