"""Generate tests/golden/fpa_golden.npz: the UNMODIFIED reference FieldProfileAnalysis (stub-imported) on the cases of fpa_cases.py.

Run here (the container that has /root/reference):  python -m tests.golden.make_fpa_golden
"""
from __future__ import annotations

import hashlib
import sys
import warnings

import numpy as np

from tests.golden.fpa_cases import CASES, case, resolve_enums


def reference_fpa(frame, ps, sid, kw):
    from oracle.refstub import import_reference

    import_reference()
    from pylinac import field_profile_analysis as fpa

    from pylinac.core.profile import Normalization

    f = fpa.FieldProfileAnalysis(np.array(frame), dpi=25.4 / ps, sid=sid)
    f.analyze(**resolve_enums(kw, Normalization))
    from oracle import skimage_shim

    skimage_shim.install()           # RectangleROI statistics rasterise with skimage.draw.polygon (restated, unpinned at that boundary)
    out = {"center": np.array([f.center_rect.mean, f.center_rect.std, f.center_rect.min, f.center_rect.max], dtype=float)}
    for ax, p in (("x", f.x_profile), ("y", f.y_profile)):
        out[f"{ax}/metric_names"] = np.array(list(p.metric_values.keys()))
        out[f"{ax}/metric_values"] = np.array([float(v) for v in p.metric_values.values()])
        out[f"{ax}/field_width_mm"] = np.array(float(p.field_width_mm))
        out[f"{ax}/center_idx"] = np.array(float(p.center_idx))
        out[f"{ax}/cax_index"] = np.array(float(p.cax_index))
        out[f"{ax}/edges"] = np.array([float(p.field_edge_idx("left")), float(p.field_edge_idx("right"))])
        out[f"{ax}/values"] = np.asarray(p.values, dtype=float)
    return out


def main():
    store = {}
    warnings.simplefilter("ignore")
    for name in CASES:
        a, ps, sid, kw = case(name)
        store[f"{name}/input_sha1"] = np.frombuffer(hashlib.sha1(a.tobytes()).digest(), dtype=np.uint8)
        ref = reference_fpa(a, ps, sid, kw)
        for k, v in ref.items():
            store[f"{name}/{k}"] = v
        print(name, dict(zip(ref["x/metric_names"].tolist(), np.round(ref["x/metric_values"], 4).tolist())), float(ref["x/center_idx"]))
    np.savez_compressed("tests/golden/fpa_golden.npz", **store)


if __name__ == "__main__":
    sys.exit(main())
