"""Generate tests/golden/mtmf_golden.npz: the UNMODIFIED reference WinstonLutzMultiTargetMultiField (stub-imported; skimage served by
oracle/skimage_shim.py) on the seeded sets of mtmf_cases.py.

Run here (the container that has /root/reference):  python -m tests.golden.make_mtmf_golden
"""
from __future__ import annotations

import hashlib
import sys
import threading
import time
import warnings

import numpy as np

from tests.golden.mtmf_cases import SETS, set_frames


def reference_mtmf(frames, ps, sid, axes, arr):
    from oracle import skimage_shim
    from oracle.refstub import reference_image_from_array

    skimage_shim.install()
    from pylinac import winston_lutz as wl

    cfg = tuple(wl.BBConfig(name=n, offset_left_mm=l, offset_up_mm=u, offset_in_mm=i, bb_size_mm=d, rad_size_mm=r) for n, l, u, i, d, r in arr)
    st = wl.WinstonLutzMultiTargetMultiField.__new__(wl.WinstonLutzMultiTargetMultiField)
    st.images = [reference_image_from_array(wl.WinstonLutzMultiTargetMultiFieldImage, np.array(f), ps, sid=sid, gantry=g, coll=c, couch=p)
                 for f, (g, c, p) in zip(frames, axes)]
    st._captured_warnings, st._warnings_lock = [], threading.Lock()
    st._is_analyzed = False
    st.is_from_cbct = False
    st.analyze(bb_arrangement=cfg)
    rd = st.results_data()
    out = {}
    names = [c.name for c in cfg]
    out["names"] = np.array(names)
    out["bb_px"] = np.array([[[img.arrangement_matches[n].bb.x, img.arrangement_matches[n].bb.y] for n in names] for img in st.images])
    out["field_px"] = np.array([[[img.arrangement_matches[n].field.x, img.arrangement_matches[n].field.y] for n in names] for img in st.images])
    out["epid_px"] = np.array([[img.epid.x, img.epid.y] for img in st.images])
    out["shape"] = np.array([img.shape for img in st.images])
    out["measured_bb"] = np.array([[b.measured_bb_position.x, b.measured_bb_position.y, b.measured_bb_position.z] for b in st.bbs])
    out["measured_field"] = np.array([[b.measured_field_position.x, b.measured_field_position.y, b.measured_field_position.z] for b in st.bbs])
    t, yaw, pitch, roll = st.bb_shift_vector
    out["shift"] = np.array([t.x, t.y, t.z, yaw, pitch, roll])
    out["max_2d"] = np.array(rd.max_2d_field_to_bb_mm)
    out["mean_2d"] = np.array(rd.mean_2d_field_to_bb_mm)
    out["median_2d"] = np.array(rd.median_2d_field_to_bb_mm)
    out["bb_maxes"] = np.array([rd.bb_maxes[n] for n in names])
    out["instructions"] = np.array(st.bb_shift_instructions())
    return out


def main():
    store = {}
    warnings.simplefilter("ignore")
    for name in SETS:
        frames, ps, sid, axes, arr = set_frames(name)
        store[f"{name}/input_sha1"] = np.frombuffer(hashlib.sha1(frames.tobytes()).digest(), dtype=np.uint8)
        t = time.time()
        ref = reference_mtmf(frames, ps, sid, axes, arr)
        for k, v in ref.items():
            store[f"{name}/{k}"] = v
        print(name, round(time.time() - t, 1), "s shift", np.round(ref["shift"], 4).tolist(), "max 2d", float(ref["max_2d"]))
    np.savez_compressed("tests/golden/mtmf_golden.npz", **store)


if __name__ == "__main__":
    sys.exit(main())
