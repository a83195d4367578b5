"""Generate tests/golden/wlset_golden.npz: the UNMODIFIED reference WinstonLutz set-level analysis (stub-imported; skimage
served by oracle/skimage_shim.py) on the seeded synthetic sets of wlset_cases.py.

Run here (the container that has /root/reference):  python -m tests.golden.make_wlset_golden
"""
from __future__ import annotations

import hashlib
import sys
import warnings

import numpy as np

from tests.golden.wlset_cases import SETS, VSHIFT_SETS, set_frames

SCALARS = ["max_2d_cax_to_bb_mm", "median_2d_cax_to_bb_mm", "mean_2d_cax_to_bb_mm", "max_2d_cax_to_epid_mm",
           "median_2d_cax_to_epid_mm", "mean_2d_cax_to_epid_mm", "gantry_3d_iso_diameter_mm", "coll_2d_iso_diameter_mm",
           "couch_2d_iso_diameter_mm", "gantry_coll_3d_iso_diameter_mm", "num_total_images", "num_gantry_images",
           "num_coll_images", "num_couch_images", "num_gantry_coll_images", "max_gantry_rms_deviation_mm",
           "max_epid_rms_deviation_mm", "max_coll_rms_deviation_mm", "max_couch_rms_deviation_mm"]


def reference_wlset(frames, ps, sid, axes, **analyze_kwargs):
    from oracle import skimage_shim
    from oracle.refstub import reference_image_from_array

    skimage_shim.install()
    from pylinac import winston_lutz as wl

    st = wl.WinstonLutz.__new__(wl.WinstonLutz)
    st.images = [reference_image_from_array(wl.WinstonLutz2D, np.array(f), ps, sid=sid, gantry=g, coll=c, couch=p)
                 for f, (g, c, p) in zip(frames, axes)]
    import threading

    st._captured_warnings, st._warnings_lock = [], threading.Lock()      # WarningCollectorMixin.__init__ (core/warnings.py:14-17)
    st._is_analyzed = False
    st.is_from_cbct = False
    st.analyze(**analyze_kwargs)
    rd = st.results_data()
    out = {k: np.asarray(getattr(rd, k)) for k in SCALARS}
    sv = st.bb_shift_vector
    out["bb_shift_vector"] = np.array([sv.x, sv.y, sv.z], dtype=float)
    out["measured_bb_position"] = np.array([st.bb.measured_bb_position.x, st.bb.measured_bb_position.y, st.bb.measured_bb_position.z])
    out["measured_field_position"] = np.array([st.bb.measured_field_position.x, st.bb.measured_field_position.y, st.bb.measured_field_position.z])
    out["variable_axes"] = np.array([str(i.variable_axis.value) for i in st.images])
    out["cax2bb_distances"] = np.array([i.cax2bb_distance for i in st.images])
    out["bbs"] = np.array([[i.bb.x, i.bb.y] for i in st.images])
    out["fields"] = np.array([[i.field_cax.x, i.field_cax.y] for i in st.images])
    out["epids"] = np.array([[i.epid.x, i.epid.y] for i in st.images])
    out["dpmm"] = np.array(st.images[0].dpmm)
    out["keys"] = np.array(list(rd.keyed_image_details.keys()))
    out["shift_instructions"] = np.array(st.bb_shift_instructions())
    return out


def main():
    store = {}
    warnings.simplefilter("ignore")
    for name in SETS:
        frames, ps, sid, axes = set_frames(name)
        store[f"{name}/input_sha1"] = np.frombuffer(hashlib.sha1(frames.tobytes()).digest(), dtype=np.uint8)
        ref = reference_wlset(frames, ps, sid, axes)
        for k, v in ref.items():
            store[f"{name}/{k}"] = v
        print(name, {k: (float(v) if v.ndim == 0 and v.dtype.kind == "f" else v.tolist()) for k, v in ref.items() if k in
                     ("gantry_3d_iso_diameter_mm", "coll_2d_iso_diameter_mm", "couch_2d_iso_diameter_mm", "bb_shift_vector", "variable_axes")})
    # analyze(apply_virtual_shift=True) (winston_lutz.py:1587-1601): the BBs are moved by the computed shift and everything is redone
    for name in VSHIFT_SETS:
        frames, ps, sid, axes = set_frames(name)
        ref = reference_wlset(frames, ps, sid, axes, apply_virtual_shift=True)
        for k, v in ref.items():
            store[f"{name}_vshift/{k}"] = v
        print(name, "virtual shift", ref["bb_shift_vector"].tolist(), float(ref["max_2d_cax_to_bb_mm"]))
    np.savez_compressed("tests/golden/wlset_golden.npz", **store)


if __name__ == "__main__":
    sys.exit(main())
