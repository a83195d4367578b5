"""CPU: the Winston-Lutz ORACLE (oracle/wl_oracle.py + the restated skimage functions) against the reference's own synthetic
expectations (tests_basic/test_winstonlutz.py:1244-1520).  The goldens of tests/golden/wl_golden.npz prove CUDA == oracle ==
reference-control-flow-with-restated-skimage; this test ties that chain to numbers the reference itself asserts (with real
skimage): known BB offsets must come back as shift vector / measured position / distance statistics within its tolerances."""
import warnings

import numpy as np
import pytest

from oracle import wl_oracle
from pylinac_b200 import _native as nat
from pylinac_b200 import winston_lutz as wl
from tests.golden.wl_synthetic_classes import CLASSES, _set


def _rows(frames, dpmm):
    rows = np.zeros(len(frames), nat.WL_RESULT_DTYPE)
    for k, f in enumerate(frames):
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            o = wl_oracle.wl2d_analyze(f, dpmm, bb_size_mm=5)
        bb, fld, ep = o["bb"], o["field_cax"], o["epid"]
        rows["bb_x"][k], rows["bb_y"][k] = bb
        rows["field_x"][k], rows["field_y"][k] = fld
        rows["epid_x"][k], rows["epid_y"][k] = ep
        rows["cax2bb_x"][k], rows["cax2bb_y"][k] = (bb[0] - fld[0]) / dpmm, (bb[1] - fld[1]) / dpmm
        rows["cax2bb_distance"][k] = np.sqrt((fld[0] - bb[0]) ** 2 + (fld[1] - bb[1]) ** 2 + 0.0) / dpmm
        rows["cax2epid_x"][k], rows["cax2epid_y"][k] = (ep[0] - fld[0]) / dpmm, (ep[1] - fld[1]) / dpmm
        rows["cax2epid_distance"][k] = np.sqrt((fld[0] - ep[0]) ** 2 + (fld[1] - ep[1]) ** 2 + 0.0) / dpmm
    return rows


@pytest.mark.parametrize("name", ["Synthetic1mmLeft", "Synthetic1mmUp", "Synthetic1mmIn1mmLeft", "Synthetic2mmRight1mmDown"])
def test_oracle_meets_reference_synthetic_expectations(name):
    left, up, inn, axes, exp = CLASSES[name]
    frames, dpmm = _set(left, up, inn, axes)
    rows = _rows(frames, dpmm)
    st = wl.WinstonLutz.__new__(wl.WinstonLutz)
    st._setup(np.zeros((len(rows), 4, 4), np.uint16), [tuple(float(v) for v in a) for a in axes], dpmm)
    refs = dict(snap_tolerance=3, gantry_reference=0, collimator_reference=0, couch_reference=0)
    st.images = [wl._SetImage(wl.WLFrameResult(rows[k]), dpmm, *st._axes[k], refs) for k in range(len(rows))]
    st._is_analyzed = True
    sv = st.bb_shift_vector
    assert abs(sv.x - left) < 0.05 and abs(sv.y - (-inn)) < 0.05 and abs(sv.z - (-up)) < 0.05, (sv.x, sv.y, sv.z)
    mp = st.measured_bb_position
    assert abs(mp.x - (-left)) < 0.03 and abs(mp.y - inn) < 0.03 and abs(mp.z - up) < 0.03
    assert abs(st.cax2bb_distance("max") - exp["bb_max"]) < 0.15
    assert abs(st.cax2bb_distance("median") - exp["bb_median"]) < 0.1
    assert abs(st.cax2bb_distance("mean") - exp["bb_mean"]) < 0.1
    assert abs(st.cax2epid_distance("max") - exp["epid_max"]) < 0.1
    if "couch_iso" in exp:
        assert abs(st.couch_iso_size - exp["couch_iso"]) < 0.15
