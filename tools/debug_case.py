"""Print the GPU summary of one golden PF case (debug aid): python tools/debug_case.py as1200"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from tests.golden.pf_cases import case_frame
from pylinac_b200 import picketfence as pf

GOLD = np.load("tests/golden/pf_golden.npz")
for name in sys.argv[1:]:
    a, ps, sid, ck, ak = case_frame(name)
    dpmm = (1 / ps) * sid / 1000.0
    ck = dict(ck)
    if ck.get("mlc") == "HD":
        ck["mlc"] = pf.MLC.HD_MILLENNIUM
    r = pf.analyze_batch(a[None], dpmm, **ck, **ak)[0]
    s = r.s
    print(name, "shape", a.shape, "dpmm", dpmm)
    for k in ("status", "orientation", "noise_median_passes", "corner_inverted", "height", "width", "n_pickets", "n_meas",
              "n_leaves_removed", "picket_spacing_px"):
        print("  ", k, s[k])
    print("   picket_idx", s["picket_idx"][: max(int(s["n_pickets"]), 0)])
    if f"{name}/picket_idx" in GOLD:
        print("   gold picket_idx", GOLD[f"{name}/picket_idx"], "n_meas", GOLD[f"{name}/n_meas"], "spacing", GOLD[f"{name}/picket_spacing"])
