"""Generate tests/golden/pf_golden.npz by running the UNMODIFIED reference
(/root/reference, stub-imported) on the seeded synthetic cases of pf_cases.py.

Run here (the container that has /root/reference):  python -m tests.golden.make_pf_golden
The inputs are regenerated from their seeds at test time (oracle/synth.py), so only the
reference's outputs are committed.
"""
from __future__ import annotations

import hashlib
import sys
import warnings

import numpy as np

from tests.golden.pf_cases import CASES, case_frame
from tests.golden.refrun import reference_pf

KEYS = ["orientation", "n_meas", "meas_leaf", "meas_picket", "meas_position", "meas_error", "meas_width_mm", "picket_idx",
        "picket_spacing", "fits", "percent_passing", "max_error", "abs_median_error", "max_error_picket", "passed",
        "offsets_from_cax_mm", "mean_picket_spacing", "mlc_skew", "number_of_pickets", "picket_widths", "shape"]


def main():
    from oracle.refstub import import_reference

    import_reference()
    from pylinac import picketfence as rpf

    store = {}
    warnings.simplefilter("ignore")
    for name in CASES:
        a, ps, sid, ck, ak = case_frame(name)
        store[f"{name}/input_sha1"] = np.frombuffer(hashlib.sha1(a.tobytes()).digest(), dtype=np.uint8)
        ck = dict(ck)
        if ck.get("mlc") == "HD":
            ck["mlc"] = rpf.MLC.HD_MILLENNIUM
        try:
            ref = reference_pf(a, ps, sid, ck, ak)
        except ValueError as e:
            store[f"{name}/raises"] = np.array(str(e)[:60])
            print(name, "raises", e)
            continue
        for k in KEYS:
            store[f"{name}/{k}"] = np.asarray(ref[k])
        store[f"{name}/max_error_leaf"] = np.array(str(ref["max_error_leaf"]))
        store[f"{name}/failed_leaves"] = np.array([str(x) for x in ref["failed_leaves"]])
        print(name, "ok", ref["number_of_pickets"], ref["n_meas"])
    np.savez_compressed("tests/golden/pf_golden.npz", **store)


if __name__ == "__main__":
    sys.exit(main())
