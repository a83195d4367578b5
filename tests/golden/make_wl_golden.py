"""Generate tests/golden/wl_golden.npz by running the UNMODIFIED reference WinstonLutz2D (stub-imported; the absent skimage
functions are served by oracle/skimage_shim.py) on the seeded synthetic cases of wl_cases.py.

Run here (the container that has /root/reference):  python -m tests.golden.make_wl_golden
"""
from __future__ import annotations

import hashlib
import sys
import warnings

import numpy as np

from tests.golden.refrun import reference_wl2d
from tests.golden.wl_cases import CASES, case_frame


def main():
    store = {}
    warnings.simplefilter("ignore")
    for name in CASES:
        a, ps, sid, g, c, p, ak = case_frame(name)
        store[f"{name}/input_sha1"] = np.frombuffer(hashlib.sha1(a.tobytes()).digest(), dtype=np.uint8)
        try:
            ref = reference_wl2d(a, ps, sid, g, c, p, ak)
        except ValueError as e:
            store[f"{name}/raises"] = np.array(str(e)[:60])
            print(name, "raises", e)
            continue
        for k, v in ref.items():
            store[f"{name}/{k}"] = np.asarray(v)
        print(name, "ok", ref["shape"].tolist(), ref["bb"], ref["field_cax"], round(ref["cax2bb_distance"], 4), ref["variable_axis"])
    np.savez_compressed("tests/golden/wl_golden.npz", **store)


if __name__ == "__main__":
    sys.exit(main())
