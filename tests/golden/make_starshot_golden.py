"""Generate tests/golden/starshot_golden.npz by running the UNMODIFIED reference
(/root/reference, stub-imported) on the seeded synthetic cases of starshot_cases.py.

Run here (the container that has /root/reference):  python -m tests.golden.make_starshot_golden
"""
from __future__ import annotations

import hashlib
import sys
import warnings

import numpy as np

from tests.golden.refrun import reference_starshot
from tests.golden.starshot_cases import CASES, case_frame

KEYS = ["iterations", "radius_px", "profile_len", "peak_idx", "peak_xy", "wobble_center", "wobble_radius_px", "wobble_radius_mm",
        "angles", "n_lines", "passed"]


def main():
    store = {}
    warnings.simplefilter("ignore")
    for name in CASES:
        a, ps, sid, ak = case_frame(name)
        store[f"{name}/input_sha1"] = np.frombuffer(hashlib.sha1(a.tobytes()).digest(), dtype=np.uint8)
        try:
            ref = reference_starshot(a, ps, sid, ak)
        except RuntimeError as e:
            store[f"{name}/raises"] = np.array(str(e)[:60])
            print(name, "raises", e)
            continue
        for k in KEYS:
            store[f"{name}/{k}"] = np.asarray(ref[k])
        print(name, "ok iterations", ref["iterations"], "lines", ref["n_lines"], "wobble", ref["wobble_center"], ref["wobble_radius_mm"])
    np.savez_compressed("tests/golden/starshot_golden.npz", **store)


if __name__ == "__main__":
    sys.exit(main())
