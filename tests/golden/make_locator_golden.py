"""Generate tests/golden/locator_golden.npz by running the UNMODIFIED reference's whole-frame locators (metrics/image.py:275-354,
727-956; stub-imported, skimage served by oracle/skimage_shim.py) on the seeded cases of locator_cases.py.

Run here (the container that has /root/reference):  python -m tests.golden.make_locator_golden
"""
from __future__ import annotations

import hashlib
import sys
import time
import warnings

import numpy as np

from tests.golden.locator_cases import CASES, case
from tests.golden.refrun import reference_locator


def main():
    store = {}
    warnings.simplefilter("ignore")
    for name in CASES:
        a, ps, sid, spec = case(name)
        store[f"{name}/input_sha1"] = np.frombuffer(hashlib.sha1(a.tobytes()).digest(), dtype=np.uint8)
        t = time.time()
        try:
            pts = reference_locator(a, ps, sid, spec)
            store[f"{name}/raised"] = np.array(0)
        except ValueError as e:
            pts = np.zeros((0, 2))
            store[f"{name}/raised"] = np.array(1)
            print(name, "raised", e)
        store[f"{name}/points"] = pts
        print(name, round(time.time() - t, 1), "s", np.round(pts, 2).tolist())
    np.savez_compressed("tests/golden/locator_golden.npz", **store)


if __name__ == "__main__":
    sys.exit(main())
