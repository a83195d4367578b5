"""Generate tests/golden/field_golden.npz by running the UNMODIFIED reference
(/root/reference, stub-imported) on the seeded synthetic cases of field_cases.py.

Run here (the container that has /root/reference):  python -m tests.golden.make_field_golden
"""
from __future__ import annotations

import hashlib
import sys
import warnings

import numpy as np

from tests.golden.field_cases import CASES, EXT_CASES, case_frame
from tests.golden.refrun import reference_field


def main():
    store = {}
    warnings.simplefilter("ignore")
    for name in CASES + EXT_CASES:
        a, ps, sid, ak = case_frame(name)
        store[f"{name}/input_sha1"] = np.frombuffer(hashlib.sha1(a.tobytes()).digest(), dtype=np.uint8)
        ref = reference_field(a, ps, sid, ak)
        for k, v in ref.items():
            store[f"{name}/{k}"] = np.asarray(v)
        print(name, "ok", {k: (float(v) if np.ndim(v) == 0 else v.tolist()) for k, v in ref.items() if "flatness" in k or "symmetry" in k},
              ref["profile_len"].tolist())
    np.savez_compressed("tests/golden/field_golden.npz", **store)


if __name__ == "__main__":
    sys.exit(main())
