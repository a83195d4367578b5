"""Generate tests/golden/vmat_golden.npz by running the UNMODIFIED reference DRGS / DRMLC / DRCS (pylinac/vmat.py) and DLG
(pylinac/dlg.py), stub-imported from /root/reference, on the seeded synthetic cases of vmat_cases.py.

Run here (the container that has /root/reference):  python -m tests.golden.make_vmat_golden
"""
from __future__ import annotations

import hashlib
import sys
import warnings

import numpy as np

from tests.golden.refrun import reference_dlg, reference_vmat
from tests.golden.vmat_cases import DLG_CASES, DRCS_CASES, VMAT_CASES, dlg_case, drcs_case, vmat_case


def main():
    store = {}
    warnings.simplefilter("ignore")
    for name in VMAT_CASES + DRCS_CASES:
        klass, a, b, ps, sid, ck, ak = (vmat_case if name in VMAT_CASES else drcs_case)(name)
        store[f"{name}/input_sha1"] = np.frombuffer(hashlib.sha1(a.tobytes() + b.tobytes()).digest(), dtype=np.uint8)
        ref = reference_vmat(klass, a, b, ps, sid, ck, ak)
        for k, v in ref.items():
            store[f"{name}/{k}"] = np.asarray(v)
        print(name, klass, "open_is_first", ref["open_is_first"], "r_dev", np.round(ref["r_dev"], 3), "passed", ref["passed"])
    for name in DLG_CASES:
        args = dlg_case(name)
        store[f"{name}/input_sha1"] = np.frombuffer(hashlib.sha1(args[0].tobytes()).digest(), dtype=np.uint8)
        ref = reference_dlg(*args)
        for k, v in ref.items():
            store[f"{name}/{k}"] = np.asarray(v)
        print(name, "dlg", ref["measured_dlg"])
    np.savez_compressed("tests/golden/vmat_golden.npz", **store)


if __name__ == "__main__":
    sys.exit(main())
