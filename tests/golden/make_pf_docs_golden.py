"""Generate tests/golden/pf_docs_golden.npz (+ pf_docs_dcm.npz): the UNMODIFIED reference PicketFence on the reference's own
docs fixtures (docs/source/files/*.dcm) with the analyze() arguments of their recipes (docs/source/picketfence.rst:455-730).

Run here (the container that has /root/reference):  python -m tests.golden.make_pf_docs_golden
"""
from __future__ import annotations

import hashlib
import lzma
import os
import sys
import warnings

import numpy as np

from tests.golden import pf_docs_cases as dc
from tests.golden.make_pf_golden import KEYS
from tests.golden.refrun import reference_pf


def main():
    store, frames = {}, {}
    warnings.simplefilter("ignore")
    if os.path.exists(dc.DCM_NPZ):
        os.remove(dc.DCM_NPZ)          # read every file from the reference tree
    for name, ak in dc.DOCS.items():
        a, ps, sid, ak = dc.docs_frame(name)
        store[f"{name}/input_sha1"] = np.frombuffer(hashlib.sha1(a.tobytes()).digest(), dtype=np.uint8)
        ref = reference_pf(a, ps, sid, {}, ak)
        for k in KEYS:
            store[f"{name}/{k}"] = np.asarray(ref[k])
        store[f"{name}/max_error_leaf"] = np.array(str(ref["max_error_leaf"]))
        store[f"{name}/failed_leaves"] = np.array([str(x) for x in ref["failed_leaves"]])
        print(name, "pickets", ref["number_of_pickets"], "kisses", ref["n_meas"], "max err", ref["max_error"], "offsets",
              np.round(ref["offsets_from_cax_mm"], 1).tolist(), "skew", round(float(ref["mlc_skew"]), 3), "passed", ref["passed"])
        frames[name] = np.frombuffer(lzma.compress(dc.docs_dcm_bytes(name), preset=6), dtype=np.uint8)
    np.savez_compressed("tests/golden/pf_docs_golden.npz", **store)
    np.savez(dc.DCM_NPZ, **frames)


if __name__ == "__main__":
    sys.exit(main())
