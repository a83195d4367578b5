"""Generate tests/golden/profile_golden.npz: the UNMODIFIED reference SingleProfile (stub-imported) on seeded 1-D profiles.

Run here (the container that has /root/reference):  python -m tests.golden.make_profile_golden
"""
from __future__ import annotations

import sys
import warnings

import numpy as np

from tests.golden.profile_cases import CASES, case_profile

FD_KEYS = ["width (exact)", "beam center index (exact)", "beam center value (@rounded)", "cax index (exact)", "cax value (@rounded)",
           "left index (exact)", "right index (exact)", "left value (@rounded)", "right value (@rounded)", "left slope", "left intercept",
           "right slope", "right intercept", "left inner index (exact)", "right inner index (exact)"]


def main():
    from oracle.refstub import import_reference

    import_reference()
    from pylinac.core import profile as rp

    store = {}
    warnings.simplefilter("ignore")
    for name in CASES:
        vals, kw, q = case_profile(name)
        rkw = dict(kw)
        for k, enum in (("interpolation", rp.Interpolation), ("normalization_method", rp.Normalization), ("edge_detection_method", rp.Edge),
                        ("centering", rp.Centering)):
            if k in rkw:
                rkw[k] = enum(rkw[k])
        sp = rp.SingleProfile(vals, **rkw)
        store[f"{name}/values"] = np.asarray(sp.values, dtype=float)
        store[f"{name}/x_indices"] = np.asarray(sp.x_indices, dtype=float)
        gc, bc = sp.geometric_center(), sp.beam_center()
        store[f"{name}/geometric_center"] = np.array([gc["index (exact)"], gc["value (exact)"]])
        store[f"{name}/beam_center"] = np.array([bc["index (exact)"], bc["value (@rounded)"]])
        fw = sp.fwxm_data(q["fwxm_x"])
        store[f"{name}/fwxm"] = np.array([fw["left index (exact)"], fw["right index (exact)"], fw["center value (@rounded)"],
                                          fw["left value (@rounded)"], fw["right value (@rounded)"]])
        pen = sp.penumbra(*q["penumbra"])
        lo, up = q["penumbra"]
        store[f"{name}/penumbra"] = np.array([pen[f"left {lo}% index (exact)"], pen[f"left {up}% index (exact)"],
                                              pen[f"right {lo}% index (exact)"], pen[f"right {up}% index (exact)"]])
        if rkw.get("edge_detection_method", rp.Edge.FWHM) == rp.Edge.INFLECTION_HILL:
            inf = sp.inflection_data()
            store[f"{name}/hill"] = np.array([inf["left index (exact)"], inf["right index (exact)"], inf["left value (@exact)"],
                                              inf["right value (@exact)"]])
            store[f"{name}/hill_params"] = np.array([inf["left Hill params"], inf["right Hill params"]])
            store[f"{name}/hill_gradients"] = np.array([pen["left gradient (exact)"], pen["right gradient (exact)"]])
        elif rkw.get("edge_detection_method", rp.Edge.FWHM) != rp.Edge.FWHM:
            inf = sp.inflection_data()
            store[f"{name}/inflection"] = np.array([inf["left index (exact)"], inf["right index (exact)"], inf["left value (@exact)"],
                                                    inf["right value (@exact)"], inf["left value (@rounded)"], inf["right value (@rounded)"]])
        fd = sp.field_data(q["in_field_ratio"], q["slope_exclusion_ratio"])
        store[f"{name}/field_data"] = np.array([float(fd[k]) for k in FD_KEYS])
        store[f"{name}/field_values"] = np.asarray(fd["field values"], dtype=float)
        store[f"{name}/top_params"] = np.asarray(fd["top params"], dtype=float)
        print(name, "ok", len(sp.values), fw["left index (exact)"], fd["width (exact)"])
    np.savez_compressed("tests/golden/profile_golden.npz", **store)


if __name__ == "__main__":
    sys.exit(main())
