"""Generate tests/golden/profile_ext_golden.npz: the UNMODIFIED reference's ProfileBase family (core/profile.py:195-1116) on seeded
profiles -- ``as_resampled`` (plain and physical, grid on / off, order 1 / 3), ``resample_to``, ``as_simple_profile`` and the
Hill-fit edges of ``HillProfile`` / ``HillProfilePhysical``.

Run here:  python -m tests.golden.make_profile_ext_golden
"""
from __future__ import annotations

import warnings

import numpy as np

from tests.golden.profile_cases import _field

CASES = {
    # name: (field args, class, ctor kwargs)
    "fwxm": ((300, 80.3, 220.6, 5.0, 11), "FWXMProfile", {"fwxm_height": 40}),
    "infl": ((400, 110.7, 290.1, 6.0, 12), "InflectionDerivativeProfile", {"edge_smoothing_ratio": 0.005}),
    "hill": ((400, 100.2, 310.4, 7.0, 13), "HillProfile", {"hill_window_ratio": 0.12}),
    "hill_fff": ((500, 130.5, 380.2, 8.0, 14, -1.2), "HillProfile", {}),
}
PHYS = {
    "fwxm_phys": ((320, 90.3, 240.6, 5.0, 21), "FWXMProfilePhysical", {"dpmm": 2.56}),
    "infl_phys": ((300, 70.7, 230.1, 6.0, 22), "InflectionDerivativeProfilePhysical", {"dpmm": 3.0}),
    "hill_phys": ((360, 100.2, 270.4, 7.0, 23), "HillProfilePhysical", {"dpmm": 2.976}),
}


def values_of(args):
    return _field(*args[:5], horns=args[5] if len(args) > 5 else 0.1)


def main():
    from oracle.refstub import import_reference

    import_reference()
    from pylinac.core import profile as rp

    warnings.simplefilter("ignore")
    store = {}

    def dump(tag, prof):
        store[f"{tag}/values"] = np.asarray(prof.values, dtype=float)
        store[f"{tag}/x_values"] = np.asarray(prof.x_values, dtype=float)
        store[f"{tag}/edges"] = np.array([prof.field_edge_idx("left"), prof.field_edge_idx("right"), prof.center_idx, prof.field_width_px])

    for name, (args, cls, kw) in CASES.items():
        prof = getattr(rp, cls)(values_of(args), **kw)
        dump(name, prof)
        dump(f"{name}/res10", prof.as_resampled())
        dump(f"{name}/res2.5_o1", prof.as_resampled(interpolation_factor=2.5, order=1))
        dump(f"{name}/res0.5", prof.as_resampled(interpolation_factor=0.5))
    for name, (args, cls, kw) in PHYS.items():
        prof = getattr(rp, cls)(values_of(args), **kw)
        dump(name, prof)
        store[f"{name}/physical_x"] = np.asarray(prof.physical_x_values, dtype=float)
        store[f"{name}/width_mm"] = np.array(prof.field_width_mm)
        r = prof.as_resampled()
        dump(f"{name}/res", r)
        store[f"{name}/res/dpmm"] = np.array(r.dpmm)
        store[f"{name}/res/width_mm"] = np.array(r.field_width_mm)
        r2 = prof.as_resampled(interpolation_resolution_mm=0.25, order=1, grid=False)
        dump(f"{name}/res_nogrid", r2)
        sp = prof.as_simple_profile()
        dump(f"{name}/simple", sp)
    # resample_to: an EPID-like physical profile onto a sparse "ion chamber" profile and back
    epid = rp.FWXMProfilePhysical(values_of(PHYS["fwxm_phys"][0]), dpmm=2.56)
    ic_x = np.linspace(10.0, 110.0, 41)
    ic = rp.FWXMProfile(np.interp(ic_x, epid.physical_x_values, epid.values) * 1.01, x_values=ic_x)
    out = epid.resample_to(ic)
    store["resample_to/values"], store["resample_to/x_values"] = np.asarray(out.values, float), np.asarray(out.x_values, float)
    store["resample_to/type"] = np.array(type(out).__name__)
    np.savez_compressed("tests/golden/profile_ext_golden.npz", **store)
    for k in ("hill/edges", "hill_fff/edges", "hill_phys/edges", "hill/res10/edges", "fwxm_phys/res/edges"):
        print(k, store[k])


if __name__ == "__main__":
    main()
