"""Generate tests/golden/gamma_golden.npz: BaseImage.gamma of the UNMODIFIED reference (core/image.py:928-1017, stub-imported) on
seeded image pairs.  Run here:  python -m tests.golden.make_gamma_golden"""
from __future__ import annotations

import sys
import warnings

import numpy as np

from tests.golden.gamma_cases import CASES, case_images


def main():
    from oracle.refstub import import_reference

    import_reference()
    from pylinac.core import image as rimage

    store = {}
    warnings.simplefilter("ignore")
    for name in CASES:
        a, b, dpi, kw = case_images(name)
        g = rimage.ArrayImage(np.array(a), dpi=dpi).gamma(rimage.ArrayImage(np.array(b), dpi=dpi), **kw)
        store[name] = np.asarray(g)
        print(name, g.dtype, float(np.nanmax(g)), int(np.isnan(g).sum()))
    np.savez_compressed("tests/golden/gamma_golden.npz", **store)


if __name__ == "__main__":
    sys.exit(main())
