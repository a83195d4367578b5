"""Debug helper: raw records of epid_global_locate next to the restated skimage sweep (oracle/skimage_shim.py) for one locator case."""
import sys

import numpy as np

sys.path.insert(0, ".")
from oracle import skimage_shim as sk
from pylinac_b200 import _native as nat
from pylinac_b200.core import image
from pylinac_b200.metrics import image as mi
from tests.golden.locator_cases import case

name = sys.argv[1]
a, ps, sid, spec = case(name)
img = image.ArrayImage(a, dpi=25.4 / ps, sid=sid)
dpmm = img.dpmm
cls = getattr(mi, spec["cls"])
m = cls.from_physical(**spec["kw"]) if spec.get("physical") else cls(**spec["kw"])
m.inject_image(img)
if isinstance(m, mi.GlobalSizedFieldLocator) and not m.is_from_physical:
    m.field_width_mm /= dpmm; m.field_height_mm /= dpmm; m.field_tolerance_mm /= dpmm
p = m._params(dpmm)
regs = nat.global_locate(nat.Context.default(), a, p)[0]
print(name, "device regions:", len(regs))
for r in regs[:40]:
    print("  k", r["threshold_index"], "root", r["label_root"], "bbox", r["bbox"].tolist(), "area", r["area"], "filled", r["area_filled"],
          "perim", round(float(r["perimeter"]), 3), "c", round(float(r["centroid_x"]), 3), round(float(r["centroid_y"]), 3),
          "wc", round(float(r["wcentroid_x"]), 3), round(float(r["wcentroid_y"]), 3))
# restated sweep
if p.mode == 1:
    sample = a
    imin, imax = sample.min(), sample.max()
    step = (imax - imin) / 50
    cutoff = imin + step * 5
    k = 0
    while cutoff <= imax and k < 60:
        b = sk.clear_border(sample > cutoff, buffer_size=3)
        lab = sk.label(b)
        props = sk.regionprops(lab, intensity_image=sample)
        if k < 3 or k % 10 == 0:
            print("  shim k", k, "cutoff", cutoff, "regions", len(props), [(r.bbox, r.area, r.area_filled, round(r.perimeter, 3)) for r in props[:6]])
        cutoff += step
        k += 1
else:
    b = (-a + a.max() + a.min()) if p.invert else a
    g = b - b.min()
    sample = (g / g.max()) * 1
    sample = sample - sample.min() + 0
    step = 1.0 / 50
    cutoff = 0.0 + step
    k = 0
    while cutoff <= 1.0 and k < 60:
        lab = sk.clear_border(sk.label(sample > cutoff, connectivity=1))
        props = sk.regionprops(lab, intensity_image=sample)
        ok = [r for r in props if (3.141592653589793 / 4 * 1.2 > r.filled_area / r.bbox_area > 3.141592653589793 / 4 * 0.8)]
        if k < 3 or k % 8 == 0:
            print("  shim k", k, "cutoff", cutoff, "regions", len(props), "round", [(r.bbox, r.area, r.area_filled, round(r.perimeter, 3)) for r in ok[:6]])
        cutoff += step
        k += 1
