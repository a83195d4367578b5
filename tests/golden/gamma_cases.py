"""Seeded image pairs for BaseImage.gamma (reference + comparison, same shape / dpi)."""
from __future__ import annotations

import numpy as np

CASES = {
    "gauss_shifted": dict(doseTA=2, distTA=1, threshold=0.1),
    "defaults": {},
    "no_ground": dict(doseTA=3, distTA=2, threshold=0.05, ground=False),
    "inverted_pair": dict(doseTA=1, distTA=1, threshold=0.2),
}


def case_images(name):
    """-> (reference uint16, comparison uint16, dpi, gamma kwargs)"""
    seed = sorted(CASES).index(name)
    rng = np.random.default_rng(40 + seed)
    h, w = 96 + 8 * seed, 120 - 4 * seed
    yy, xx = np.mgrid[0:h, 0:w]
    cy, cx = h / 2 + rng.uniform(-3, 3), w / 2 + rng.uniform(-3, 3)
    base = 30000 * np.exp(-(((yy - cy) / (h / 5)) ** 2 + ((xx - cx) / (w / 5)) ** 2))
    a = (base + rng.normal(0, 80, (h, w)) + 700).clip(0, 65535).astype(np.uint16)
    shifted = 30000 * np.exp(-(((yy - cy - 0.8) / (h / 5)) ** 2 + ((xx - cx + 1.1) / (w / 5)) ** 2))
    b = (1.02 * shifted + rng.normal(0, 80, (h, w)) + 650).clip(0, 65535).astype(np.uint16)
    if name == "inverted_pair":
        a = (int(a.max()) + int(a.min()) - a.astype(np.int64)).astype(np.uint16)
        b = (int(b.max()) + int(b.min()) - b.astype(np.int64)).astype(np.uint16)
    return a, b, 72.0, dict(CASES[name])
