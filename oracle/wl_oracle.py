"""TEST INFRASTRUCTURE ONLY -- never imported by the product (pylinac_b200/).

CPU restatement (numpy / scipy) of the per-image (2-D) Winston-Lutz path, the parity oracle of the CUDA pipeline:

    WLBaseImage.analyze / _clean_edges / find_field_centroids / find_bb_centroids / find_bb_matches   winston_lutz.py:668-829, 1109-1133
    WinstonLutz2D.analyze / cax2bb_* / cax2epid_*                                                       winston_lutz.py:1137-1231
    SizedDiskLocator.calculate (from_center_physical)                                                   metrics/image.py:526-612, 661-667
    find_features / deduplicate_points_and_boundaries                                                   metrics/utils.py:14-37, 66-190
    predicates is_right_size_bb / is_round / is_right_circumference / is_symmetric / is_solid           metrics/features.py:7-68
    stretch / invert / ground / normalize                                                               core/array_utils.py:64-168

skimage (label, clear_border, regionprops) is not installed here: the restatement of oracle/skimage_shim.py is used, both by this
oracle and -- installed into the stub package -- by the UNMODIFIED reference that generated tests/golden/wl_golden.npz
(tests/golden/make_wl_golden.py).  The oracle is therefore pinned to the reference's own control flow and float arithmetic; the
skimage boundary itself (perimeter, convex area) is UNPINNED and says so here and in DESIGN.md.
"""
from __future__ import annotations

import math

import numpy as np
from scipy import ndimage

from oracle import skimage_shim as sk


def _invert(a):
    return -a + a.max() + a.min()


def _stretch01(a):
    """core/array_utils.py:142-168 with min=0, max=1"""
    g = a - a.min()
    n = g / g.max()
    s = n * (1 - 0)
    return s - s.min() + 0


def find_bbs(sample, top, left, dpmm, radius_mm, tol_mm, max_number=1, min_separation_mm=5):
    """find_features (metrics/utils.py:66-190) with the WinstonLutz2D detection conditions."""
    s = _stretch01(sample)
    imin, imax = s.min(), s.max()
    step = (imax - imin) / 50
    cutoff = imin + step
    total = []
    passes = 0
    while cutoff <= imax and len(total) < max_number:
        passes += 1
        lab = sk.clear_border(sk.label(s > cutoff))
        keep = []
        for rg in sk.regionprops(lab, s):
            filled = rg.area_filled
            bb_area = filled / dpmm**2
            if not (max(np.pi * (radius_mm - tol_mm) ** 2, 2) < bb_area < np.pi * (radius_mm + tol_mm) ** 2):
                continue
            ratio = filled / rg.area_bbox
            if not (np.pi / 4 * 1.2 > ratio > np.pi / 4 * 0.8):
                continue
            per = rg.perimeter / dpmm
            if not (2 * np.pi * (radius_mm + tol_mm) > per > 2 * np.pi * (radius_mm - tol_mm)):
                continue
            y0, x0, y1, x1 = rg.bbox
            yy, xx = abs(y1 - y0), abs(x1 - x0)
            if xx > max(yy * 1.05, yy + 3) or xx < min(yy * 0.95, yy - 3):
                continue
            if not rg.solidity > 0.9:
                continue
            keep.append(rg)
        if keep:
            originals = list(total)
            for rg in keep:
                wc = rg.centroid_weighted
                p = (wc[1], wc[0])
                if all(math.dist(p, o) >= min_separation_mm * dpmm for o in originals):
                    total.append(p)
        cutoff += step
    return [(x + left, y + top) for x, y in total], passes


def wl2d_analyze(frame, dpmm, *, bb_size_mm=5, low_density_bb=False, open_field=False, bb_proximity_mm=20):
    a = np.array(frame)
    # check_inversion_by_histogram((0.01, 50, 99.99))
    p_low, p_mid, p_high = (np.percentile(a, q) for q in (0.01, 50, 99.99))
    inverted = bool(abs(p_mid - p_low) > abs(p_mid - p_high))
    if inverted:
        a = _invert(a)
    # _clean_edges(window_size=2)
    safety = min(a.shape) / 10
    crops = 0
    while safety > 0:
        near_min, near_max = np.percentile(a, [5, 99.5])
        rng = near_max - near_min
        edge = np.concatenate((a[:2, :].flatten(), a[:, :2].flatten(), a[-2:, :].flatten(), a[:, -2:].flatten()))
        if not (edge.min() < (near_min - rng / 10) or edge.max() > (near_max + rng / 10)):
            break
        a = a[2:-2, 2:-2]
        crops += 1
        safety -= 1
    a = a - a.min()                   # ground
    a = a / a.max()                   # normalize
    H, W = a.shape
    epid = (W / 2 - 0.5, H / 2 - 0.5)
    # find_field_centroids
    if open_field:
        field = epid
    else:
        mn, mx = np.percentile(a, [5, 99.9])
        binary = np.where(a >= (mx - mn) / 2 + mn, 1, 0)
        filled = ndimage.binary_fill_holes(binary)
        com = ndimage.center_of_mass(filled)
        field = (com[-1], com[0])
    # find_bb_centroids: SizedDiskLocator.from_center_physical((0, 0), window 40 + bb, radius bb / 2, ...)
    tol = float(np.interp(bb_size_mm, (1.5, 30), (2, 4)))
    win = (40 + bb_size_mm) * dpmm
    ex, ey = W / 2, H / 2
    left = max(math.floor(ex - win / 2), 0)
    right = math.ceil(ex + win / 2)
    top = max(math.floor(ey - win / 2), 0)
    bottom = math.ceil(ey + win / 2)
    sample = a[top:bottom, left:right]
    if not low_density_bb:
        sample = _invert(sample)
    pts, passes = find_bbs(sample, top, left, dpmm, bb_size_mm / 2, tol)
    if len(pts) < 1:
        raise ValueError("Couldn't find the minimum number of disks in the image.")
    # find_bb_matches / find_field_matches: nearest detected point to the nominal (ISO: the EPID centre) within the proximity
    def match(points):
        d = [math.dist(epid, p) for p in points]
        k = int(np.argmin(d))
        return points[k] if d[k] < bb_proximity_mm * dpmm else None

    fm, bm = match([field]), match(pts)
    if (fm is None) != (bm is None):
        raise ValueError("The number of detected fields and BBs do not match")
    if fm is None:
        raise ValueError("No fields were detected")
    return {
        "inverted": inverted, "crops": crops, "shape": np.array([H, W]), "field_cax": np.array(fm), "bb": np.array(bm),
        "epid": np.array(epid), "threshold_passes": passes,
        "cax2bb_vector": np.array([(bm[0] - fm[0]) / dpmm, (bm[1] - fm[1]) / dpmm]),
        "cax2bb_distance": math.dist(fm, bm) / dpmm,
        "cax2epid_vector": np.array([(epid[0] - fm[0]) / dpmm, (epid[1] - fm[1]) / dpmm]),
        "cax2epid_distance": math.dist(fm, epid) / dpmm,
    }
