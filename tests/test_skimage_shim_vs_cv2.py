"""Differential test of the restated scikit-image functions (oracle/skimage_shim.py: the only unpinned boundary of the Winston-Lutz
path) against INDEPENDENT implementations from OpenCV on random blobs:

    label(connectivity=1)            vs cv2.connectedComponents(connectivity=4)   (same partition, raster-order numbering)
    clear_border                     vs components whose cv2 bounding box touches the frame
    regionprops area / bbox / centroid  vs cv2.connectedComponentsWithStats
    area_filled                      vs cv2.floodFill of the background from a padded corner
    area_convex (solidity)           vs pixel centres inside cv2.convexHull of the same diamond offsets (cv2.pointPolygonTest >= 0)

``perimeter`` has no OpenCV counterpart (cv2.arcLength measures a different contour): it is checked on shapes whose value under the
published definition is known in closed form.  This narrows, but does not close, the skimage boundary: the DEFINITIONS (diamond
offsets for the hull, the 4-neighbourhood border weights of the perimeter) remain restated from the published algorithms."""
import numpy as np
import pytest

cv2 = pytest.importorskip("cv2")

from oracle import skimage_shim as sk


def _blobs(seed, shape=(90, 110)):
    rng = np.random.default_rng(seed)
    img = np.zeros(shape, np.uint8)
    for _ in range(rng.integers(3, 9)):
        c = (int(rng.integers(0, shape[1])), int(rng.integers(0, shape[0])))
        axes = (int(rng.integers(2, 14)), int(rng.integers(2, 14)))
        cv2.ellipse(img, c, axes, float(rng.uniform(0, 180)), 0, 360, 1, -1)
    noise = rng.random(shape) < 0.03
    img[noise] ^= 1
    for _ in range(2):   # a ring (a region with a hole) and a concave "L"
        c = (int(rng.integers(15, shape[1] - 15)), int(rng.integers(15, shape[0] - 15)))
        cv2.circle(img, c, 8, 1, 2)
    return img


@pytest.mark.parametrize("seed", range(12))
def test_label_clear_border_and_region_geometry(seed):
    img = _blobs(seed)
    lab = sk.label(img, connectivity=1)
    n_cv, lab_cv, stats, cents = cv2.connectedComponentsWithStats(img, connectivity=4)
    assert lab.max() == n_cv - 1
    # same partition
    pairs = {(int(a), int(b)) for a, b in zip(lab.ravel(), lab_cv.ravel())}
    assert len(pairs) == n_cv and len({a for a, _ in pairs}) == n_cv
    # raster-order numbering: label k is the k-th component met in a row-major scan
    first = {}
    for idx, v in enumerate(lab.ravel()):
        if v and v not in first:
            first[int(v)] = idx
    assert [k for k, _ in sorted(first.items(), key=lambda kv: kv[1])] == list(range(1, lab.max() + 1))
    # clear_border == drop the components whose bounding box touches the frame
    cleared = sk.clear_border(lab)
    h, w = img.shape
    to_cv = dict(pairs)
    for v in range(1, lab.max() + 1):
        x, y, bw, bh, _ = stats[to_cv[v]]
        touches = x == 0 or y == 0 or x + bw == w or y + bh == h
        assert (v in cleared) != touches, v
    # regionprops geometry
    sample = np.random.default_rng(seed).random(img.shape)
    for rg in sk.regionprops(cleared, sample):
        x, y, bw, bh, area = stats[to_cv[rg.label]]
        assert rg.area == area and tuple(rg.bbox) == (y, x, y + bh, x + bw)
        assert rg.area_bbox == bw * bh
        # filled area: flood the background of the padded region mask from a corner with OpenCV
        mask = (lab[rg.bbox[0]:rg.bbox[2], rg.bbox[1]:rg.bbox[3]] == rg.label).astype(np.uint8)
        pad = np.pad(mask, 1)
        ff = pad.copy()
        cv2.floodFill(ff, None, (0, 0), 2, flags=4)
        assert rg.area_filled == int((ff != 2).sum())
        # weighted centroid against the definition
        rr, cc = np.nonzero(lab == rg.label)
        wts = sample[rr, cc]
        np.testing.assert_allclose(rg.centroid_weighted, ((rr * wts).sum() / wts.sum(), (cc * wts).sum() / wts.sum()), rtol=1e-12)
        # convex area: pixel centres inside (or on) OpenCV's hull of the diamond offsets
        pts = []
        for r, c in zip(*np.nonzero(mask)):
            pts += [(c, r + 0.5), (c, r - 0.5), (c + 0.5, r), (c - 0.5, r)]
        hull = cv2.convexHull((np.array(pts, np.float32) * 2).astype(np.int32))         # doubled coordinates: exact integers
        inside = sum(cv2.pointPolygonTest(hull, (2.0 * c, 2.0 * r), False) >= 0 for r in range(mask.shape[0]) for c in range(mask.shape[1]))
        assert sk.convex_area(mask) == inside
        assert rg.solidity == pytest.approx(rg.area / inside)


def test_perimeter_closed_forms():
    """skimage.measure.perimeter(neighborhood=4): border pixels weighted 1 (straight), sqrt(2) (diagonal step), (1 + sqrt(2)) / 2 (mixed).
    Axis-aligned rectangle n x m: every border pixel is a straight step -> 2 (n - 1) + 2 (m - 1); a one-pixel-wide diagonal line of
    k pixels: only interior diagonal steps -> (k - 2) sqrt(2) + end-point terms of weight 0; single pixel -> 0."""
    rect = np.zeros((20, 30), np.uint8)
    rect[4:14, 5:25] = 1
    assert sk.perimeter(rect) == pytest.approx(2 * 9 + 2 * 19)
    assert sk.perimeter(np.pad(np.ones((1, 1), np.uint8), 2)) == 0.0
    diag = np.eye(12, dtype=np.uint8)
    diag = np.pad(diag, 2)
    assert sk.perimeter(diag) == pytest.approx(10 * np.sqrt(2))
