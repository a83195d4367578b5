"""TEST INFRASTRUCTURE ONLY -- never imported by the product (pylinac_b200/).

CPU restatement of the edge / line operators behind ``JawOrthogonality.analyze`` (pylinac/contrib/orthogonality.py:7-39, 29-86):

    skimage.feature.canny(image)                     sigma = 1, low / high thresholds 0.1 / 0.2 (float image), mode 'constant'
    skimage.transform.hough_line(edges, theta)       classic straight-line Hough transform
    skimage.transform.hough_line_peaks(h, theta, d)  min_distance = 9, min_angle = 10, threshold = 0.5 * max

**PARITY UNPINNED.**  scikit-image is not installed in this container and the reference holds no golden vectors for this path (its
only test downloads images and checks that ``analyze()`` runs, tests_basic/contrib/test_orthogonality.py).  The three functions are
restated from the published skimage algorithms (Canny 1986 with bilinear non-maximum suppression along the gradient, hysteresis by
8-connected components; accumulator index round(x cos t + y sin t) + offset; prominent peaks = maximum filter, threshold, components
sorted by intensity with neighbourhood suppression) on top of scipy.ndimage.  The CUDA path (csrc/edges.cu) is tested against THIS
restatement bit for bit on the edge map and the accumulator, and against geometric ground truth (rotated synthetic fields, OpenCV's
Hough transform) for the angles; that is all the pinning available.
"""
from __future__ import annotations

import numpy as np
from scipy import ndimage as ndi


def _gaussian(image, sigma):
    return ndi.gaussian_filter(image, sigma, mode="constant", cval=0.0, truncate=4.0)


def canny(image, sigma=1.0, low_threshold=0.1, high_threshold=0.2):
    image = np.asarray(image, dtype=np.float64)
    rows, cols = image.shape
    eroded = np.ones(image.shape, bool)
    eroded[:1, :] = eroded[-1:, :] = eroded[:, :1] = eroded[:, -1:] = False
    bleed_over = _gaussian(np.ones(image.shape), sigma) + np.finfo(np.float64).eps
    smoothed = _gaussian(image, sigma) / bleed_over
    jsobel = ndi.sobel(smoothed, axis=1)
    isobel = ndi.sobel(smoothed, axis=0)
    magnitude = isobel * isobel
    magnitude += jsobel * jsobel
    np.sqrt(magnitude, out=magnitude)
    low_masked = nonmaximum_suppression_bilinear(isobel, jsobel, magnitude, eroded, low_threshold)
    low_mask = low_masked > 0
    labels, count = ndi.label(low_mask, np.ones((3, 3), bool))
    if count == 0:
        return low_mask
    high_mask = low_mask & (low_masked >= high_threshold)
    good = np.zeros(count + 1, bool)
    good[np.unique(labels[high_mask])] = True
    return good[labels]


def nonmaximum_suppression_bilinear(isobel, jsobel, magnitude, eroded_mask, low_threshold):
    """skimage/feature/_canny_cy.pyx, vectorised: a pixel survives when its magnitude is >= the magnitude interpolated one step along
    the gradient on both sides"""
    out = np.zeros(magnitude.shape)
    m = magnitude[1:-1, 1:-1]
    i_ = isobel[1:-1, 1:-1]
    j_ = jsobel[1:-1, 1:-1]
    ok = eroded_mask[1:-1, 1:-1] & (m >= low_threshold)
    is_down, is_up, is_left, is_right = i_ <= 0, i_ >= 0, j_ <= 0, j_ >= 0
    cond1 = (is_up & is_right) | (is_down & is_left)
    cond2 = (is_down & is_right) | (is_up & is_left)
    ai, aj = np.abs(i_), np.abs(j_)
    g1 = ai > aj
    with np.errstate(divide="ignore", invalid="ignore"):
        w = np.where(g1, aj / ai, ai / aj)
    M = magnitude

    def sh(dy, dx):
        return M[1 + dy : M.shape[0] - 1 + dy, 1 + dx : M.shape[1] - 1 + dx]

    # (neigh1_1, neigh1_2, neigh2_1, neigh2_2) per case
    cases = {
        (True, True): (sh(1, 0), sh(1, 1), sh(-1, 0), sh(-1, -1)),      # cond1, |i| > |j|
        (True, False): (sh(0, 1), sh(1, 1), sh(0, -1), sh(-1, -1)),     # cond1, |i| <= |j|
        (False, True): (sh(-1, 0), sh(-1, 1), sh(1, 0), sh(1, -1)),     # cond2, |i| > |j|
        (False, False): (sh(0, 1), sh(-1, 1), sh(0, -1), sh(1, -1)),    # cond2, |i| <= |j|
    }
    res = np.zeros(m.shape)
    for (c1, grad1), (n11, n12, n21, n22) in cases.items():
        sel = ok & (cond1 if c1 else (cond2 & ~cond1)) & (g1 if grad1 else ~g1)
        with np.errstate(invalid="ignore"):
            c_plus = (n12 * w + n11 * (1.0 - w)) <= m
            c_minus = (n22 * w + n21 * (1.0 - w)) <= m
        keep = sel & c_plus & c_minus
        res[keep] = m[keep]
    out[1:-1, 1:-1] = res
    return out


def hough_line(img, theta):
    img = np.asarray(img)
    ctheta, stheta = np.cos(theta), np.sin(theta)
    offset = int(np.ceil(np.sqrt(img.shape[0] ** 2 + img.shape[1] ** 2)))
    max_distance = 2 * offset + 1
    accum = np.zeros((max_distance, len(theta)), dtype=np.uint64)
    bins = np.linspace(-offset, offset, max_distance)
    ys, xs = np.nonzero(img)
    for j in range(len(theta)):
        r = ctheta[j] * xs + stheta[j] * ys
        idx = np.where(r > 0.0, r + 0.5, r - 0.5).astype(np.int64) + offset      # skimage's round(): half away from zero by truncation
        accum[:, j] = np.bincount(idx, minlength=max_distance).astype(np.uint64)
    return accum, theta, bins


def prominent_peaks(image, min_xdistance, min_ydistance, threshold=None, num_peaks=np.inf):
    img = image.astype(np.float64).copy()
    rows, cols = img.shape
    if threshold is None:
        threshold = 0.5 * np.max(img)
    img_max = ndi.maximum_filter1d(img, size=2 * min_ydistance + 1, axis=0, mode="constant", cval=0)
    img_max = ndi.maximum_filter1d(img_max, size=2 * min_xdistance + 1, axis=1, mode="constant", cval=0)
    img *= img == img_max
    img_t = img > threshold
    lab, cnt = ndi.label(img_t, np.ones((3, 3), bool))
    props = []
    for k, sl in enumerate(ndi.find_objects(lab)):
        sel = lab[sl] == k + 1
        rr, cc = np.nonzero(sel)
        props.append((float(img_max[sl][sel].max()), rr.mean() + sl[0].start, cc.mean() + sl[1].start, k))
    props = sorted(props, key=lambda p: p[0])[::-1]      # stable sort, then reversed: ties end up in descending label order
    peaks, ys, xs = [], [], []
    yext, xext = np.mgrid[-min_ydistance : min_ydistance + 1, -min_xdistance : min_xdistance + 1]
    for _, cy, cx, _k in props:
        yi, xi = int(np.round(cy)), int(np.round(cx))
        accum = img_max[yi, xi]
        if accum > threshold:
            ynh, xnh = yi + yext, xi + xext
            inside = np.logical_and(ynh > 0, ynh < rows)
            ynh, xnh = ynh[inside], xnh[inside]
            low = xnh < 0
            ynh[low] = rows - ynh[low]
            xnh[low] += cols
            high = xnh >= cols
            ynh[high] = rows - ynh[high]
            xnh[high] -= cols
            img_max[ynh, xnh] = 0
            peaks.append(accum)
            ys.append(yi)
            xs.append(xi)
    peaks, ys, xs = np.array(peaks), np.array(ys, dtype=int), np.array(xs, dtype=int)
    if num_peaks < len(peaks):
        keep = np.argsort(peaks)[::-1][: int(num_peaks)]
        peaks, ys, xs = peaks[keep], ys[keep], xs[keep]
    return peaks, xs, ys


def hough_line_peaks(hspace, angles, dists, min_distance=9, min_angle=10, threshold=None, num_peaks=np.inf):
    min_angle = min(min_angle, hspace.shape[1])
    h, a, d = prominent_peaks(hspace, min_xdistance=min_angle, min_ydistance=min_distance, threshold=threshold, num_peaks=num_peaks)
    if a.size > 0:
        return h, angles[a], dists[d]
    return h, np.array([]), np.array([])


def stretch01(a):
    a = np.asarray(a)
    g = a - a.min()
    n = g / g.max()
    s = n * 1
    return s - s.min() + 0


def jaw_orthogonality(array):
    """contrib/orthogonality.py:29-86 -> (line_angles dict, result dict)"""
    edge = canny(stretch01(array))
    tested = np.linspace(-np.pi / 2, np.pi / 2, num=360 * 10, endpoint=False)
    h, theta, d = hough_line(edge, tested)
    _, angles, dists = hough_line_peaks(h, theta, d)
    order = np.argsort(np.abs(angles))
    sa, sd = angles[order], dists[order]
    la = {}
    if sd[0] < sd[1]:
        la["left"], la["right"] = (sa[0], sd[0]), (sa[1], sd[1])
    else:
        la["left"], la["right"] = (sa[1], sd[1]), (sa[0], sd[0])
    if sd[2] < sd[3]:
        la["bottom"], la["top"] = (sa[2], sd[2]), (sa[3], sd[3])
    else:
        la["bottom"], la["top"] = (sa[3], sd[3]), (sa[2], sd[2])
    res = {"top_left": abs(np.rad2deg(la["left"][0] - la["top"][0])), "top_right": abs(np.rad2deg(la["right"][0] - la["top"][0])),
           "bottom_left": abs(np.rad2deg(la["left"][0] - la["bottom"][0])), "bottom_right": abs(np.rad2deg(la["right"][0] - la["bottom"][0]))}
    return la, res, edge, h
