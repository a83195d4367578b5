"""TEST INFRASTRUCTURE ONLY -- never imported by the product (pylinac_b200/).

scikit-image is not installed in this container, so the reference's Winston-Lutz / disk-locator path
(metrics/utils.py:66-190, metrics/features.py:7-68) cannot run as is.  This module restates the few skimage functions that path
calls, on top of scipy, and installs them into the stub ``skimage`` package of oracle/refstub.py so that the UNMODIFIED
reference code runs end to end:

    skimage.measure.label(connectivity=1)        -> scipy.ndimage.label with the 4-neighbour structure (raster-order numbering)
    skimage.segmentation.clear_border            -> zero every label that touches the array border
    skimage.measure.regionprops(label, intensity)-> bbox, area, area_filled / filled_area, area_bbox / bbox_area, perimeter,
                                                    solidity (area / area_convex), centroid_weighted / weighted_centroid

PARITY NOTE: the definitions of ``perimeter`` (4-neighbourhood border, weights 1 / sqrt2 / (1+sqrt2)/2 from a
3x3 [[10,2,10],[2,1,2],[10,2,10]] convolution) and ``area_convex`` (pixel centres inside the convex hull of the region's pixel
corners-on-a-diamond offsets) are restated from the published skimage algorithms without the skimage source at hand: the
skimage boundary of the Winston-Lutz path is UNPINNED.  Both only enter accept / reject predicates with wide margins.
"""
from __future__ import annotations

import numpy as np
from scipy import ndimage

_STREL4 = np.array([[0, 1, 0], [1, 1, 1], [0, 1, 0]], dtype=bool)


def label(image, connectivity=None, **kwargs):
    """skimage.measure.label: connectivity None = full (8 neighbours in 2-D), 1 = 4 neighbours; raster-order numbering"""
    structure = _STREL4 if connectivity == 1 else np.ones((3, 3), bool)
    lab, _ = ndimage.label(np.asarray(image) != 0, structure=structure)
    return lab


def clear_border(labels, buffer_size=0, **kwargs):
    """skimage.segmentation.clear_border: every object with a pixel in the outer ``buffer_size + 1`` rows / columns is removed.
    A boolean input is labelled with full connectivity first (what skimage's internal re-labelling does); a label image keeps its
    labels (equal-valued pixels that touch are one object either way)."""
    labels = np.array(labels)
    border = np.zeros(labels.shape, bool)
    ext = buffer_size + 1
    border[:ext, :] = border[-ext:, :] = border[:, :ext] = border[:, -ext:] = True
    objs = label(labels, connectivity=None) if labels.dtype == bool else labels
    touching = np.unique(objs[border])
    touching = touching[touching != 0]
    labels[np.isin(objs, touching)] = 0
    return labels


def perimeter(image):
    """skimage.measure.perimeter(image, neighborhood=4)"""
    image = np.asarray(image, dtype=np.uint8)
    eroded = ndimage.binary_erosion(image, _STREL4, border_value=0)
    border = image - eroded
    weights = np.zeros(50, dtype=np.float64)
    weights[[5, 7, 15, 17, 25, 27]] = 1
    weights[[21, 33]] = np.sqrt(2)
    weights[[13, 23]] = (1 + np.sqrt(2)) / 2
    conv = ndimage.convolve(border, np.array([[10, 2, 10], [2, 1, 2], [10, 2, 10]]), mode="constant", cval=0)
    hist = np.bincount(conv.ravel(), minlength=50)
    return float(hist @ weights)


def convex_area(image):
    """np.sum(skimage.morphology.convex_hull_image(image)): pixel centres inside (or on) the hull of the pixels' diamond offsets."""
    image = np.asarray(image, dtype=bool)
    rr, cc = np.nonzero(image)
    if len(rr) < 1:
        return 0
    pts = np.stack([rr, cc], axis=1).astype(float)
    offs = np.array([[0.5, 0], [-0.5, 0], [0, 0.5], [0, -0.5]])
    pts = (pts[:, None, :] + offs[None, :, :]).reshape(-1, 2)
    hull = _monotone_chain(pts)
    gr, gc = np.mgrid[0:image.shape[0], 0:image.shape[1]]
    inside = np.ones(image.shape, bool)
    n = len(hull)
    for k in range(n):
        a, b = hull[k], hull[(k + 1) % n]
        cross = (b[0] - a[0]) * (gc - a[1]) - (b[1] - a[1]) * (gr - a[0])
        inside &= cross >= -1e-10
    return int(inside.sum())


def _monotone_chain(pts):
    """Convex hull (Andrew's monotone chain), counter-clockwise in (row, col) with the orientation test used above."""
    pts = np.unique(pts, axis=0)
    pts = pts[np.lexsort((pts[:, 1], pts[:, 0]))]

    def cross(o, a, b):
        return (a[0] - o[0]) * (b[1] - o[1]) - (a[1] - o[1]) * (b[0] - o[0])

    lower, upper = [], []
    for p in pts:
        while len(lower) >= 2 and cross(lower[-2], lower[-1], p) <= 0:
            lower.pop()
        lower.append(p)
    for p in pts[::-1]:
        while len(upper) >= 2 and cross(upper[-2], upper[-1], p) <= 0:
            upper.pop()
        upper.append(p)
    return np.array(lower[:-1] + upper[:-1])


class RegionProperties:
    def __init__(self, lab, label_img, intensity, sl=None):
        self.label = lab
        if sl is None:
            sl = ndimage.find_objects((label_img == lab).astype(np.int32))[0]
        self.slice = sl
        self.bbox = (sl[0].start, sl[1].start, sl[0].stop, sl[1].stop)
        self.image = label_img[sl] == lab
        self._intensity = None if intensity is None else intensity[sl]

    @property
    def centroid(self):
        rr, cc = np.nonzero(self.image)
        return (float(rr.mean()) + self.bbox[0], float(cc.mean()) + self.bbox[1])

    @property
    def equivalent_diameter_area(self):
        return float(np.sqrt(4 * self.area / np.pi))

    equivalent_diameter = equivalent_diameter_area

    @property
    def area(self):
        return float(self.image.sum())

    @property
    def area_filled(self):
        return float(ndimage.binary_fill_holes(self.image).sum())

    filled_area = area_filled

    @property
    def area_bbox(self):
        return float(self.image.size)

    bbox_area = area_bbox

    @property
    def perimeter(self):
        return perimeter(self.image)

    @property
    def area_convex(self):
        return float(convex_area(self.image))

    @property
    def solidity(self):
        return self.area / self.area_convex

    @property
    def centroid_weighted(self):
        w = self._intensity * self.image
        rr, cc = np.mgrid[0:self.image.shape[0], 0:self.image.shape[1]]
        tot = w.sum()
        return (float((rr * w).sum() / tot) + self.bbox[0], float((cc * w).sum() / tot) + self.bbox[1])

    weighted_centroid = centroid_weighted


def regionprops(label_image, intensity_image=None, **kwargs):
    label_image = np.asarray(label_image)
    if label_image.dtype == bool:
        label_image = label_image.astype(np.int32)
    slices = ndimage.find_objects(label_image)          # one pass for every label (None for labels that do not occur)
    return [RegionProperties(i + 1, label_image, intensity_image, sl) for i, sl in enumerate(slices) if sl is not None]


def find_boundaries(label_img, **kwargs):
    return np.zeros(np.asarray(label_img).shape, bool)


def polygon(r, c, shape=None):
    """skimage.draw.polygon(r, c, shape): integer pixel coordinates inside the polygon or on its boundary (skimage's
    point_in_polygon returns non-zero for vertex / edge hits), rows int(max(0, min r)) .. ceil(max r), clipped to ``shape``.
    Restated without the skimage source at hand (UNPINNED, like the other functions of this module); evaluated here with exact
    rational arithmetic on the float vertices (winding by the crossing rule, boundary by collinearity + range test)."""
    from fractions import Fraction

    r = [Fraction(float(v)) for v in r]
    c = [Fraction(float(v)) for v in c]
    minr, maxr = int(max(0, min(r))), int(np.ceil(float(max(r))))
    minc, maxc = int(max(0, min(c))), int(np.ceil(float(max(c))))
    if shape is not None:
        maxr, maxc = min(shape[0] - 1, maxr), min(shape[1] - 1, maxc)
    n = len(r)
    rr, cc = [], []
    for y in range(minr, maxr + 1):
        for x in range(minc, maxc + 1):
            inside, on = False, False
            for i in range(n):
                y0, x0, y1, x1 = r[i - 1], c[i - 1], r[i], c[i]
                cross = (x1 - x0) * (y - y0) - (y1 - y0) * (x - x0)
                if cross == 0 and min(x0, x1) <= x <= max(x0, x1) and min(y0, y1) <= y <= max(y0, y1):
                    on = True
                    break
                if (y0 > y) != (y1 > y):
                    xi = x0 + (y - y0) * (x1 - x0) / (y1 - y0)
                    if x < xi:
                        inside = not inside
            if on or inside:
                rr.append(y)
                cc.append(x)
    return np.array(rr, dtype=np.intp), np.array(cc, dtype=np.intp)


class EuclideanTransform:
    """skimage.transform.EuclideanTransform(rotation=, translation=): params = [[cos, -sin, tx], [sin, cos, ty], [0, 0, 1]]"""

    def __init__(self, matrix=None, rotation=None, translation=None, **kwargs):
        if matrix is not None:
            self.params = np.array(matrix, dtype=float)
            return
        rotation = 0.0 if rotation is None else rotation
        translation = (0, 0) if translation is None else translation
        cs, sn = np.cos(rotation), np.sin(rotation)
        self.params = np.array([[cs, -sn, translation[0]], [sn, cs, translation[1]], [0, 0, 1]], dtype=float)

    # skimage's ProjectiveTransform.__add__: "self, then other" = other.params @ self.params (used by vmat.py:1071-1072)
    def __add__(self, other):
        return EuclideanTransform(matrix=other.params @ self.params)

    @property
    def translation(self):
        return self.params[0:2, 2]

    @property
    def rotation(self):
        import math

        return math.atan2(self.params[1, 0], self.params[1, 1])


def matrix_transform(coords, matrix):
    """skimage.transform.matrix_transform: homogeneous coordinates times the transposed matrix, de-homogenised"""
    m = matrix.params if hasattr(matrix, "params") else np.asarray(matrix)
    coords = np.atleast_2d(np.asarray(coords, dtype=float))
    src = np.hstack([coords, np.ones((coords.shape[0], 1))])
    dst = src @ m.T
    dst[dst[:, 2] == 0, 2] = np.finfo(float).eps
    return dst[:, :2] / dst[:, 2:3]


def install():
    """Bind the restated functions where the reference looks them up (after oracle.refstub.import_reference()): the modules that
    did ``from skimage import measure, segmentation`` hold stub objects, so the names are replaced in those modules."""
    import types

    from oracle.refstub import import_reference

    import_reference()
    import pylinac.metrics.features as rfeatures
    import pylinac.metrics.utils as rutils

    rutils.measure = types.SimpleNamespace(label=label, regionprops=regionprops)
    rutils.segmentation = types.SimpleNamespace(clear_border=clear_border, find_boundaries=find_boundaries)
    rutils.RegionProperties = RegionProperties
    rutils.find_boundaries = find_boundaries       # plotting only: an empty outline of the right shape
    rfeatures.RegionProperties = RegionProperties
    import pylinac.metrics.image as rmimage

    rmimage.measure = types.SimpleNamespace(label=label, regionprops=regionprops)
    rmimage.segmentation = types.SimpleNamespace(clear_border=clear_border, find_boundaries=find_boundaries)
    import pylinac.core.roi as rroi

    rroi.polygon = polygon
    import pylinac.core.geometry as rgeo

    rgeo.transform = types.SimpleNamespace(EuclideanTransform=EuclideanTransform, matrix_transform=matrix_transform)
    import sys

    if "pylinac.vmat" in sys.modules or True:
        import pylinac.vmat as rvmat

        rvmat.EuclideanTransform = EuclideanTransform
