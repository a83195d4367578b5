"""Canny / Hough / JawOrthogonality on the GPU (csrc/edges.cu) against the restated skimage operators of oracle/edges_oracle.py
(PARITY UNPINNED: scikit-image is absent and the reference holds no vectors for contrib/orthogonality.py) and against geometric
ground truth: edge maps and accumulators bit for bit, the selected lines identical, the corner angles of a rotated synthetic field
within the 0.05 degree angular resolution of the transform plus the pixelation of the rotated edge."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _field(rotation, seed, size=(100, 120), offset=(3, -4), shape=None):
    from oracle import synth

    fr = synth.as1200(1000.0) if shape is None else synth.Frame(shape, 0.5, sid=1000.0)
    fr.add_perfect_field(size, cax_offset_mm=offset, alpha=0.6, rotation=rotation)
    fr.gaussian(1.5)
    fr.noise(0.002, seed=seed)
    return fr.image


@pytest.mark.parametrize("rotation,seed", [(0.0, 1), (0.7, 2), (-1.3, 3)])
def test_canny_and_hough_match_the_restated_operators(rotation, seed):
    from oracle import edges_oracle as eo
    from pylinac_b200 import _native as nat

    img = _field(rotation, seed, shape=(512, 640), size=(120, 150))
    st = eo.stretch01(img)
    ctx = nat.Context.default()
    edges = nat.canny(ctx, st)
    want = eo.canny(st)
    assert edges.shape == want.shape and edges.dtype == bool
    assert np.array_equal(edges, want), int((edges != want).sum())
    theta = np.linspace(-np.pi / 2, np.pi / 2, 720, endpoint=False)
    acc, offset = nat.hough_line(ctx, edges, theta)
    try:
        h, _, d = eo.hough_line(want, theta)
        got = acc.download()[0]
        assert offset == int(np.ceil(np.hypot(*img.shape))) and got.shape == h.shape
        assert np.array_equal(got.astype(np.uint64), h)
        from pylinac_b200.contrib.orthogonality import hough_line_peaks

        hv, ang, dist = hough_line_peaks(ctx, acc, theta, d)
        hv2, ang2, dist2 = eo.hough_line_peaks(h, theta, d)
        assert np.array_equal(hv, hv2) and np.array_equal(ang, ang2) and np.array_equal(dist, dist2)
    finally:
        acc.free()


@pytest.mark.parametrize("rotation", [0.0, 0.7, -1.3])
def test_jaw_orthogonality_matches_oracle_and_geometry(rotation):
    from oracle import edges_oracle as eo
    from pylinac_b200.contrib.orthogonality import JawOrthogonality

    img = _field(rotation, 7)
    j = JawOrthogonality(img)
    j.analyze()
    la, res, edge, _ = eo.jaw_orthogonality(img)
    assert np.array_equal(j.edge_image, edge)
    for k in ("left", "right", "top", "bottom"):
        assert j.line_angles[k]["angle"] == la[k][0] and j.line_angles[k]["dist"] == la[k][1]
    assert j.results() == pytest.approx(res, abs=1e-12)
    # geometry: a rectangle has right angles; the vertical jaws sit at -rotation, the horizontal ones 90 degrees from them
    for v in j.results().values():
        assert abs(v - 90.0) < 0.8
    assert abs(np.rad2deg(j.line_angles["left"]["angle"]) + rotation) < 0.6
    assert abs(abs(np.rad2deg(j.line_angles["top"]["angle"])) - (90 - abs(rotation))) < 0.6 or abs(rotation) < 1e-9


def test_hough_agrees_with_opencv_on_the_strongest_line():
    import cv2

    from pylinac_b200 import _native as nat
    from oracle import edges_oracle as eo

    img = _field(0.7, 11, shape=(512, 640), size=(120, 150))
    ctx = nat.Context.default()
    edges = nat.canny(ctx, eo.stretch01(img))
    theta = np.linspace(-np.pi / 2, np.pi / 2, 3600, endpoint=False)
    acc, offset = nat.hough_line(ctx, edges, theta)
    try:
        a = acc.download()[0]
    finally:
        acc.free()
    r, t = np.unravel_index(np.argmax(a), a.shape)
    lines = cv2.HoughLines(edges.astype(np.uint8) * 255, 1, np.pi / 3600, int(a.max() * 0.9))
    assert lines is not None
    rho, th = lines[0][0]
    # OpenCV: theta in [0, pi), rho signed; skimage: theta in [-pi/2, pi/2)
    mine_theta, mine_rho = theta[t], r - offset
    if mine_theta < 0:
        mine_theta, mine_rho = mine_theta + np.pi, -mine_rho
    assert abs(mine_theta - th) < np.deg2rad(0.3) and abs(mine_rho - rho) < 2.5
