"""oracle/edges_oracle.py (restated skimage canny / hough_line / hough_line_peaks; PARITY UNPINNED, see its header) against what can be
checked without scikit-image: geometric ground truth of rotated synthetic fields and OpenCV's Hough transform on the same edge map."""
import numpy as np
import pytest

from oracle import edges_oracle as eo
from oracle import synth


def _field(rotation, seed):
    fr = synth.Frame((512, 640), 0.5, sid=1000.0)
    fr.add_perfect_field((120, 150), cax_offset_mm=(3, -4), alpha=0.6, rotation=rotation)
    fr.gaussian(1.5)
    fr.noise(0.002, seed=seed)
    return fr.image


@pytest.mark.parametrize("rotation", [0.0, 0.7, -1.3])
def test_jaw_orthogonality_of_a_rotated_rectangle(rotation):
    la, res, edge, h = eo.jaw_orthogonality(_field(rotation, 5))
    assert 600 < edge.sum() < 3000                       # thin outline of a 240 x 300 px rectangle
    for v in res.values():
        assert abs(v - 90.0) < 0.8
    assert abs(np.rad2deg(la["left"][0]) + rotation) < 0.6
    assert la["left"][1] < la["right"][1] and la["bottom"][1] < la["top"][1]


def test_hough_accumulator_agrees_with_opencv():
    import cv2

    img = _field(0.7, 9)
    edge = eo.canny(eo.stretch01(img))
    theta = np.linspace(-np.pi / 2, np.pi / 2, 3600, endpoint=False)
    h, _, d = eo.hough_line(edge, theta)
    r, t = np.unravel_index(np.argmax(h), h.shape)
    lines = cv2.HoughLines(edge.astype(np.uint8) * 255, 1, np.pi / 3600, int(h.max() * 0.9))
    rho, th = lines[0][0]
    mt, mr = theta[t], d[r]
    if mt < 0:
        mt, mr = mt + np.pi, -mr
    assert abs(mt - th) < np.deg2rad(0.3) and abs(mr - rho) < 2.5
