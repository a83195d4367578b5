"""Whole-frame locators (GlobalSizedDiskLocator / GlobalSizedFieldLocator / GlobalFieldLocator, csrc/locate.cu) against goldens of
the UNMODIFIED reference (tests/golden/make_locator_golden.py; skimage restated by oracle/skimage_shim.py, unpinned there).  Point
lists must agree in number, ORDER and position (weighted / unweighted centroids: fp64 sums in another order -> 1e-9 px)."""
import numpy as np
import pytest

from tests.golden.locator_cases import CASES, case

pytestmark = pytest.mark.gpu
GOLD = np.load("tests/golden/locator_golden.npz", allow_pickle=False)


def _metric(spec):
    from pylinac_b200.metrics import image as mi

    cls = getattr(mi, spec["cls"])
    return cls.from_physical(**spec["kw"]) if spec.get("physical") else cls(**spec["kw"])


@pytest.mark.parametrize("name", CASES)
def test_global_locator_matches_reference_golden(name):
    from pylinac_b200.core import image

    a, ps, sid, spec = case(name)
    img = image.ArrayImage(a, dpi=25.4 / ps, sid=sid)
    if int(GOLD[f"{name}/raised"]):
        with pytest.raises(ValueError, match="Couldn't find the minimum number"):
            img.compute(metrics=_metric(spec))
        return
    pts = img.compute(metrics=_metric(spec))
    got = np.array([[p.x, p.y] for p in pts]).reshape(-1, 2)
    want = GOLD[f"{name}/points"]
    assert got.shape == want.shape, (got, want)
    np.testing.assert_allclose(got, want, rtol=0, atol=1e-9)


def test_disk_locator_batch_equals_single_frames():
    from pylinac_b200.metrics import image as mi

    a, ps, sid, spec = case("disks4")
    b = a[::-1, ::-1].copy()
    dpmm = (1 / ps) * sid / 1000.0
    out = mi.locate_disks_batch(np.stack([a, b, a]), dpmm, **spec["kw"])
    want = GOLD["disks4/points"]
    np.testing.assert_allclose(np.array([[p.x, p.y] for p in out[0]]), want, atol=1e-9)
    np.testing.assert_allclose(np.array([[p.x, p.y] for p in out[2]]), want, atol=1e-9)
    assert len(out[1]) == len(want)
    # the flipped frame finds the mirrored disks (as a set: the visiting order changes with the flip)
    mirrored = np.array([[a.shape[1] - 1 - x, a.shape[0] - 1 - y] for x, y in want])
    got = np.array([[p.x, p.y] for p in out[1]])
    for m in mirrored:
        assert np.min(np.hypot(*(got - m).T)) < 0.05


def test_large_field_goes_through_the_big_tile_path():
    """A 150 mm field on a 1280 x 1280 panel (tile 448 x 448 > the 200 KB shared tile is not needed; 300 mm is): both sizes found."""
    from oracle import synth
    from pylinac_b200.core import image
    from pylinac_b200.metrics import image as mi

    for size, tol in ((150, 5), (300, 8)):
        fr = synth.as1200(1000.0)
        fr.add_perfect_field((size, size), cax_offset_mm=(3.0, -2.0), alpha=0.6)
        fr.gaussian(1.0)
        fr.noise(0.001, seed=size)
        img = image.ArrayImage(fr.image, dpi=25.4 / fr.pixel_size, sid=1000.0)
        pts = img.compute(metrics=mi.GlobalSizedFieldLocator.from_physical(size, size, tol, max_number=1))
        assert len(pts) == 1
        # field centre: image centre + offset (x = second offset component)
        assert abs(pts[0].x - (640 - 0.5 - 2.0 * fr.dpmm)) < 1.0 and abs(pts[0].y - (640 - 0.5 + 3.0 * fr.dpmm)) < 1.0


def test_disk_locator_chunks_large_batches():
    """More frames than one labelling chunk (32): every frame still gets its own, identical answer."""
    from pylinac_b200.metrics import image as mi

    a, ps, sid, spec = case("disks4")
    dpmm = (1 / ps) * sid / 1000.0
    out = mi.locate_disks_batch(np.stack([a] * 35), dpmm, **spec["kw"])
    want = GOLD["disks4/points"]
    for k in (0, 31, 32, 34):
        np.testing.assert_allclose(np.array([[p.x, p.y] for p in out[k]]), want, atol=1e-9)
