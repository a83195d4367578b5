"""Image combination and resampling against the UNMODIFIED reference / scipy (goldens: tests/golden/make_combine_golden.py):
``load_multiples`` (core/image.py:306-360), ``convert_to_dtype`` (core/array_utils.py:172-198), ``scipy.ndimage.zoom`` semantics of
the device zoom kernel, ``equate_images`` (core/image.py:169-220) and the two ``from_multiple_images`` constructors end to end
(picketfence.py:357-400, starshot.py:148-174) including the reference's in-memory DICOM write / read of the composite."""
import hashlib

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from tests.golden import combine_cases as cc

G = np.load("tests/golden/combine_golden.npz")


def _sha(a):
    return np.frombuffer(hashlib.sha1(np.ascontiguousarray(a).tobytes()).digest(), dtype=np.uint8)


@pytest.mark.parametrize("method", ["mean", "max", "sum"])
@pytest.mark.parametrize("stretch", [True, False])
def test_load_multiples_arrays(method, stretch):
    from pylinac_b200.core import image

    img = image.load_multiples([a.copy() for a in cc.small_stack()], method=method, stretch_each=stretch, dpi=100, sid=1000)
    want = G[f"lm/{method}/{int(stretch)}"]
    assert img.array.dtype == want.dtype
    np.testing.assert_allclose(img.array, want, rtol=0, atol=1e-15 * max(1.0, float(np.abs(want).max())))
    assert img._raw_pixels is True


def test_load_multiples_fill_dtype_and_errors():
    from pylinac_b200.core import image

    st = cc.small_stack()
    img = image.load_multiples([a.copy() for a in st[:2]], method="mean", stretch_each=True, dtype=np.uint16, dpi=100, sid=1000)
    np.testing.assert_array_equal(img.array, G["lm/mean/u16"])
    with pytest.raises(ValueError):
        image.load_multiples([st[0], st[0][:-1]], dpi=100, sid=1000)


def test_convert_to_dtype_matches_reference():
    from pylinac_b200.core import array_utils as au

    st = cc.small_stack()
    with np.errstate(invalid="ignore"):
        np.testing.assert_array_equal(au.convert_to_dtype(st[2], np.uint16), G["ctd/f_u16"])
        np.testing.assert_array_equal(au.convert_to_dtype((st[0] >> 4).astype(np.uint8), np.uint16), G["ctd/u8_u16"])
        np.testing.assert_array_equal(au.convert_to_dtype(st[1], np.uint8), G["ctd/u16_u8"])


@pytest.mark.parametrize("name", list(cc.ZOOM_CASES))
def test_zoom_matches_scipy(name):
    from pylinac_b200.core import array_utils as au

    shape, z, order, mode = cc.ZOOM_CASES[name]
    got = au.zoom(cc.zoom_input(name), z, order=order, mode=mode)
    assert tuple(got.shape) == tuple(G[f"zoom/{name}/shape"])
    if name == "z2d_big":
        got = got[::3, ::3]
    want = G[f"zoom/{name}"]
    # scipy evaluates the same separable B-spline sums in another order (recursive prefilter gain vs one tridiagonal solve):
    # agreement is rounding-level relative to the data range (~1500)
    np.testing.assert_allclose(got, want, rtol=0, atol=2e-9)


def test_zoom_rejects_bad_input():
    from pylinac_b200.core import array_utils as au

    with pytest.raises(ValueError):
        au.zoom(np.zeros((4, 4)), 2.0, mode="reflect")


@pytest.mark.parametrize("tag,swap", [("eq", False), ("eq2", True)])
def test_equate_images(tag, swap):
    from pylinac_b200.core import image

    (a, adpi), (b, bdpi) = cc.equate_inputs()
    one, two = image.load(a, dpi=adpi, sid=1000), image.load(b, dpi=bdpi, sid=1000)
    if swap:
        one, two = two, one
    i1, i2 = image.equate_images(one, two)
    shapes = G[f"{tag}/shapes"]
    assert tuple(i1.shape) == tuple(shapes[0]) and tuple(i2.shape) == tuple(shapes[1])
    np.testing.assert_array_equal(_sha(np.asarray(i1.array)), G[f"{tag}/a_sha1"])        # the crop is exact
    np.testing.assert_allclose(np.asarray(i2.array)[::2, ::2], G[f"{tag}/b"], rtol=0, atol=2e-9)
    np.testing.assert_allclose([i1.dpi, i2.dpi], G[f"{tag}/dpi"], rtol=1e-14)
    # the inputs are untouched (deep copies, core/image.py:186-187)
    assert one.shape in ((480, 640), (360, 300)) and two.shape in ((480, 640), (360, 300))


def _write_parts(tmp_path, parts):
    from tests.dicom_writer import write_dicom

    return [write_dicom(tmp_path / f"part{k}.dcm", p, pixel_spacing_mm=cc.PS, sid=cc.SID) for k, p in enumerate(parts)]


@pytest.mark.parametrize("tag,kw", [("mean", {}), ("sum_nostretch", {"method": "sum", "stretch_each": False})])
def test_pf_from_multiple_images(tmp_path, tag, kw):
    from pylinac_b200 import picketfence as pfm

    paths = _write_parts(tmp_path, cc.pf_parts())
    with np.errstate(invalid="ignore"):
        pf = pfm.PicketFence.from_multiple_images(paths, **kw)
    # the composite the reference writes to its in-memory DICOM, bit for bit
    stored = pf._raw._stored
    assert stored.dtype == np.uint16
    np.testing.assert_array_equal(stored[::8, ::8], G[f"pf/{tag}/stored_sub"])
    np.testing.assert_array_equal(_sha(stored), G[f"pf/{tag}/stored_sha1"])
    pf.analyze()
    rd = pf.results_data()
    assert rd.number_of_pickets == int(G[f"pf/{tag}/number_of_pickets"])
    pos = np.array([list(m.position) for m in pf.mlc_meas])
    np.testing.assert_array_equal([m.leaf_num for m in pf.mlc_meas], G[f"pf/{tag}/leaf"])
    np.testing.assert_array_equal([m.picket_num for m in pf.mlc_meas], G[f"pf/{tag}/picket"])
    np.testing.assert_allclose(pos, G[f"pf/{tag}/position"], rtol=0, atol=1e-6)
    np.testing.assert_allclose(np.array([list(m.error) for m in pf.mlc_meas]), G[f"pf/{tag}/error"], rtol=0, atol=1e-6)
    assert abs(rd.max_error_mm - float(G[f"pf/{tag}/max_error"])) < 1e-6
    np.testing.assert_allclose(rd.offsets_from_cax_mm, G[f"pf/{tag}/offsets"], rtol=0, atol=1e-6)


def test_starshot_from_multiple_images(tmp_path):
    from pylinac_b200 import starshot as sm

    paths = _write_parts(tmp_path, cc.star_parts())
    with np.errstate(invalid="ignore"):
        st = sm.Starshot.from_multiple_images(paths)
    stored = st.image._stored
    np.testing.assert_array_equal(stored[::8, ::8], G["star/stored_sub"])
    np.testing.assert_array_equal(_sha(stored), G["star/stored_sha1"])
    st.analyze()
    w = G["star/wobble"]
    assert abs(st.wobble.center.x - w[0]) < 0.01 and abs(st.wobble.center.y - w[1]) < 0.01
    assert abs(st.wobble.radius - w[2]) < 0.01 and abs(st.wobble.radius_mm - w[3]) < 0.01
    assert len(st.lines) == int(G["star/npeaks"])
