"""GPU parity of the operator layer (SURVEY.md section 8a rows a1-a10, a15-a19, config 1 "plumbing"): every primitive of
core.array_utils / core.image / core.profile through the C-ABI vs the numpy / scipy expression the reference evaluates
(core/array_utils.py:38-212, core/image.py:695-926, core/profile.py:2021-2176, 2545-2649).  Integer results bit-exact, fp64
results exact or to 1 ulp where the operation order is numpy's."""
import numpy as np
import pytest
from scipy import ndimage

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def frame():
    from oracle import synth

    return synth.bench_pf_frame(7)


def test_normalize_invert_ground_stretch(frame):
    from pylinac_b200.core import array_utils as au

    a = frame
    np.testing.assert_array_equal(au.normalize(a), a / a.max())
    np.testing.assert_array_equal(au.normalize(a, value=1234.5), a / 1234.5)
    np.testing.assert_array_equal(au.invert(a), -a + a.max() + a.min())
    b = a[100:200, 300:500].copy()
    np.testing.assert_array_equal(au.invert(b), -b + b.max() + b.min())
    np.testing.assert_array_equal(au.bit_invert(a), np.invert(a))
    with pytest.raises(ValueError):
        au.bit_invert(a.astype(np.float64))
    g = a.astype(np.int32) + 77
    np.testing.assert_array_equal(au.ground(g), g - g.min())
    np.testing.assert_array_equal(au.ground(g, value=5), g - g.min() + 5)
    f64 = a[:64, :64].astype(np.float64) * 0.37 + 3
    np.testing.assert_array_equal(au.invert(f64), -f64 + f64.max() + f64.min())
    np.testing.assert_array_equal(au.ground(f64), f64 - f64.min())
    # stretch: ground -> normalize -> scale -> offset (core/array_utils.py:142-168)
    s = au.stretch(f64, min=0, max=1)
    ref = f64 - f64.min()
    ref = ref / ref.max()
    np.testing.assert_allclose(s, ref, rtol=0, atol=1e-15)


@pytest.mark.parametrize("size", [3, 5])
def test_median_filter(frame, size):
    from pylinac_b200.core import array_utils as au

    a = frame[200:520, 100:612]
    np.testing.assert_array_equal(au.filter(a, size=size, kind="median"), ndimage.median_filter(a, size=size))


@pytest.mark.parametrize("sigma", [1, 2, 4])
def test_gaussian_filter_integer_and_float(frame, sigma):
    from pylinac_b200.core import array_utils as au

    a = frame[200:520, 100:612]
    np.testing.assert_array_equal(au.filter(a, size=sigma, kind="gaussian"), ndimage.gaussian_filter(a, sigma=sigma))
    f = a.astype(np.float64) / 7.0
    np.testing.assert_allclose(au.filter(f, size=sigma, kind="gaussian"), ndimage.gaussian_filter(f, sigma=sigma), rtol=1e-15, atol=1e-12)
    p = f[17]                                       # 1-D profile
    np.testing.assert_allclose(au.filter(p, size=sigma, kind="gaussian"), ndimage.gaussian_filter(p, sigma=sigma), rtol=1e-15, atol=1e-12)
    with pytest.raises(ValueError):
        au.filter(a, size=3, kind="bogus")
    with pytest.raises(ValueError):
        au.filter(a, size=1.5, kind="median")        # floats outside (0, 1) are rejected (core/array_utils.py:121-128)
    with pytest.raises(ValueError):
        au.filter(a, size=3.5, kind="gaussian")


def test_config1_plumbing_gaussian_then_threshold(frame):
    """BASELINE.json configs[0]: one synthetic EPID frame through Image.filter(gaussian) + threshold."""
    from pylinac_b200.core import image

    img = image.load(frame.copy(), dpi=25.4 / 0.390625, sid=1000)
    img.filter(size=2, kind="gaussian")
    ref = ndimage.gaussian_filter(frame, sigma=2)
    np.testing.assert_array_equal(img.array, ref)
    t = 0.5 * float(ref.max())
    img.threshold(t)
    np.testing.assert_array_equal(img.array, np.where(ref >= t, ref, 0))
    img2 = image.load(frame.copy(), dpi=25.4 / 0.390625, sid=1000)
    bi = img2.as_binary(t)
    np.testing.assert_array_equal(bi.array, np.where(frame >= t, 1, 0))
    img2.threshold(t, kind="low")
    np.testing.assert_array_equal(img2.array, np.where(frame <= t, frame, 0))


def test_image_operators(frame):
    from pylinac_b200.core import image

    img = image.load(frame.copy(), dpi=25.4 / 0.390625, sid=1000)
    assert abs(img.dpmm - 2.56) < 1e-12
    assert (img.center.x, img.center.y) == (frame.shape[1] / 2 - 0.5, frame.shape[0] / 2 - 0.5)
    mn = img.ground()
    assert mn == frame.min()
    np.testing.assert_array_equal(img.array, frame - frame.min())
    img.normalize()
    np.testing.assert_array_equal(img.array, (frame - frame.min()) / (frame - frame.min()).max())
    img = image.load(frame.copy(), dpi=65, sid=1000)
    img.crop(pixels=8)
    np.testing.assert_array_equal(img.array, frame[8:-8, 8:-8])
    with pytest.raises(ValueError):
        img.crop(pixels=-1)
    with pytest.raises(ValueError):
        img.crop(pixels=4000)
    img = image.load(frame.copy(), dpi=65, sid=1000)
    img.roll(direction="y", amount=3)
    np.testing.assert_array_equal(img.array, np.roll(frame, 3, axis=0))
    img.rot90(1)
    np.testing.assert_array_equal(img.array, np.rot90(np.roll(frame, 3, axis=0), 1))
    assert img.dist2edge_min((10, 20)) == 10


def test_check_inversion_variants(frame):
    from pylinac_b200.core import image

    inv = (-frame + frame.max() + frame.min()).astype(np.uint16)
    for src in (frame, inv):
        img = image.load(src.copy(), dpi=65, sid=1000)
        was = img.check_inversion_by_histogram(percentiles=(5, 50, 95))
        p = [np.percentile(src, q) for q in (5, 50, 95)]
        expect = abs(p[1] - p[0]) > abs(p[1] - p[2])
        assert was == expect
        np.testing.assert_array_equal(img.array, (-src + src.max() + src.min()) if expect else src)
        img = image.load(src.copy(), dpi=65, sid=1000)
        img.check_inversion(box_size=10, position=(0.01, 0.01))
        rp, cp, b = max(int(0.01 * src.shape[0]), 1), max(int(0.01 * src.shape[1]), 1), 10
        boxes = (src[rp:rp + b, cp:cp + b], src[-rp - b:-rp, cp:cp + b], src[rp:rp + b, -cp - b:-cp], src[-rp - b:-rp, -cp - b:-cp])
        expect = np.mean(boxes) > np.mean(src.flatten())
        np.testing.assert_array_equal(img.array, (-src + src.max() + src.min()) if expect else src)


def test_percentiles_and_stats_are_exact(frame):
    from pylinac_b200 import _native as nat

    ctx = nat.Context.default()
    b = nat.Batch.upload(ctx, np.stack([frame, frame[::-1, ::-1] // 3]))
    qs = [0.01, 0.5, 4, 50, 90, 96, 99.5, 99.99]          # at most 8 per call
    st = nat.frame_stats(ctx, b, percentiles=qs)
    for i, src in enumerate((frame, frame[::-1, ::-1] // 3)):
        assert st["min"][i] == src.min() and st["max"][i] == src.max() and st["sum"][i] == src.sum(dtype=np.int64)
        np.testing.assert_array_equal(st["rowsum"][i], src.sum(axis=1, dtype=np.int64))
        np.testing.assert_array_equal(st["colsum"][i], src.sum(axis=0, dtype=np.int64))
        np.testing.assert_array_equal(st["percentiles"][i], np.percentile(src, qs))
    view = (100, 200, 300, 400)
    sv = nat.frame_stats(ctx, b, view=view, percentiles=[90])
    sub = frame[100:400, 200:600]
    assert sv["percentiles"][0, 0] == np.percentile(sub, 90) and sv["sum"][0] == sub.sum(dtype=np.int64)
    b.free()


def test_find_peaks_and_profile_classes_match_the_reference_semantics():
    from oracle.pf_oracle import ref_find_peaks
    from pylinac_b200.core.profile import FWXMProfile, MultiProfile, find_peaks

    rng = np.random.default_rng(4)
    x = np.linspace(0, 1, 1500)
    prof = sum(np.exp(-0.5 * ((x - c) / 0.012) ** 2) * h for c, h in ((0.12, 0.7), (0.33, 1.0), (0.5, 0.4), (0.71, 0.9), (0.9, 0.55)))
    prof = prof + rng.normal(0, 0.01, x.size)
    for kw in ({}, {"threshold": 0.3, "peak_separation": 0.05}, {"threshold": 0.5, "max_number": 2, "peak_sort": "peak_heights"},
               {"fwxm_height": 0.8, "max_number": 1}, {"search_region": (0.2, 0.8), "threshold": 0.2, "peak_separation": 30},
               {"required_prominence": 0.3, "min_width": 3}):
        idx, props = find_peaks(prof, **kw)
        ridx, rprops = ref_find_peaks(prof, **kw)
        np.testing.assert_array_equal(idx, ridx)
        for k in ("peak_heights", "prominences", "left_bases", "right_bases"):
            np.testing.assert_array_equal(props[k], rprops[k], err_msg=k)
        for k in ("widths", "width_heights", "left_ips", "right_ips"):
            np.testing.assert_allclose(props[k], rprops[k], rtol=0, atol=1e-12, err_msg=k)
    mp = MultiProfile(prof)
    pi, pv = mp.find_peaks(threshold=0.3, min_distance=0.05)
    ri, rp = ref_find_peaks(prof, threshold=0.3, peak_separation=0.05)
    np.testing.assert_array_equal(pi, ri)
    vi, vv = mp.find_valleys(threshold=0.5, min_distance=0.05)
    rvi, _ = ref_find_peaks(-prof, threshold=0.5, peak_separation=0.05)
    np.testing.assert_array_equal(vi, rvi)
    np.testing.assert_array_equal(vv, prof[rvi])
    fi, fv = mp.find_fwxm_peaks(threshold=0.3, min_distance=0.05)
    _, rp2 = ref_find_peaks(prof, threshold=0.3, peak_separation=0.05)
    np.testing.assert_array_equal(fi, [int(round(l + (r - l) / 2)) for l, r in zip(rp2["left_ips"], rp2["right_ips"])])
    single = np.exp(-0.5 * ((x - 0.47) / 0.1) ** 2)
    fw = FWXMProfile(single, fwxm_height=80)
    _, rp3 = ref_find_peaks(single, fwxm_height=0.8, max_number=1)
    assert abs(fw.center_idx - (abs(rp3["right_ips"][0] - rp3["left_ips"][0]) / 2 + rp3["left_ips"][0])) < 1e-12
    assert abs(fw.field_width_px - (rp3["right_ips"][0] - rp3["left_ips"][0])) < 1e-12


def test_circle_profiles_match_map_coordinates(frame):
    """CircleProfile / CollapsedCircleProfile (core/profile.py:2179-2283, 2405-2483) vs the reference expression."""
    from pylinac_b200.core.profile import CircleProfile, CollapsedCircleProfile

    img = frame
    for cls, kw in ((CircleProfile, dict(sampling_ratio=1.0)), (CircleProfile, dict(sampling_ratio=2.0, start_angle=0.3, ccw=False)),
                    (CollapsedCircleProfile, dict(sampling_ratio=3, width_ratio=0.1, num_profiles=20)),
                    (CollapsedCircleProfile, dict(sampling_ratio=1, width_ratio=0.25, num_profiles=7, start_angle=1.0))):
        cx, cy, r = 500.3, 520.7, 371.5
        p = cls((cx, cy), r, img, **kw)
        collapsed = cls is CollapsedCircleProfile
        wr, nprof = kw.get("width_ratio", 0.0), kw.get("num_profiles", 1)
        radii = np.linspace(r * (1 - wr), r * (1 + wr), nprof) if collapsed else np.array([r])
        size = np.pi * max(radii) * 2 * kw.get("sampling_ratio", 1.0)
        interval = (2 * np.pi) / size
        sa = kw.get("start_angle", 0)
        rads = np.arange(0 + sa, (2 * np.pi) + sa - interval, interval)
        if kw.get("ccw", True):
            rads = rads[::-1]
        ref = np.zeros(len(rads))
        for rr in radii:
            ref += ndimage.map_coordinates(img, [np.sin(rads) * rr + cy, np.cos(rads) * rr + cx], order=0)
        if collapsed:
            ref /= nprof
        assert len(p.values) == len(rads)
        # cos / sin of the device agree with numpy to ~1 ulp: a sample can flip only when it sits on a pixel boundary
        mism = np.flatnonzero(p.values != ref)
        assert len(mism) <= 2, len(mism)
        np.testing.assert_allclose(p.x_locations, np.cos(rads) * r + cx, rtol=0, atol=1e-9)
        np.testing.assert_allclose(p.y_locations, np.sin(rads) * r + cy, rtol=0, atol=1e-9)
    with pytest.raises(ValueError):
        CircleProfile((900, 900), 400, img)


@pytest.mark.parametrize("name", ["gauss_shifted", "defaults", "no_ground", "inverted_pair"])
def test_gamma_matches_the_reference(name):
    """BaseImage.gamma (core/image.py:928-1017) against maps produced by the unmodified reference (tests/golden/make_gamma_golden.py):
    same nan pattern; values to 1e-6 relative (the float32 hypot / sqrt of the denominator may differ from glibc's in the last ulp)."""
    from pylinac_b200.core import image
    from tests.golden.gamma_cases import case_images

    a, b, dpi, kw = case_images(name)
    g = image.ArrayImage(a, dpi=dpi).gamma(image.ArrayImage(b, dpi=dpi), **kw)
    ref = np.load("tests/golden/gamma_golden.npz")[name]
    assert g.dtype == np.float64 and g.shape == ref.shape
    assert np.array_equal(np.isnan(g), np.isnan(ref))
    m = ~np.isnan(ref)
    np.testing.assert_allclose(g[m], ref[m], rtol=1e-6, atol=1e-12)
    with pytest.raises(AttributeError):
        image.ArrayImage(a, dpi=dpi).gamma(image.ArrayImage(b, dpi=dpi + 5))
    with pytest.raises(AttributeError):
        image.ArrayImage(a, dpi=dpi).gamma(image.ArrayImage(b[:-4], dpi=dpi))
    with pytest.raises(ValueError):      # integer images cannot take the nan threshold mask: the reference raises the same way
        image.ArrayImage(a, dpi=dpi).gamma(image.ArrayImage(b, dpi=dpi), normalize=False)


@pytest.mark.parametrize("dtype", [np.uint16, np.uint8, np.float64])
@pytest.mark.parametrize("angle,mode", [(17.3, "edge"), (-90.0, "edge"), (45.0, "constant"), (180.0, "edge")])
def test_rotate_matches_bilinear_warp(dtype, angle, mode):
    """BaseImage.rotate (core/image.py:780-783 -> skimage.transform.rotate defaults; scikit-image is absent, so the check is an
    independent restatement with scipy.ndimage.affine_transform(order=1) on the same inverse map and img_as_float scaling)."""
    from scipy import ndimage

    from pylinac_b200.core import image

    rng = np.random.default_rng(3)
    a = (rng.random((61, 83)) * (255 if dtype == np.uint8 else 60000)).astype(dtype)
    img = image.ArrayImage(a.copy(), dpi=100, sid=1000)
    img.rotate(angle, mode=mode)
    assert img.array.dtype == np.float64 and img.shape == a.shape
    scale = {np.uint8: 1 / 255.0, np.uint16: 1 / 65535.0, np.float64: 1.0}[dtype]
    t = np.deg2rad(angle)
    ca, sa = np.cos(t), np.sin(t)
    cy, cx = a.shape[0] / 2 - 0.5, a.shape[1] / 2 - 0.5
    m = np.array([[ca, sa], [-sa, ca]])                       # (row, col) of the input as a function of (row, col) of the output
    off = np.array([cy, cx]) - m @ np.array([cy, cx])
    want = ndimage.affine_transform(a.astype(np.float64) * scale, m, offset=off, order=1,
                                    mode="nearest" if mode == "edge" else "grid-constant", cval=0.0)
    np.testing.assert_allclose(img.array, want, rtol=0, atol=1e-9 * max(1.0, float(want.max())))
    if angle == 180.0:                                         # exact half turn: the flipped image, scaled
        # sin(pi) = 1.2e-16 in fp64: sample positions are off by ~1e-14 px, times the pixel-to-pixel gradient
        np.testing.assert_allclose(img.array, a[::-1, ::-1].astype(np.float64) * scale, rtol=0,
                                   atol=1e-13 * max(a.shape) * max(1.0, float(want.max())))
    with pytest.raises(ValueError):
        img.rotate(10, mode="wrap")
