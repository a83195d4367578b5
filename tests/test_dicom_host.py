"""Host-side ingest: pylinac_b200.dicom + DicomImage / LinacDicomImage / FileImage / load().

The reference reads files with pydicom (core/io.py:73-84) and then touches a handful of tags (core/image.py:363-389, 1383-1444,
1509-1580, 1612-1730).  These tests read the reference's OWN DICOM fixtures (docs/source/files/*.dcm, committed lzma-compressed)
through this repo's parser, and synthetic files with the tags clinical EPID images carry (rescale, sign flip, axis angles)."""
import io

import numpy as np
import pytest

from pylinac_b200 import dicom
from pylinac_b200.core import image
from tests.dicom_writer import write_dicom
from tests.golden import pf_docs_cases as dc


@pytest.mark.parametrize("name", list(dc.DOCS))
def test_reference_docs_files_parse(name, tmp_path):
    data = dc.docs_dcm_bytes(name)
    tail, ps, sid, _ = dc.docs_frame(name)
    ds = dicom.dcmread(data)
    assert (ds.Rows, ds.Columns, ds.BitsAllocated) == (1280, 1280, 16)
    assert ds.pixel_array.dtype == np.uint16 and np.array_equal(ds.pixel_array, tail)
    assert ds.ImagePlanePixelSpacing == [ps, ps] and ds.RTImageSID == sid and ds.RadiationMachineSAD == 1000.0
    assert ds.Modality == "RTIMAGE" and ds.TransferSyntaxUID == "1.2.840.10008.1.2"
    # file path, stream and bytes are all accepted (core/image.py:244-286, io.py:73-84)
    p = tmp_path / (name + ".dcm")
    p.write_bytes(data)
    img = image.load(str(p))
    assert isinstance(img, image.DicomImage)
    assert np.array_equal(image.load(io.BytesIO(data)).array, img.array)
    # RescaleSlope 1 / RescaleIntercept 0 are present: pydicom's apply_rescale yields float64 (core/image.py:374)
    assert img.array.dtype == np.float64 and np.array_equal(img.array, tail.astype(np.float64))
    assert img.dpmm == pytest.approx(1 / ps * sid / 1000.0) and img.sid == sid and img.sad == 1000.0
    assert (img.cax.x, img.cax.y) == (639.5, 639.5)
    lin = image.LinacDicomImage(str(p))
    assert (lin.gantry_angle, lin.collimator_angle, lin.couch_angle) == (0.0, 0.0, 0.0)
    f16 = image.frame_u16(img)
    assert f16.dtype == np.uint16 and np.array_equal(f16, tail)


@pytest.mark.parametrize("explicit,preamble", [(True, True), (False, True), (False, False)])
def test_synthetic_dicom_tags(tmp_path, explicit, preamble):
    rng = np.random.default_rng(5)
    a = rng.integers(0, 60000, (48, 64)).astype(np.uint16)
    p = write_dicom(tmp_path / "x.dcm", a, pixel_spacing_mm=0.392, sid=1500.0, gantry=270.04, coll=12.0, couch=359.96, slope=2.5,
                    intercept=-100.0, explicit=explicit, preamble=preamble, translation=[1.0, -2.0, -500.0])
    assert dicom.is_dicom(p)
    img = image.LinacDicomImage(p, axes_precision=1)
    assert img.shape == (48, 64)
    np.testing.assert_array_equal(img.array, a.astype(np.float64) * 2.5 - 100.0)      # pydicom apply_rescale
    assert img.dpmm == pytest.approx(1 / 0.392 * 1.5)
    assert (img.gantry_angle, img.collimator_angle, img.couch_angle) == (270.0, 12.0, 0.0)   # 359.96 -> 360.0 -> 0 (core/image.py:1700-1730)
    assert img.cax.x == pytest.approx(img.center.x - 1.0 * img.dpmm / 1.5) and img.cax.y == pytest.approx(img.center.y - 2.0 * img.dpmm / 1.5)
    # overrides win over tags
    assert image.LinacDicomImage(p, gantry=10.0).gantry_angle == 10.0
    # raw pixels: no rescale
    raw = image.DicomImage(p, raw_pixels=True)
    assert raw.array.dtype == np.uint16 and np.array_equal(raw.array, a)
    # the device pipelines take the stored integers of an untouched rescaled image
    assert np.array_equal(image.frame_u16(img), a)
    img.array = img.array * 1.0001   # no longer the rescale of the stored values and not integer valued
    with pytest.raises(ValueError):
        image.frame_u16(img)


def test_sign_flip_and_forced_inversion(tmp_path):
    a = np.arange(30 * 40, dtype=np.uint16).reshape(30, 40) * 7 + 11
    p = write_dicom(tmp_path / "neg.dcm", a, sign=-1, slope=1.0, intercept=0.0)
    img = image.DicomImage(p)
    exp = a.astype(np.float64)
    exp = exp.max() - exp + exp.min()
    np.testing.assert_array_equal(img.array, exp)                   # core/image.py:381-388
    np.testing.assert_array_equal(image.DicomImage(p, invert_pixels=False).array, a.astype(np.float64))
    f = image.frame_u16(img)
    assert f.dtype == np.uint16 and np.array_equal(f, (int(a.max()) + int(a.min()) - a.astype(np.int64)).astype(np.uint16))
    # without rescale tags the dtype is preserved
    p2 = write_dicom(tmp_path / "plain.dcm", a)
    assert image.DicomImage(p2).array.dtype == np.uint16


def test_load_dispatch_and_errors(tmp_path):
    a = (np.random.default_rng(1).random((20, 30)) * 1000).astype(np.uint16)
    assert isinstance(image.load(a, dpi=100), image.ArrayImage)
    with pytest.raises(TypeError):
        image.load(str(tmp_path))                                    # a directory: neither DICOM, image nor array
    with pytest.raises(FileExistsError):
        image.DicomImage(str(tmp_path / "missing.dcm"))              # core/image.py:483
    bad = tmp_path / "bad.dcm"
    bad.write_bytes(b"\x00" * 256)
    assert not dicom.is_dicom(str(bad))
    # TIFF / PNG through Pillow -> FileImage (core/image.py:1733-1812)
    from PIL import Image as pImage

    tif = tmp_path / "f.tif"
    pImage.fromarray(a).save(tif, dpi=(72, 72))
    img = image.load(str(tif), sid=1500)
    assert isinstance(img, image.FileImage) and np.array_equal(img.array, a)
    assert img.dpi == pytest.approx(72 * 1.5) and img.dpmm == pytest.approx(72 * 1.5 / 25.4)
    png = tmp_path / "f.png"
    pImage.fromarray(np.stack([a >> 8] * 3, axis=-1).astype(np.uint8)).save(png)
    rgb = image.load(str(png), dpi=50)
    assert rgb.array.ndim == 2 and rgb.dpi == 50


def test_winston_lutz_2d_reads_axis_tags(tmp_path):
    """WinstonLutz2D(path) must be a LinacDicomImage like the reference's (winston_lutz.py:629, 1137)."""
    from pylinac_b200 import winston_lutz as wl

    a = np.zeros((64, 64), np.uint16)
    p = write_dicom(tmp_path / "wl.dcm", a, gantry=90.0, coll=0.0, couch=45.0, slope=1.0, intercept=0.0)
    w = wl.WinstonLutz2D(p)
    assert (w.gantry_angle, w.collimator_angle, w.couch_angle) == (90.0, 0.0, 45.0)
    w = wl.WinstonLutz2D(p, gantry=180.0)
    assert (w.gantry_angle, w.couch_angle) == (180.0, 45.0)
    assert w._frame_u16().dtype == np.uint16          # identity rescale -> float64 array -> stored integers


def test_axis_values_from_file_names(tmp_path):
    """LinacDicomImage._get_axis_value (core/image.py:1655-1730): with use_filenames the `<axis><number>` convention wins over the tags, an
    absent keyword gives missing_axis_value (tags are not consulted), a keyword without a number raises (expected values derived by hand from that function; pydicom is not
    installed here, so the reference cannot read the file itself)."""
    from pylinac_b200.core import image
    from pylinac_b200 import winston_lutz as wl

    a = np.zeros((64, 64), np.uint16)
    p = write_dicom(tmp_path / "wl_gantry45_COLL270.dcm", a, gantry=90.0, coll=0.0, couch=45.0)
    img = image.LinacDicomImage(p, use_filenames=True)
    assert (img.gantry_angle, img.collimator_angle, img.couch_angle) == (45.0, 270.0, 0.0)
    img = image.LinacDicomImage(p, use_filenames=True, missing_axis_value=7)
    assert img.couch_angle == 7
    with pytest.raises(ValueError, match="not found in the filename"):
        image.LinacDicomImage(p, use_filenames=True, missing_axis_value="raise").couch_angle
    assert image.LinacDicomImage(p, use_filenames=True, couch=12).couch_angle == 12.0      # explicit values first
    assert image.LinacDicomImage(p).gantry_angle == 90.0                                    # tags otherwise
    bad = write_dicom(tmp_path / "wl_gantry_x.dcm", a, gantry=90.0, coll=0.0, couch=45.0)
    with pytest.raises(ValueError, match="could not read a number"):
        image.LinacDicomImage(bad, use_filenames=True).gantry_angle
    w = wl.WinstonLutz2D(p, use_filenames=True)
    assert (w.gantry_angle, w.collimator_angle, w.couch_angle) == (45.0, 270.0, 0.0)


def test_batched_ingest_reads_pixels_straight_into_one_array(tmp_path):
    """dicom.read_frames: header-only parse + readinto of the pixel bytes into the caller's [n, rows, cols] array; equals the
    per-file reader, for explicit / implicit VR files with and without preamble; mismatching shapes are refused."""
    rng = np.random.default_rng(3)
    arrs = [rng.integers(0, 65535, (48, 40)).astype(np.uint16) for _ in range(5)]
    paths = [write_dicom(tmp_path / f"f{i}.dcm", a, explicit=bool(i % 2), preamble=i != 3, slope=1.0 if i == 2 else None,
                         intercept=0.0 if i == 2 else None) for i, a in enumerate(arrs)]
    out = np.zeros((5, 48, 40), np.uint16)
    frames, headers = dicom.read_frames(paths, out=out, threads=3)
    assert frames is out
    for i, a in enumerate(arrs):
        np.testing.assert_array_equal(frames[i], a)
        np.testing.assert_array_equal(dicom.dcmread(paths[i]).pixel_array, a)
        assert headers[i]["Rows"] == 48 and headers[i]["PixelCount"] == 48 * 40
        assert "pixel_array" not in headers[i].__dict__
    frames2, _ = dicom.read_frames(paths)
    np.testing.assert_array_equal(frames2, out)
    other = write_dicom(tmp_path / "other.dcm", np.zeros((32, 40), np.uint16))
    with pytest.raises(ValueError, match="differs from the first"):
        dicom.read_frames(paths + [other])
    with pytest.raises(ValueError, match="out must be"):
        dicom.read_frames(paths, out=np.zeros((5, 48, 41), np.uint16))
    # a header longer than the first read is parsed from the whole file
    h = dicom.read_header(paths[0], head_bytes=64)
    assert h["PixelCount"] == 48 * 40


def test_zip_archives_and_image_file_discovery(tmp_path):
    """core/io.py TemporaryZipDirectory + core/image.py retrieve_image_files, and WinstonLutz.from_zip (winston_lutz.py:1397-1410): the
    set is built from the DICOM files of the archive (frames and axis angles are read before the temporary directory is removed)."""
    import zipfile

    from pylinac_b200 import winston_lutz as wl

    rng = np.random.default_rng(8)
    d = tmp_path / "imgs"
    d.mkdir()
    frames, axes = [], [(0.0, 0.0, 0.0), (90.0, 0.0, 0.0), (180.0, 0.0, 0.0)]
    for i, (g, c, p) in enumerate(axes):
        a = rng.integers(100, 60000, (32, 32)).astype(np.uint16)
        frames.append(a)
        write_dicom(d / f"wl{i}.dcm", a, gantry=g, coll=c, couch=p)
    (d / "notes.txt").write_text("not an image")
    z = tmp_path / "set.zip"
    with zipfile.ZipFile(z, "w") as zf:
        for f in sorted(d.iterdir()):
            zf.write(f, arcname=f"sub/{f.name}")
    with image.TemporaryZipDirectory(str(z)) as tmp:
        files = image.retrieve_image_files(tmp)
        assert [f.rsplit("/", 1)[1] for f in files] == ["wl0.dcm", "wl1.dcm", "wl2.dcm"]
    import os
    assert not os.path.exists(tmp)
    st = wl.WinstonLutz.from_zip(str(z))
    assert st._axes == axes and st._frames.shape == (3, 32, 32)
    for k in range(3):
        np.testing.assert_array_equal(st._frames[k], frames[k])
    with open(z, "rb") as fh:                       # binary stream, like the reference accepts
        assert wl.WinstonLutz.from_zip(fh)._frames.shape == (3, 32, 32)


def test_picketfence_from_bb_setup_overrides_the_cax(monkeypatch):
    """picketfence.py:402-437: the BB's offset from the centre of the BB image, in mm, becomes the central-axis override of the
    picket-fence image (the locator itself is a GPU metric: stubbed here; bright BB first, dark BB when that raises)."""
    from pylinac_b200 import picketfence as pf
    from pylinac_b200.core.geometry import Point

    bb = image.ArrayImage(np.zeros((200, 300), np.uint16), dpi=25.4 * 2.0, sid=1000)        # dpmm = 2
    calls = []

    def fake_compute(self, metrics):
        calls.append(bool(metrics.invert))
        if metrics.invert:
            raise ValueError("no bright BB")
        return [Point(x=149.5 + 10.0, y=99.5 - 4.0)]

    monkeypatch.setattr(image.BaseImage, "compute", fake_compute)
    frame = np.zeros((64, 64), np.uint16)
    inst = pf.PicketFence.from_bb_setup(frame, bb_image=bb, bb_diameter=5, image_kwargs={"dpi": 25.4 * 2.56, "sid": 1000})
    assert calls == [True, False]
    assert inst._from_bb_setup and inst._bb_image is bb
    assert (inst._central_axis.x, inst._central_axis.y) == (5.0, -2.0)
