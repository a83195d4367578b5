"""Generate tests/golden/combine_golden.npz with the UNMODIFIED reference (stub-imported from /root/reference) and scipy:

* ``load_multiples`` (core/image.py:306-360) on ndarray inputs, every method x stretch_each;
* ``convert_to_dtype`` (core/array_utils.py:172-198) float -> uint16 / uint8 -> uint16 / uint16 -> uint8;
* ``equate_images`` (core/image.py:169-220);
* ``scipy.ndimage.zoom`` (what ``equate_images`` and ``ProfileBase.as_resampled`` call) for the cases of combine_cases.ZOOM_CASES;
* ``PicketFence.from_multiple_images`` (picketfence.py:357-400) and ``Starshot.from_multiple_images`` (starshot.py:148-174) end to
  end: the in-memory DICOM the reference writes and re-reads is intercepted at ``Dataset.save_as`` / ``retrieve_dicom_file``
  (pydicom is not installed here), so the stored pixel array the reference analyses is recorded together with its results.

Run here:  python -m tests.golden.make_combine_golden
"""
from __future__ import annotations

import hashlib
import io
import types
import warnings

import numpy as np

from oracle.refstub import FakeDicomDataset, import_reference
from tests.golden import combine_cases as cc


class _SavingDataset(FakeDicomDataset):
    """FakeDicomDataset whose ``save_as`` keeps what DicomImage.save wrote (core/image.py:1485-1488)."""

    saved: dict = {}

    def save_as(self, target):
        px = np.frombuffer(self.PixelData, dtype=self.pixel_array.dtype).reshape(self.Rows, self.Columns).copy()
        _SavingDataset.saved[id(target)] = (px, self)


def _sha(a):
    return np.frombuffer(hashlib.sha1(np.ascontiguousarray(a).tobytes()).digest(), dtype=np.uint8)


def _with_fake_dicom(parts, ps, sid, fn):
    """Run ``fn(list_of_streams)`` with retrieve_dicom_file resolving each stream to an in-memory dataset of ``parts`` and any
    OTHER stream to the dataset last written by ``save_as`` (the reference's write-then-read of the composite)."""
    from pylinac.core import image as rimage

    streams = [io.BytesIO(b"part%d" % k) for k in range(len(parts))]
    table = {id(s): _SavingDataset(p, ps, sid=sid) for s, p in zip(streams, parts)}

    def retrieve(path):
        if id(path) in table:
            return table[id(path)]
        px, src = _SavingDataset.saved[id(path)]
        ds = _SavingDataset(px, ps, sid=sid)
        return ds

    old = rimage.retrieve_dicom_file, rimage.pixels
    rimage.retrieve_dicom_file = retrieve
    rimage.pixels = types.SimpleNamespace(apply_rescale=lambda arr, md: arr)
    try:
        return fn(streams)
    finally:
        rimage.retrieve_dicom_file, rimage.pixels = old


def main():
    import_reference()
    from scipy import ndimage

    from pylinac import picketfence as rpf
    from pylinac import starshot as rstar
    from pylinac.core import array_utils as rau
    from pylinac.core import image as rimage

    warnings.simplefilter("ignore")
    store = {}
    # ---- load_multiples on arrays
    stack = cc.small_stack()
    for method in ("mean", "max", "sum"):
        for stretch in (True, False):
            img = rimage.load_multiples([a.copy() for a in stack], method=method, stretch_each=stretch, dpi=100, sid=1000)
            store[f"lm/{method}/{int(stretch)}"] = np.asarray(img.array)
    img = rimage.load_multiples([a.copy() for a in stack[:2]], method="mean", stretch_each=True, dtype=np.uint16, dpi=100, sid=1000)
    store["lm/mean/u16"] = np.asarray(img.array)
    # ---- convert_to_dtype
    store["ctd/f_u16"] = rau.convert_to_dtype(stack[2], np.uint16)
    store["ctd/u8_u16"] = rau.convert_to_dtype((stack[0] >> 4).astype(np.uint8), np.uint16)
    store["ctd/u16_u8"] = rau.convert_to_dtype(stack[1], np.uint8)
    # ---- zoom
    for name, (shape, z, order, mode) in cc.ZOOM_CASES.items():
        out = ndimage.zoom(cc.zoom_input(name), z, order=order, mode=mode)
        store[f"zoom/{name}/shape"] = np.array(out.shape)
        store[f"zoom/{name}"] = out[::3, ::3] if name == "z2d_big" else out
    # ---- equate_images
    (a, adpi), (b, bdpi) = cc.equate_inputs()
    i1, i2 = rimage.equate_images(rimage.load(a, dpi=adpi, sid=1000), rimage.load(b, dpi=bdpi, sid=1000))
    store["eq/a_sha1"], store["eq/b"] = _sha(np.asarray(i1.array)), np.asarray(i2.array)[::2, ::2]
    store["eq/shapes"] = np.array([i1.shape, i2.shape])
    store["eq/dpi"] = np.array([i1.dpi, i2.dpi])
    i1, i2 = rimage.equate_images(rimage.load(b, dpi=bdpi, sid=1000), rimage.load(a, dpi=adpi, sid=1000))
    store["eq2/a_sha1"], store["eq2/b"] = _sha(np.asarray(i1.array)), np.asarray(i2.array)[::2, ::2]
    store["eq2/shapes"] = np.array([i1.shape, i2.shape])
    store["eq2/dpi"] = np.array([i1.dpi, i2.dpi])

    # ---- PicketFence.from_multiple_images
    def run_pf(streams, **kw):
        pf = rpf.PicketFence.from_multiple_images(streams, **kw)
        pf.analyze()
        return pf

    for tag, kw in (("mean", {}), ("sum_nostretch", {"method": "sum", "stretch_each": False})):
        pf = _with_fake_dicom(cc.pf_parts(), cc.PS, cc.SID, lambda s: run_pf(s, **kw))
        rd = pf.results_data()
        px = [v[0] for v in _SavingDataset.saved.values()][-1]
        store[f"pf/{tag}/stored_sha1"] = _sha(px)
        store[f"pf/{tag}/stored_sub"] = px[::8, ::8]
        store[f"pf/{tag}/shape"] = np.array(pf.image.shape)
        store[f"pf/{tag}/position"] = np.array([list(m.position) for m in pf.mlc_meas], dtype=np.float64)
        store[f"pf/{tag}/error"] = np.array([list(m.error) for m in pf.mlc_meas], dtype=np.float64)
        store[f"pf/{tag}/leaf"] = np.array([m.leaf_num for m in pf.mlc_meas])
        store[f"pf/{tag}/picket"] = np.array([m.picket_num for m in pf.mlc_meas])
        store[f"pf/{tag}/max_error"] = np.array(rd.max_error_mm)
        store[f"pf/{tag}/abs_median_error"] = np.array(rd.absolute_median_error_mm)
        store[f"pf/{tag}/number_of_pickets"] = np.array(rd.number_of_pickets)
        store[f"pf/{tag}/offsets"] = np.array(rd.offsets_from_cax_mm)
        print("pf", tag, rd.number_of_pickets, len(pf.mlc_meas), rd.max_error_mm, px.dtype, px.min(), px.max())
        _SavingDataset.saved.clear()

    # ---- Starshot.from_multiple_images (DICOM inputs; the generic loader resolves streams to DicomImage via is_dicom)
    def run_star(streams):
        old = rimage.is_dicom_image
        rimage.is_dicom_image = lambda file=None: True
        try:
            st = rstar.Starshot.from_multiple_images(streams)
        finally:
            rimage.is_dicom_image = old
        st.analyze()
        return st

    st = _with_fake_dicom(cc.star_parts(), cc.PS, cc.SID, run_star)
    px = [v[0] for v in _SavingDataset.saved.values()][-1]
    store["star/stored_sha1"] = _sha(px)
    store["star/stored_sub"] = px[::8, ::8]
    store["star/wobble"] = np.array([st.wobble.center.x, st.wobble.center.y, st.wobble.radius, st.wobble.radius_mm])
    store["star/circle"] = np.array([st.circle_profile.center.x, st.circle_profile.center.y, st.circle_profile.radius])
    store["star/npeaks"] = np.array(len(st.lines))
    print("star", store["star/wobble"], px.dtype, px.min(), px.max())
    np.savez_compressed("tests/golden/combine_golden.npz", **store)


if __name__ == "__main__":
    main()
