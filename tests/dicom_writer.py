"""Test helper: write small uncompressed DICOM RT-Image files (what the reference's ``array_to_dicom`` + pydicom would write,
core/array_utils.py:216-311) without pydicom, so the ingest path (pylinac_b200/dicom.py, DicomImage, LinacDicomImage) can be tested
with arbitrary tags: rescale slope / intercept, PixelIntensityRelationshipSign, axis angles, explicit or implicit VR."""
from __future__ import annotations

import struct

import numpy as np

_LONG = {b"OB", b"OW", b"SQ", b"UN", b"UT"}


def _ds(v) -> bytes:
    if isinstance(v, (list, tuple)):
        txt = "\\".join(repr(float(x)) for x in v)
    else:
        txt = repr(float(v))
    b = txt.encode("ascii")
    return b + b" " * (len(b) & 1)


def _str(v: str) -> bytes:
    b = v.encode("ascii")
    return b + b" " * (len(b) & 1)


def _uid(v: str) -> bytes:
    b = v.encode("ascii")
    return b + b"\x00" * (len(b) & 1)


def _elem(group, elem, vr: bytes, value: bytes, explicit: bool) -> bytes:
    if explicit:
        if vr in _LONG:
            return struct.pack("<HH2sHI", group, elem, vr, 0, len(value)) + value
        return struct.pack("<HH2sH", group, elem, vr, len(value)) + value
    return struct.pack("<HHI", group, elem, len(value)) + value


def write_dicom(path, array: np.ndarray, *, pixel_spacing_mm=0.4, sid=1000.0, sad=1000.0, gantry=None, coll=None, couch=None,
                slope=None, intercept=None, sign=None, explicit=True, preamble=True, translation=None, use_image_plane_tag=True):
    """array: 2-D uint16 (or uint8).  Returns the path."""
    a = np.ascontiguousarray(array)
    assert a.ndim == 2 and a.dtype in (np.uint16, np.uint8)
    bits = 8 * a.dtype.itemsize
    ex = explicit
    body = []
    body.append(_elem(0x0008, 0x0016, b"UI", _uid("1.2.840.10008.5.1.4.1.1.481.1"), ex))
    body.append(_elem(0x0008, 0x0060, b"CS", _str("RTIMAGE"), ex))
    body.append(_elem(0x0008, 0x0070, b"LO", _str("pylinac_b200 tests"), ex))
    body.append(_elem(0x0018, 0x1110, b"DS", _ds(sid), ex))
    body.append(_elem(0x0028, 0x0002, b"US", struct.pack("<H", 1), ex))
    body.append(_elem(0x0028, 0x0010, b"US", struct.pack("<H", a.shape[0]), ex))
    body.append(_elem(0x0028, 0x0011, b"US", struct.pack("<H", a.shape[1]), ex))
    if not use_image_plane_tag:
        body.append(_elem(0x0028, 0x0030, b"DS", _ds([pixel_spacing_mm, pixel_spacing_mm]), ex))
    body.append(_elem(0x0028, 0x0100, b"US", struct.pack("<H", bits), ex))
    body.append(_elem(0x0028, 0x0101, b"US", struct.pack("<H", bits), ex))
    body.append(_elem(0x0028, 0x0103, b"US", struct.pack("<H", 0), ex))
    if sign is not None:
        body.append(_elem(0x0028, 0x1041, b"SS", struct.pack("<h", int(sign)), ex))
    if intercept is not None:
        body.append(_elem(0x0028, 0x1052, b"DS", _ds(intercept), ex))
    if slope is not None:
        body.append(_elem(0x0028, 0x1053, b"DS", _ds(slope), ex))
    if translation is not None:
        body.append(_elem(0x3002, 0x000D, b"DS", _ds(translation), ex))
    if use_image_plane_tag:
        body.append(_elem(0x3002, 0x0011, b"DS", _ds([pixel_spacing_mm, pixel_spacing_mm]), ex))
    body.append(_elem(0x3002, 0x0022, b"DS", _ds(sad), ex))
    body.append(_elem(0x3002, 0x0026, b"DS", _ds(sid), ex))
    if gantry is not None:
        body.append(_elem(0x300A, 0x011E, b"DS", _ds(gantry), ex))
    if coll is not None:
        body.append(_elem(0x300A, 0x0120, b"DS", _ds(coll), ex))
    if couch is not None:
        body.append(_elem(0x300A, 0x0122, b"DS", _ds(couch), ex))
    body.append(_elem(0x7FE0, 0x0010, b"OW", a.astype(a.dtype.newbyteorder("<")).tobytes(), ex))
    out = b""
    if preamble:
        ts = "1.2.840.10008.1.2.1" if explicit else "1.2.840.10008.1.2"
        meta = _elem(0x0002, 0x0002, b"UI", _uid("1.2.840.10008.5.1.4.1.1.481.1"), True) + _elem(0x0002, 0x0010, b"UI", _uid(ts), True)
        meta = _elem(0x0002, 0x0000, b"UL", struct.pack("<I", len(meta)), True) + meta
        out = b"\x00" * 128 + b"DICM" + meta
    out += b"".join(body)
    with open(path, "wb") as f:
        f.write(out)
    return str(path)
