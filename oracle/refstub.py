"""TEST INFRASTRUCTURE ONLY -- never imported by the product (pylinac_b200/).

Import the *unmodified* reference (``/root/reference/pylinac``) in a container
that lacks its non-numerical dependencies (matplotlib, plotly, reportlab,
quaac, pydicom, skimage, argue, py_linq).  Everything numerical the PF /
Starshot / FieldAnalysis paths execute (numpy, scipy) is the real thing; only
presentation / file-format modules are replaced by inert stand-ins.

Used by ``tests/golden/make_*.py`` to generate the committed golden vectors.
``/root/reference`` does not exist on the GPU box, so nothing that runs there
may call :func:`import_reference`.
"""
from __future__ import annotations

import importlib.abc
import importlib.machinery
import io
import statistics
import sys
import types

import numpy as np

REFERENCE_ROOT = "/root/reference"

_AUTO_STUB_ROOTS = {
    "matplotlib",
    "mpl_toolkits",
    "plotly",
    "reportlab",
    "quaac",
    "pydicom",
    "skimage",
    "parameterized",
}


class _AutoAttr:
    """Inert object: any attribute / call / subscript returns another inert object."""

    def __init__(self, name="stub"):
        self.__name__ = name

    def __getattr__(self, item):
        if item.startswith("__") and item.endswith("__"):
            raise AttributeError(item)
        return _AutoAttr(f"{self.__name__}.{item}")

    def __call__(self, *a, **k):
        return _AutoAttr(self.__name__ + "()")

    def __getitem__(self, item):
        return _AutoAttr(self.__name__ + "[]")

    def __iter__(self):
        return iter(())

    def __mro_entries__(self, bases):
        return (object,)

    def __or__(self, other):
        return self

    def __ror__(self, other):
        return self


class _StubModule(types.ModuleType):
    __path__: list = []

    def __getattr__(self, item):
        if item.startswith("__") and item.endswith("__"):
            raise AttributeError(item)
        val = _make_stub_class(item) if item[:1].isupper() else _AutoAttr(f"{self.__name__}.{item}")
        setattr(self, item, val)
        return val


def _make_stub_class(name):
    # A real class so that it can be used in ``except``, ``isinstance``, as a base
    # class and in annotations.
    return type(name, (Exception,) if name.endswith("Error") else (object,), {"__init__": lambda self, *a, **k: None})


class _StubFinder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    def find_spec(self, fullname, path, target=None):
        if fullname.split(".")[0] in _AUTO_STUB_ROOTS:
            return importlib.machinery.ModuleSpec(fullname, self, is_package=True)
        return None

    def create_module(self, spec):
        m = _StubModule(spec.name)
        m.__path__ = []
        return m

    def exec_module(self, module):
        pass


# ---------------------------------------------------------------- argue shim
def _argue_module():
    m = types.ModuleType("argue")

    def _noop_decorator(*a, **k):
        def deco(f):
            return f

        return deco

    m.bounds = _noop_decorator
    m.options = _noop_decorator
    m.verify_bounds = lambda *a, **k: None
    m.POSITIVE = (0, float("inf"))

    class BoundsError(ValueError):
        pass

    m.BoundsError = BoundsError
    return m


# -------------------------------------------------------------- py_linq shim
class _Grouping(list):
    def __init__(self, key, items):
        super().__init__(items)
        self.key = key


class Enumerable:
    """Minimal py_linq.Enumerable (pyproject pins py_linq~=1.4) -- only the calls
    pylinac/picketfence.py makes (:448-539, :811-823, :1886-1894)."""

    def __init__(self, data=None):
        self._data = list(data) if data is not None else []

    def __iter__(self):
        return iter(self._data)

    def __len__(self):
        return len(self._data)

    def select(self, f):
        return Enumerable(f(x) for x in self._data)

    def select_many(self, f=lambda x: x):
        out = []
        for x in self._data:
            out.extend(list(f(x)))
        return Enumerable(out)

    def where(self, f):
        return Enumerable(x for x in self._data if f(x))

    def count(self, f=None):
        if f is None:
            return len(self._data)
        return sum(1 for x in self._data if f(x))

    def to_list(self):
        return list(self._data)

    def first(self):
        return self._data[0]

    def single(self):
        assert len(self._data) == 1
        return self._data[0]

    def distinct(self, key=lambda x: x):
        seen, out = set(), []
        for x in self._data:
            k = key(x)
            if k not in seen:
                seen.add(k)
                out.append(x)
        return Enumerable(out)

    def order_by_descending(self, key):
        # py_linq sorts with Python's stable sort (reverse=True keeps original order of ties)
        return Enumerable(sorted(self._data, key=key, reverse=True))

    def order_by(self, key):
        return Enumerable(sorted(self._data, key=key))

    def group_by(self, key_names=None, key=lambda x: x, result_func=None):
        groups: dict = {}
        for x in self._data:
            groups.setdefault(key(x), []).append(x)
        return Enumerable(_Grouping(k, v) for k, v in groups.items())

    def median(self, f=lambda x: x):
        return statistics.median([f(x) for x in self._data])

    def max(self, f=lambda x: x):
        return max(f(x) for x in self._data)

    def min(self, f=lambda x: x):
        return min(f(x) for x in self._data)


def _py_linq_module():
    m = types.ModuleType("py_linq")
    m.Enumerable = Enumerable
    return m


_installed = False


def import_reference():
    """Install the stubs, put /root/reference on sys.path, return the ``pylinac`` package."""
    global _installed
    if not _installed:
        sys.meta_path.insert(0, _StubFinder())
        sys.modules["argue"] = _argue_module()
        sys.modules["py_linq"] = _py_linq_module()
        if REFERENCE_ROOT not in sys.path:
            sys.path.insert(0, REFERENCE_ROOT)
        _installed = True
    import pylinac  # noqa

    return pylinac


# ------------------------------------------------------- array -> reference image
class FakeDicomDataset:
    """What DicomImage/LinacDicomImage read from a pydicom Dataset
    (core/image.py:1430-1444, 1509-1547, 1612-1730), backed by an ndarray."""

    def __init__(self, array, pixel_spacing_mm, sid=1000.0, sad=1000.0, gantry=0.0, coll=0.0, couch=0.0):
        self.pixel_array = np.asarray(array)
        self.ImagePlanePixelSpacing = [pixel_spacing_mm, pixel_spacing_mm]
        self.RTImageSID = sid
        self.RadiationMachineSAD = sad
        self.GantryAngle = gantry
        self.BeamLimitingDeviceAngle = coll
        self.PatientSupportAngle = couch
        self.Rows, self.Columns = self.pixel_array.shape
        self.file_meta = types.SimpleNamespace(TransferSyntaxUID="1.2.840.10008.1.2")

    def get(self, name, default=None):
        return self.__dict__.get(name, default)

    def __contains__(self, item):
        return item in self.__dict__


def reference_image_from_array(cls, array, pixel_spacing_mm, sid=1000.0, gantry=0.0, coll=0.0, couch=0.0, **kwargs):
    """Build ``cls`` (a DicomImage subclass of the reference) from an ndarray by
    swapping ``retrieve_dicom_file`` (core/io.py:73-84) for an in-memory dataset
    and ``apply_rescale`` for identity (synthetic images carry no slope/intercept,
    core/array_utils.py:251-311)."""
    import_reference()
    from pylinac.core import image as rimage

    ds = FakeDicomDataset(array, pixel_spacing_mm, sid=sid, gantry=gantry, coll=coll, couch=couch)
    old_retrieve, old_pixels = rimage.retrieve_dicom_file, rimage.pixels
    rimage.retrieve_dicom_file = lambda path: ds
    rimage.pixels = types.SimpleNamespace(apply_rescale=lambda arr, md: arr)
    try:
        return cls(io.BytesIO(b"fake"), **kwargs)
    finally:
        rimage.retrieve_dicom_file = old_retrieve
        rimage.pixels = old_pixels


def read_dicom_pixel_tail(path, rows, cols, dtype=np.uint16):
    """Uncompressed little-endian DICOM: PixelData is the last element, so the
    frame is the last rows*cols*itemsize bytes."""
    n = rows * cols * np.dtype(dtype).itemsize
    with open(path, "rb") as f:
        f.seek(-n, 2)
        return np.frombuffer(f.read(n), dtype=dtype).reshape(rows, cols).copy()
