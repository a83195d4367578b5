"""CPU: the C-ABI shared library loads and exports every symbol include/epid.h declares (no compute calls)."""
import ctypes
import os
import re

from pylinac_b200 import _native as nat

HEADER = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "include", "epid.h")


def declared_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(epid_[a-z0-9_]+)\s*\(", src)))


def test_header_functions_are_exported_and_bound():
    names = declared_functions()
    assert len(names) >= 40
    handle = ctypes.CDLL(nat.LIB_PATH)
    for n in names:
        assert hasattr(handle, n), f"libepid.so does not export {n}"
    bound = set(nat.exported_symbols())
    assert set(names) <= bound, sorted(set(names) - bound)
    assert bound <= set(names), sorted(bound - set(names))


def test_struct_layouts_match_the_numpy_dtypes():
    assert nat.PF_SUMMARY_DTYPE.itemsize == ctypes.sizeof(nat.PFSummary)
    assert nat.PF_MEAS_DTYPE.itemsize == ctypes.sizeof(nat.PFMeas)
    assert nat.STAR_RESULT_DTYPE.itemsize == ctypes.sizeof(nat.StarResult)
    for dt, st in ((nat.PF_SUMMARY_DTYPE, nat.PFSummary), (nat.PF_MEAS_DTYPE, nat.PFMeas), (nat.STAR_RESULT_DTYPE, nat.StarResult)):
        for name, *_ in st._fields_:
            assert dt.fields[name][1] == getattr(st, name).offset, name


def test_no_device_is_reported_not_faked():
    """Without a GPU every compute entry point must fail loudly (EPID_ERR_NO_DEVICE), never fall back to the CPU."""
    import numpy as np
    import pytest

    if nat.device_count() > 0:
        pytest.skip("a CUDA device is present")
    with pytest.raises(nat.NoDeviceError):
        nat.Context(0)
    from pylinac_b200 import picketfence as pf

    with pytest.raises(nat.NoDeviceError):
        pf.analyze_batch(np.zeros((1, 64, 64), np.uint16), 2.56)


def test_product_never_imports_the_oracle_and_has_one_scipy_call_site():
    """The oracle is test infrastructure: nothing under pylinac_b200/ may import it (a product path through the oracle would void
    every parity claim).  scipy is allowed in exactly one place: the set-level Winston-Lutz minimisation the reference itself does
    on the host (winston_lutz.py:1614-1640)."""
    import ast
    import pathlib

    root = pathlib.Path(__file__).resolve().parents[1] / "pylinac_b200"
    scipy_sites = []
    for path in root.rglob("*.py"):
        tree = ast.parse(path.read_text())
        for node in ast.walk(tree):
            names = []
            if isinstance(node, ast.Import):
                names = [a.name for a in node.names]
            elif isinstance(node, ast.ImportFrom) and node.module:
                names = [node.module]
            for nm in names:
                top = nm.split(".")[0]
                assert top != "oracle", f"{path} imports the oracle"
                # torch.distributed is rendezvous / barrier plumbing of the multi-GPU helpers only (parallel.py)
                banned = ("triton", "cupy", "numba") + (() if path.name == "parallel.py" else ("torch",))
                assert top not in banned, f"{path} imports {top}"
                if top == "scipy":
                    scipy_sites.append(path.name)
    # set-level scalar host work on RESULT rows, exactly the calls the reference makes there: optimize.minimize over the back-projection
    # segments (winston_lutz.py:1614-1640) and Rotation.as_euler for a non-default axes order of align_points (winston_lutz.py:3655-3658)
    assert sorted(scipy_sites) == ["winston_lutz.py", "winston_lutz_mtmf.py"], scipy_sites
