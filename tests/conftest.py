import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box via gpurun)")
    config.addinivalue_line("markers", "reference: needs /root/reference (only present in the build container)")


def _has_gpu():
    try:
        from pylinac_b200 import _native

        return _native.device_count() > 0
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    have_ref = os.path.isdir("/root/reference/pylinac")
    for item in items:
        if "reference" in item.keywords and not have_ref:
            item.add_marker(pytest.mark.skip(reason="/root/reference not present"))
