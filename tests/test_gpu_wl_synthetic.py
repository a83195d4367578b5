"""The reference's OWN synthetic Winston-Lutz expectations (tests_basic/test_winstonlutz.py:1244-1520, tolerances of
``WinstonLutzMixin`` :1155-1206 and ``SyntheticWLMixin`` :1303-1347) reproduced against the CUDA pipeline.

These are the only offline-checkable statements the reference makes about the BB finder end to end (the skimage boundary itself is
not installed here): a 5 mm BB with a known 3-D offset, an AS1200 panel at SID 1000, a 20 x 20 mm perfect field blurred with 1.5 mm,
imaged at 8 gantry / couch positions, must yield the BB shift vector to 0.05 mm, the measured BB position to 0.03 mm and the
CAX-to-BB / CAX-to-EPID / couch-isocentre statistics of each test class to 0.1 - 0.15 mm.  Frames come from the restated generator
(oracle/synth.py, generate_winstonlutz recipe image_generator/utils.py:139-263)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from tests.golden.wl_synthetic_classes import CLASSES, _set


@pytest.mark.parametrize("name", list(CLASSES))
def test_reference_synthetic_wl_class(name):
    from pylinac_b200 import winston_lutz as wl

    left, up, inn, axes, exp = CLASSES[name]
    frames, dpmm = _set(left, up, inn, axes)
    st = wl.WinstonLutz.from_arrays(frames, axes, dpmm=dpmm)
    st.analyze(bb_size_mm=5)
    assert len(st.images) == len(axes)
    # SyntheticWLMixin.test_bb_shift_vector / test_bb3d_measured_position
    sv = st.bb_shift_vector
    assert abs(sv.x - left) < 0.05 and abs(sv.y - (-inn)) < 0.05 and abs(sv.z - (-up)) < 0.05, (sv.x, sv.y, sv.z)
    mp = st.measured_bb_position
    assert abs(mp.x - (-left)) < 0.03 and abs(mp.y - inn) < 0.03 and abs(mp.z - up) < 0.03, (mp.x, mp.y, mp.z)
    # WinstonLutzMixin
    assert abs(st.cax2bb_distance("max") - exp["bb_max"]) < 0.15
    assert abs(st.cax2bb_distance("median") - exp["bb_median"]) < 0.1
    assert abs(st.cax2bb_distance("mean") - exp["bb_mean"]) < 0.1
    assert abs(st.cax2epid_distance("max") - exp["epid_max"]) < 0.1
    if "couch_iso" in exp:
        assert abs(st.couch_iso_size - exp["couch_iso"]) < 0.15
    assert abs(st.gantry_iso_size - 0) < 0.15          # mixin defaults: gantry_iso_size = collimator_iso_size = 0
