"""Multi-target multi-field Winston-Lutz end to end on the GPU (whole-frame field locator + windowed BB searches per image) against
goldens of the UNMODIFIED reference (tests/golden/make_mtmf_golden.py; skimage restated by oracle/skimage_shim.py, unpinned there).
Bar: BB / field positions <= 0.01 px."""
import warnings

import numpy as np
import pytest

from tests.golden.mtmf_cases import SETS, set_frames

pytestmark = pytest.mark.gpu
GOLD = np.load("tests/golden/mtmf_golden.npz")


@pytest.mark.parametrize("name", list(SETS))
def test_mtmf_matches_reference_golden(name):
    from pylinac_b200 import winston_lutz_mtmf as mt
    from tests.test_mtmf_host import check

    frames, ps, sid, axes, arr = set_frames(name)
    cfgs = tuple(mt.BBConfig(name=n, offset_left_mm=l, offset_up_mm=u, offset_in_mm=i, bb_size_mm=d, rad_size_mm=r) for n, l, u, i, d, r in arr)
    st = mt.WinstonLutzMultiTargetMultiField.from_arrays(frames, axes, dpmm=(1 / ps) * sid / 1000.0)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        st.analyze(bb_arrangement=cfgs)
    names = list(GOLD[f"{name}/names"])
    for k, img in enumerate(st.images):
        assert tuple(img.shape) == tuple(GOLD[f"{name}/shape"][k])
        np.testing.assert_allclose([img.epid.x, img.epid.y], GOLD[f"{name}/epid_px"][k], atol=1e-9)
        np.testing.assert_allclose([[img.arrangement_matches[n].field.x, img.arrangement_matches[n].field.y] for n in names],
                                   GOLD[f"{name}/field_px"][k], rtol=0, atol=1e-6)
        np.testing.assert_allclose([[img.arrangement_matches[n].bb.x, img.arrangement_matches[n].bb.y] for n in names],
                                   GOLD[f"{name}/bb_px"][k], rtol=0, atol=1e-6)
    check(st, name, 1e-6)
