"""The restated image generator (oracle/synth.py, test infrastructure) against the pieces of the reference's generator that run here
without scikit-image: the BB projection of generate_winstonlutz (winston_lutz.py:3401-3460)."""
import os

import numpy as np
import pytest

from oracle import synth

pytestmark = pytest.mark.skipif(not os.path.isdir("/root/reference/pylinac"), reason="needs the reference tree")


def test_bb_projection_matches_the_reference_function():
    from oracle.refstub import import_reference

    import_reference()
    from pylinac.winston_lutz import bb_projection_with_rotation as ref

    rng = np.random.default_rng(0)
    for _ in range(300):
        left, up, inn = rng.uniform(-5, 5, 3)
        g, c = rng.uniform(0, 360, 2)
        a, b = ref(left, up, inn, g, c), synth.bb_projection_with_rotation(left, up, inn, g, c)
        assert abs(a[0] - b[0]) < 1e-12 and abs(a[1] - b[1]) < 1e-12
