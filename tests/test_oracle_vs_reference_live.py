"""CPU, only where /root/reference exists (this container): the oracle restatements against the UNMODIFIED reference run LIVE
(stub-imported, oracle/refstub.py) on fresh seeded frames that are not part of the committed goldens -- the goldens pin fixed
cases, this guards against an oracle that only fits those.  Skipped on boxes without the reference tree."""
import os
import warnings

import numpy as np
import pytest

pytestmark = pytest.mark.skipif(not os.path.isdir("/root/reference/pylinac"), reason="/root/reference is not present on this box")


@pytest.mark.parametrize("seed", [901, 902, 903])
def test_pf_oracle_equals_reference_on_fresh_frames(seed):
    from oracle import pf_oracle, synth
    from tests.golden.refrun import reference_pf
    from tests.test_oracle_pf import CLOSE, EXACT

    rng = np.random.default_rng(seed)
    fr = synth.epid1024()
    a = synth.picketfence_frame(fr, pickets=int(rng.integers(5, 11)), picket_spacing_mm=int(rng.integers(18, 28)),
                                picket_width_mm=int(rng.integers(2, 5)), picket_offset_error=rng.uniform(-0.6, 0.6, 12),
                                noise_sigma=float(rng.uniform(0.001, 0.004)), seed=seed,
                                orientation="left_right" if seed % 2 else "up_down")
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        ref = reference_pf(a, fr.pixel_size, 1000.0, {}, {})
        o = pf_oracle.pf_analyze(a, 1 / fr.pixel_size)
    for k in EXACT:
        assert np.array_equal(np.asarray(o[k]), np.asarray(ref[k])), k
    for k in CLOSE:
        np.testing.assert_array_equal(np.asarray(o[k]), np.asarray(ref[k]), err_msg=k)


@pytest.mark.parametrize("seed", [911, 912])
def test_starshot_oracle_equals_reference_on_fresh_frames(seed):
    from oracle import starshot_oracle, synth
    from tests.golden.refrun import reference_starshot

    rng = np.random.default_rng(seed)
    spokes = int(rng.choice([4, 6, 8]))
    fr = synth.epid1024()
    a = synth.starshot_frame(fr, spokes=spokes, offsets_mm=[tuple(rng.uniform(-0.6, 0.6, 2)) for _ in range(spokes)],
                             noise_sigma=0.003, seed=seed)
    dpmm = 1 / fr.pixel_size
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        ref = reference_starshot(a, fr.pixel_size, 1000.0)
        o = starshot_oracle.starshot_analyze(a, dpmm)
    for k in ("iterations", "profile_len", "peak_idx", "n_lines", "passed"):
        assert np.array_equal(np.asarray(o[k]), np.asarray(ref[k])), k
    for k in ("wobble_center", "wobble_radius_px", "angles"):
        np.testing.assert_allclose(np.asarray(o[k]), np.asarray(ref[k]), rtol=1e-12, atol=1e-12, err_msg=k)


@pytest.mark.parametrize("seed", [921, 922])
def test_field_oracle_equals_reference_on_fresh_frames(seed):
    from oracle import field_oracle, synth
    from tests.golden.refrun import reference_field

    rng = np.random.default_rng(seed)
    fr = synth.as1200(1000.0)
    a = synth.openfield_frame(fr, field_size_mm=(int(rng.integers(80, 200)), int(rng.integers(80, 200))),
                              cax_offset_mm=tuple(rng.uniform(-8, 8, 2)), seed=seed)
    dpmm = 1 / fr.pixel_size
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        ref = reference_field(a, fr.pixel_size, 1000.0)
        o = field_oracle.field_analyze(a, dpmm)
    for k in ("field_size_horizontal_mm", "field_size_vertical_mm", "beam_center_index_x_y", "left_penumbra_mm", "top_penumbra_mm",
              "flatness_horizontal", "symmetry_vertical", "cax_to_left_mm", "cax_to_top_mm"):
        np.testing.assert_allclose(np.asarray(o[k], dtype=float), np.asarray(ref[k], dtype=float), rtol=0, atol=1e-7, err_msg=k)


@pytest.mark.parametrize("seed", [931, 932])
def test_wl_oracle_equals_reference_on_fresh_frames(seed):
    from oracle import synth, wl_oracle
    from tests.golden.refrun import reference_wl2d

    rng = np.random.default_rng(seed)
    fr = synth.epid1024()
    g, p = float(rng.integers(0, 360)), float(rng.choice([0, 45, 315]))
    a = synth.winstonlutz_frame(fr, offset_mm_left=rng.uniform(-1.5, 1.5), offset_mm_up=rng.uniform(-1.5, 1.5),
                                offset_mm_in=rng.uniform(-1.5, 1.5), gantry=g, couch=p, noise_sigma=0.003, seed=seed)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        ref = reference_wl2d(a, fr.pixel_size, 1000.0, g, 0.0, p)
        o = wl_oracle.wl2d_analyze(a, 1 / fr.pixel_size)
    for k in ("field_cax", "bb", "epid", "cax2bb_vector", "cax2bb_distance", "cax2epid_distance"):
        np.testing.assert_allclose(np.asarray(o[k], dtype=float), np.asarray(ref[k], dtype=float), rtol=0, atol=1e-9, err_msg=k)
