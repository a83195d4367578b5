"""GPU: BASELINE.json's full batch sizes through size-independent properties (the oracle needs ~0.3 s per frame, so it is not run
on thousands of frames).  A batch is built by tiling a handful of unique seeded frames in a shuffled order; because frames are
analysed independently, every copy must reproduce its unique frame's result BIT FOR BIT wherever it sits in the batch (chunk
boundaries, work-item boundaries of the persistent kernels, fast / fallback paths), and the unique results themselves are the
ones the golden tests pin to the reference.

configs[1] PicketFence 512 x 1024^2;  configs[2] Winston-Lutz 2048 x 1024^2;  configs[3] Starshot 256 x 1024^2;
configs[4] FieldAnalysis 4096 x 1280^2 (run as 4 calls of 1024 frames: 3.4 GB of host frames per call)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _tile(unique: np.ndarray, n: int, seed: int):
    order = np.random.default_rng(seed).integers(0, len(unique), n)
    order[: len(unique)] = np.arange(len(unique))          # every unique frame occurs at least once
    return unique[order], order


def _assert_rows_equal(rows, ref_rows, order, skip=()):
    for name in rows.dtype.names:
        if name in skip:
            continue
        np.testing.assert_array_equal(rows[name], ref_rows[name][order], err_msg=name)


def test_pf_config2_512_frames():
    from oracle import synth
    from pylinac_b200 import picketfence as pf

    unique = np.stack([synth.bench_pf_frame(i) for i in range(100, 108)])
    ref = pf.analyze_batch(unique, 2.56)
    frames, order = _tile(unique, 512, 1)
    res = pf.analyze_batch(frames, 2.56)
    assert (res.summary["status"] == 0).all() and (res.summary["n_meas"] == 500).all()
    _assert_rows_equal(res.summary, ref.summary, order)
    for name in res.meas.dtype.names:
        np.testing.assert_array_equal(res.meas[name][:, :500], ref.meas[name][order][:, :500], err_msg=name)
    # gold check of the unique frames' headline numbers against the oracle happens in test_gpu_pf.py (same generator)


def test_wl_config3_2048_frames():
    from pylinac_b200 import winston_lutz as wl
    from tests.golden.wl_cases import case_frame

    names = ["g0", "g90", "g180", "g270", "couch45", "big_offset", "noisy", "field30", "fff"]
    unique = np.stack([case_frame(nm)[0] for nm in names])
    dpmm = 1 / case_frame("g0")[1]
    ref = wl.analyze_batch(unique, dpmm)
    gold = np.load("tests/golden/wl_golden.npz")
    for k, nm in enumerate(names):
        assert ref[k].status == 0
        np.testing.assert_allclose([ref[k].bb.x, ref[k].bb.y], gold[f"{nm}/bb"], rtol=0, atol=1e-9)
    frames, order = _tile(unique, 2048, 2)
    res = wl.analyze_batch(frames, dpmm)
    _assert_rows_equal(res.rows, ref.rows, order)


def test_starshot_config4_256_frames():
    from pylinac_b200 import starshot as ss
    from tests.golden.starshot_cases import case_frame

    names = ["perfect6", "offset6", "noisy6", "spokes4", "spokes8", "inverted"]
    unique = np.stack([case_frame(nm)[0] for nm in names])
    dpmm = 1 / case_frame("perfect6")[1]
    ref = ss.analyze_batch(unique, dpmm)
    gold = np.load("tests/golden/starshot_golden.npz")
    for k, nm in enumerate(names):
        assert ref[k].status == 0
        np.testing.assert_allclose([float(ref.rows["wobble_x"][k]), float(ref.rows["wobble_y"][k])], gold[f"{nm}/wobble_center"], rtol=0, atol=1e-6)
    frames, order = _tile(unique, 256, 3)
    res = ss.analyze_batch(frames, dpmm)
    _assert_rows_equal(res.rows, ref.rows, order)


def test_field_config5_4096_frames():
    from pylinac_b200 import field_analysis as fa
    from tests.golden.field_cases import case_frame

    names = ["as1200_150", "as1200_offset", "fwhm_edges", "geometric", "siemens", "elekta", "slope", "no_ground"]
    unique = np.stack([case_frame(nm)[0] for nm in names])       # all 1280 x 1280; analysed here with default arguments
    dpmm = 1 / case_frame("as1200_150")[1]
    ref = fa.analyze_batch(unique, dpmm)
    assert (ref.rows["status"] == 0).all()
    gold = np.load("tests/golden/field_golden.npz")
    np.testing.assert_allclose(ref.rows["beam_center_index_x_y"][0], gold["as1200_150/beam_center_index_x_y"], rtol=0, atol=1e-6)
    for call in range(4):
        frames, order = _tile(unique, 1024, 10 + call)
        res = fa.analyze_batch(frames, dpmm)
        _assert_rows_equal(res.rows, ref.rows, order)
