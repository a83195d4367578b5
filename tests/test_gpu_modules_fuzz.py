"""GPU: Starshot, FieldAnalysis and Winston-Lutz 2-D in CUDA against their oracles on seeded RANDOM frames (the oracles are pinned
to the unmodified reference by the golden tests and by tests/test_oracle_vs_reference_live.py on the same generators)."""
import warnings

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("seed", range(4000, 4008))
def test_starshot_random_case_matches_oracle(seed):
    from oracle import starshot_oracle, synth
    from pylinac_b200 import starshot as ss

    rng = np.random.default_rng(seed)
    spokes = int(rng.choice([4, 6, 8]))
    fr = synth.epid1024() if rng.random() < 0.7 else synth.as1200(1000.0)
    a = synth.starshot_frame(fr, spokes=spokes, offsets_mm=[tuple(rng.uniform(-0.7, 0.7, 2)) for _ in range(spokes)],
                             noise_sigma=float(rng.uniform(0.001, 0.006)), seed=seed)
    kw = {}
    if rng.random() < 0.3:
        kw["radius"] = float(rng.uniform(0.5, 0.9))
    if rng.random() < 0.3:
        kw["fwhm"] = False
    dpmm = 1 / fr.pixel_size
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        try:
            o = starshot_oracle.starshot_analyze(a, dpmm, **kw)
        except RuntimeError:
            o = None
    r = ss.analyze_batch(a[None], dpmm, **kw)[0]
    if o is None:
        assert r.status != 0
        return
    assert r.status == 0
    row = r.r
    assert int(row["iterations"]) == int(o["iterations"]) and int(row["n_lines"]) == int(o["n_lines"])
    npk = int(row["n_peaks"])
    assert np.array_equal(row["peak_idx"][:npk], o["peak_idx"])
    np.testing.assert_allclose([float(row["wobble_x"]), float(row["wobble_y"])], o["wobble_center"], rtol=0, atol=1e-6)
    np.testing.assert_allclose(float(row["wobble_radius_mm"]), float(o["wobble_radius_mm"]), rtol=0, atol=1e-6)
    assert bool(row["passed"]) == bool(o["passed"])


@pytest.mark.parametrize("seed", range(4100, 4108))
def test_field_random_case_matches_oracle(seed):
    from oracle import field_oracle, synth
    from pylinac_b200 import field_analysis as fa

    rng = np.random.default_rng(seed)
    fr = synth.as1200(1000.0) if rng.random() < 0.6 else synth.epid1024()
    a = synth.openfield_frame(fr, field_size_mm=(int(rng.integers(60, 200)), int(rng.integers(60, 200))),
                              cax_offset_mm=tuple(rng.uniform(-8, 8, 2)), seed=seed, field="fff" if rng.random() < 0.25 else "filtered")
    kw = {}
    if rng.random() < 0.4:
        kw["edge_detection_method"] = "FWHM"
    if rng.random() < 0.3:
        kw["protocol"] = str(rng.choice(["SIEMENS", "ELEKTA"]))
    if rng.random() < 0.3:
        kw["in_field_ratio"] = float(rng.uniform(0.6, 0.85))
    dpmm = 1 / fr.pixel_size
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        o = field_oracle.field_analyze(a, dpmm, **kw)
    r = fa.analyze_batch(a[None], dpmm, **kw)
    assert int(r.rows["status"][0]) == 0
    for k in ("field_size_horizontal_mm", "field_size_vertical_mm", "beam_center_index_x_y", "left_penumbra_mm", "right_penumbra_mm",
              "top_penumbra_mm", "bottom_penumbra_mm", "flatness_horizontal", "flatness_vertical", "symmetry_horizontal",
              "symmetry_vertical", "cax_to_left_mm", "cax_to_top_mm", "left_slope_percent_mm"):
        np.testing.assert_allclose(np.asarray(r.rows[k][0], dtype=float), np.asarray(o[k], dtype=float), rtol=0, atol=1e-6, err_msg=k)


@pytest.mark.parametrize("seed", range(4200, 4208))
def test_wl_random_case_matches_oracle(seed):
    from oracle import synth, wl_oracle
    from pylinac_b200 import winston_lutz as wl

    rng = np.random.default_rng(seed)
    fr = synth.epid1024()
    bb = float(rng.choice([5.0, 5.0, 8.0]))
    a = synth.winstonlutz_frame(fr, bb_size_mm=bb, field_size_mm=(int(rng.integers(18, 40)),) * 2,
                                offset_mm_left=rng.uniform(-2, 2), offset_mm_up=rng.uniform(-2, 2), offset_mm_in=rng.uniform(-2, 2),
                                gantry=float(rng.integers(0, 360)), couch=float(rng.choice([0, 0, 45, 315])),
                                noise_sigma=float(rng.uniform(0.001, 0.006)), seed=seed)
    if rng.random() < 0.25:
        a = (a.max() - a + a.min()).astype(np.uint16)
    dpmm = 1 / fr.pixel_size
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        try:
            o = wl_oracle.wl2d_analyze(a, dpmm, bb_size_mm=bb)
        except ValueError:
            o = None
    r = wl.analyze_batch(a[None], dpmm, bb_size_mm=bb)[0]
    if o is None:
        assert r.status != 0
        return
    assert r.status == 0
    assert bool(r.r["inverted"]) == o["inverted"] and int(r.r["threshold_passes"]) == o["threshold_passes"]
    np.testing.assert_allclose([r.bb.x, r.bb.y], o["bb"], rtol=0, atol=1e-9)
    np.testing.assert_allclose([r.field_cax.x, r.field_cax.y], o["field_cax"], rtol=0, atol=1e-9)
    np.testing.assert_allclose(r.cax2bb_distance, o["cax2bb_distance"], rtol=0, atol=1e-9)
