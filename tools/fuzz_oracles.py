"""Fuzz the oracle restatements against the UNMODIFIED reference run live (stub-imported from /root/reference): random panels,
frame contents and analyze() arguments; PicketFence bit-identical, Starshot / FieldAnalysis / Winston-Lutz to 1e-7 .. 1e-9, and
the same exception type when the reference raises.  Only runs where /root/reference exists (the build container).

    python tools/fuzz_oracles.py          # round 1: 40 PF + 15 Starshot + 20 Field + 20 WL cases, 0 mismatches
"""
import os
import sys
import time
import warnings

import numpy as np

warnings.simplefilter("ignore")
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import field_oracle, pf_oracle, starshot_oracle, synth, wl_oracle
from tests.golden.refrun import reference_field, reference_pf, reference_starshot, reference_wl2d
from tests.test_oracle_pf import CLOSE, EXACT

bad = 0
t0 = time.time()
for seed in range(2000, 2040):
    rng = np.random.default_rng(seed)
    panel = rng.choice(["epid1024", "as1200", "as1000"])
    fr = {"epid1024": synth.epid1024, "as1200": lambda: synth.as1200(1000.0), "as1000": lambda: synth.as1000(1000.0)}[panel]()
    np_ = int(rng.integers(4, 11)); sp = int(rng.integers(15, 30))
    a = synth.picketfence_frame(fr, pickets=np_, picket_spacing_mm=sp, picket_width_mm=int(rng.integers(2, 6)),
                                picket_offset_error=rng.uniform(-0.8, 0.8, 12), noise_sigma=float(rng.uniform(0.0005, 0.006)), seed=seed,
                                orientation="left_right" if rng.random() < 0.4 else "up_down", blur_mm=float(rng.uniform(0.6, 2.0)))
    if rng.random() < 0.2:
        a = (a.max() - a + a.min()).astype(np.uint16)
    ak = {}
    if rng.random() < 0.3: ak["separate_leaves"] = True; ak["nominal_gap_mm"] = float(rng.integers(2, 6))
    if rng.random() < 0.3: ak["sag_adjustment"] = float(rng.uniform(-2, 2))
    if rng.random() < 0.3: ak["fwxm"] = int(rng.integers(30, 80))
    if rng.random() < 0.3: ak["leaf_analysis_width_ratio"] = float(rng.uniform(0.3, 0.8))
    if rng.random() < 0.3: ak["tolerance"] = float(rng.uniform(0.1, 0.6))
    ck = {}
    if rng.random() < 0.3: ck["crop_mm"] = int(rng.integers(0, 8))
    if rng.random() < 0.2: ck["filter"] = int(rng.choice([3, 5]))
    try:
        ref = reference_pf(a, fr.pixel_size, 1000.0, dict(ck), dict(ak))
        rerr = None
    except Exception as e:
        ref, rerr = None, type(e).__name__
    try:
        o = pf_oracle.pf_analyze(a, 1 / fr.pixel_size, **ck, **ak)
        oerr = None
    except Exception as e:
        o, oerr = None, type(e).__name__
    if rerr or oerr:
        if rerr != oerr:
            bad += 1; print("seed", seed, "exception mismatch", rerr, oerr, panel, ck, ak)
        continue
    for k in EXACT + CLOSE:
        if not np.array_equal(np.asarray(o[k]), np.asarray(ref[k])):
            bad += 1; print("seed", seed, "MISMATCH", k, panel, ck, ak); break
print("picket fence done", bad, round(time.time() - t0))

for seed in range(5000, 5015):
    rng = np.random.default_rng(seed)
    spokes = int(rng.choice([4, 6, 8]))
    fr = synth.epid1024() if rng.random() < 0.7 else synth.as1200(1000.0)
    a = synth.starshot_frame(fr, spokes=spokes, offsets_mm=[tuple(rng.uniform(-0.9, 0.9, 2)) for _ in range(spokes)], noise_sigma=float(rng.uniform(0.001, 0.008)), seed=seed)
    kw = {}
    if rng.random() < 0.3: kw["radius"] = float(rng.uniform(0.4, 0.9))
    if rng.random() < 0.3: kw["fwhm"] = False
    if rng.random() < 0.3: kw["min_peak_height"] = float(rng.uniform(0.1, 0.5))
    if rng.random() < 0.2: kw["recursive"] = False
    res = []
    for fn in (lambda: reference_starshot(a, fr.pixel_size, 1000.0, dict(kw)), lambda: starshot_oracle.starshot_analyze(a, 1 / fr.pixel_size, **kw)):
        try: res.append(fn())
        except Exception as e: res.append(type(e).__name__)
    r, o = res
    if isinstance(r, str) or isinstance(o, str):
        if r != o: bad += 1; print("star", seed, "exc mismatch", r if isinstance(r, str) else "ok", o if isinstance(o, str) else "ok", kw)
        continue
    for k in ("iterations", "profile_len", "peak_idx", "n_lines", "passed"):
        if not np.array_equal(np.asarray(o[k]), np.asarray(r[k])): bad += 1; print("star", seed, "MISMATCH", k, kw); break
    else:
        if not np.allclose(o["wobble_center"], r["wobble_center"], rtol=0, atol=1e-9): bad += 1; print("star", seed, "wobble", kw)
print("starshot done", bad, round(time.time() - t0))
for seed in range(5100, 5120):
    rng = np.random.default_rng(seed)
    fr = synth.as1200(1000.0) if rng.random() < 0.6 else synth.epid1024()
    a = synth.openfield_frame(fr, field_size_mm=(int(rng.integers(50, 220)), int(rng.integers(50, 220))), cax_offset_mm=tuple(rng.uniform(-10, 10, 2)), seed=seed,
                              field="fff" if rng.random() < 0.25 else "filtered", noise_sigma=float(rng.uniform(0.0005, 0.004)))
    kw = {}
    if rng.random() < 0.4: kw["edge_detection_method"] = "FWHM"
    if rng.random() < 0.4: kw["protocol"] = str(rng.choice(["SIEMENS", "ELEKTA", "NONE"]))
    if rng.random() < 0.3: kw["in_field_ratio"] = float(rng.uniform(0.6, 0.85))
    if rng.random() < 0.3: kw["centering"] = str(rng.choice(["Geometric center", "Manual"]))
    if rng.random() < 0.3: kw["vert_width"] = float(rng.uniform(0, 0.05)); kw["horiz_width"] = float(rng.uniform(0, 0.05))
    if rng.random() < 0.3: kw["interpolation_resolution_mm"] = float(rng.choice([0.1, 0.2, 0.5]))
    if rng.random() < 0.2: kw["normalization_method"] = str(rng.choice(["Max", "Geometric center"]))
    if rng.random() < 0.2: kw["penumbra"] = (10, 90)
    try:
        r = reference_field(a, fr.pixel_size, 1000.0, dict(kw)); o = field_oracle.field_analyze(a, 1 / fr.pixel_size, **kw)
    except Exception as e:
        print("field", seed, "exception", type(e).__name__, str(e)[:80], kw); bad += 1; continue
    for k in r:
        if k.startswith("top_") or k not in o: continue
        if not np.allclose(np.asarray(o[k], dtype=float), np.asarray(r[k], dtype=float), rtol=0, atol=1e-7, equal_nan=True):
            bad += 1; print("field", seed, "MISMATCH", k, np.asarray(o[k]), np.asarray(r[k]), kw); break
print("field done", bad, round(time.time() - t0))
for seed in range(5200, 5220):
    rng = np.random.default_rng(seed)
    fr = synth.epid1024() if rng.random() < 0.7 else synth.as1200(1000.0)
    bb = float(rng.choice([3.0, 5.0, 5.0, 8.0]))
    g, c, p = float(rng.integers(0, 360)), float(rng.choice([0, 0, 30, 330])), float(rng.choice([0, 0, 45, 315]))
    a = synth.winstonlutz_frame(fr, bb_size_mm=bb, field_size_mm=(int(rng.integers(16, 45)),) * 2, offset_mm_left=rng.uniform(-3, 3), offset_mm_up=rng.uniform(-3, 3),
                                offset_mm_in=rng.uniform(-3, 3), gantry=g, coll=c, couch=p, noise_sigma=float(rng.uniform(0.001, 0.008)), seed=seed,
                                field="fff" if rng.random() < 0.2 else "perfect")
    if rng.random() < 0.25: a = (a.max() - a + a.min()).astype(np.uint16)
    kw = {"bb_size_mm": bb}
    if rng.random() < 0.2: kw["open_field"] = True
    res = []
    for fn in (lambda: reference_wl2d(a, fr.pixel_size, 1000.0, g, c, p, dict(kw)), lambda: wl_oracle.wl2d_analyze(a, 1 / fr.pixel_size, **kw)):
        try: res.append(fn())
        except Exception as e: res.append(type(e).__name__)
    r, o = res
    if isinstance(r, str) or isinstance(o, str):
        if r != o: bad += 1; print("wl", seed, "exc mismatch", r if isinstance(r, str) else "ok", o if isinstance(o, str) else "ok", kw)
        continue
    for k in ("field_cax", "bb", "epid", "cax2bb_vector", "cax2bb_distance", "cax2epid_distance"):
        if not np.allclose(np.asarray(o[k], dtype=float), np.asarray(r[k], dtype=float), rtol=0, atol=1e-9): bad += 1; print("wl", seed, "MISMATCH", k, kw); break
print("all done", bad, "mismatches", round(time.time() - t0), "s")
sys.exit(1 if bad else 0)
