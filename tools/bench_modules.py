"""Device-resident throughput of the Starshot, FieldAnalysis and Winston-Lutz batch pipelines (BASELINE.json configs[2..4]) next to
the CPU oracle port on one host core.  Not the judged benchmark (that is bench.py / PicketFence); prints one JSON line each.

    python tools/bench_modules.py [--star 256] [--field 512] [--iters 3]
"""
import argparse
import json
import os
import sys
import time
import warnings

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

from oracle import field_oracle, starshot_oracle, synth
from pylinac_b200 import _native as nat
from pylinac_b200 import field_analysis as fa
from pylinac_b200 import starshot as ss

ap = argparse.ArgumentParser()
ap.add_argument("--star", type=int, default=256)
ap.add_argument("--field", type=int, default=512)
ap.add_argument("--iters", type=int, default=3)
args = ap.parse_args()
ctx = nat.Context.default(0)
warnings.simplefilter("ignore")


def timed(fn, iters):
    fn()
    ctx.sync()
    t0 = time.perf_counter()
    for _ in range(iters):
        out = fn()
    ctx.sync()
    return (time.perf_counter() - t0) / iters, out


# ---- Starshot: 256 synthetic 1024x1024 star images (docs recipe, per-frame spoke offsets), tiled from 16 unique frames
rng = np.random.default_rng(0)
uniq = np.stack([synth.starshot_frame(synth.epid1024(), offsets_mm=[tuple(rng.uniform(-0.5, 0.5, 2)) for _ in range(6)],
                                      noise_sigma=0.002, seed=100 + i) for i in range(16)])
frames = np.concatenate([uniq] * (args.star // 16))
b = nat.Batch.upload(ctx, frames)
params = ss.make_params(2.56)
dt, rows = timed(lambda: nat.starshot_analyze(ctx, b, params), args.iters)
t0 = time.perf_counter()
o = starshot_oracle.starshot_analyze(uniq[0], 2.56)
cpu = time.perf_counter() - t0
err = max(abs(float(rows["wobble_x"][0]) - o["wobble_center"][0]), abs(float(rows["wobble_y"][0]) - o["wobble_center"][1]))
print(json.dumps({"module": "starshot", "frames": len(frames), "ms_per_batch": dt * 1e3, "frames_per_s": len(frames) / dt,
                  "status_ok": int((rows["status"] == 0).sum()), "iterations_per_frame": float(rows["iterations"].mean()),
                  "cpu_oracle_s_per_frame_1core": cpu, "max_center_err_px_vs_oracle": err, "includes": "D2H of results"}))
b.free()
# ---- FieldAnalysis: 1280x1280 open-field frames (AS1200 at SID 1000), tiled from 16 unique frames
uniq = np.stack([synth.openfield_frame(synth.as1200(1000.0), cax_offset_mm=tuple(rng.uniform(-3, 3, 2)), seed=200 + i) for i in range(16)])
frames = np.concatenate([uniq] * (args.field // 16))
b = nat.Batch.upload(ctx, frames)
dpmm = 1 / 0.336
fp = fa.make_params(dpmm)
dt, rows = timed(lambda: nat.field_analyze(ctx, b, fp), args.iters)
t0 = time.perf_counter()
o = field_oracle.field_analyze(uniq[0], dpmm)
cpu = time.perf_counter() - t0
err = abs(float(rows["field_size_horizontal_mm"][0]) - o["field_size_horizontal_mm"])
print(json.dumps({"module": "field_analysis", "frames": len(frames), "ms_per_batch": dt * 1e3, "frames_per_s": len(frames) / dt,
                  "status_ok": int((rows["status"] == 0).sum()), "gb_per_s_one_read": frames.nbytes / dt / 1e9,
                  "cpu_oracle_s_per_frame_1core": cpu, "field_size_err_mm_vs_oracle": err, "includes": "D2H of results"}))
b.free()
# ---- Winston-Lutz 2-D: 1024x1024 BB + field frames (gantry / couch sweep, seeded BB offsets), tiled from 16 unique frames
from oracle import wl_oracle
from pylinac_b200 import winston_lutz as wlm

uniq = np.stack([synth.winstonlutz_frame(synth.epid1024(), offset_mm_left=rng.uniform(-1, 1), offset_mm_up=rng.uniform(-1, 1),
                                         offset_mm_in=rng.uniform(-1, 1), gantry=22.5 * i, couch=(0, 45, 90, 270, 315)[i % 5] if i % 4 == 0 else 0,
                                         noise_sigma=0.002, seed=300 + i) for i in range(16)])
frames = np.concatenate([uniq] * 32)
b = nat.Batch.upload(ctx, frames)
wp = wlm.make_params(2.56)
dt, rows = timed(lambda: nat.wl2d_analyze(ctx, b, wp), args.iters)
t0 = time.perf_counter()
o = wl_oracle.wl2d_analyze(uniq[0], 2.56)
cpu = time.perf_counter() - t0
err = max(abs(float(rows["bb_x"][0]) - o["bb"][0]), abs(float(rows["bb_y"][0]) - o["bb"][1]))
print(json.dumps({"module": "winston_lutz_2d", "frames": len(frames), "ms_per_batch": dt * 1e3, "frames_per_s": len(frames) / dt,
                  "status_ok": int((rows["status"] == 0).sum()), "threshold_passes_per_frame": float(rows["threshold_passes"].mean()),
                  "gb_per_s_one_read": frames.nbytes / dt / 1e9, "cpu_oracle_s_per_frame_1core": cpu, "max_bb_err_px_vs_oracle": err,
                  "includes": "D2H of results"}))
b.free()
