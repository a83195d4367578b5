"""Small driver for ncu runs: a 512-frame device-resident batch tiled from 8 unique synthetic frames, `iters` pipeline passes.
    ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/x.csv python tools/prof_pf.py 2
"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from pylinac_b200 import _native as nat
from pylinac_b200 import picketfence as pf
from oracle import synth

iters = int(sys.argv[1]) if len(sys.argv) > 1 else 2
n = int(sys.argv[2]) if len(sys.argv) > 2 else 512
uniq = np.stack([synth.bench_pf_frame(i) for i in range(8)])
frames = np.concatenate([uniq] * (n // 8))
ctx = nat.Context.default(0)
b = nat.Batch.upload(ctx, frames)
params = pf.make_params(2.56, (1024, 1024))
nat.pf_bench(ctx, b, params, 1)   # warm-up: module load, attribute calls, scratch allocation
total, stats, launches = nat.pf_bench(ctx, b, params, iters)
print(f"{n} frames x {iters} iters: {total / iters:.3f} ms/iter -> {n * iters / total * 1e3:.0f} fps; stats kernel {stats / iters:.3f} ms; launches {launches}")
s, m = nat.pf_analyze(ctx, b, params)
print("status ok:", int((s["status"] == 0).sum()), "n_meas", s["n_meas"][:4], "max_err", s["max_error_mm"][:4])
