"""EPID_OPT_PF_SPLIT experiment: device-resident 512-frame batch, S sub-batches on S streams (S = 1, 2, 3, 4): ms per step and a
bit-identity check of the result rows against the single-stream run."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from pylinac_b200 import _native as nat
from pylinac_b200 import picketfence as pf
from oracle import synth

OPT_SPLIT = 4
n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 20
uniq = np.stack([synth.bench_pf_frame(i) for i in range(16)])
frames = np.concatenate([uniq] * (n // 16))
ctx = nat.Context.default(0)
b = nat.Batch.upload(ctx, frames)
params = pf.make_params(2.56, (1024, 1024))
ref = None
for S in (1, 2, 3, 4, 1):
    ctx.set_option(OPT_SPLIT, S)
    nat.pf_bench(ctx, b, params, 3)
    best = 1e9
    for rep in range(3):
        total, _, launches = nat.pf_bench(ctx, b, params, iters)
        best = min(best, total / iters)
    s, m = nat.pf_analyze(ctx, b, params)
    if ref is None:
        ref = (s.copy(), m.copy())
    same = s.tobytes() == ref[0].tobytes() and m.tobytes() == ref[1].tobytes()
    print(f"split {S}: {best:.3f} ms/step -> {n / best * 1e3:.0f} frames/s; launches/step {launches / iters:.1f}; results identical to split 1: {same}; ok {int((s['status'] == 0).sum())}")
