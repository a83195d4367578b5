"""Small driver for ncu runs of the VMAT pipeline: `n` device-resident DRGS pairs (1280 x 1280), `iters` passes; prints pairs/s."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from pylinac_b200 import _native as nat
from pylinac_b200 import vmat as vm
from bench import _gen_vmat_pair

n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 3
pairs = [_gen_vmat_pair(i) for i in range(4)]
f1 = np.stack([pairs[k % 4][k % 2] for k in range(n)])
f2 = np.stack([pairs[k % 4][1 - k % 2] for k in range(n)])
ctx = nat.Context.default(0)
b1, b2 = nat.Batch.upload(ctx, f1), nat.Batch.upload(ctx, f2)
p = vm._make_params(1 / 0.336, 1.5, (5, 100), [-60, -40, -20, 0, 20, 40, 60], True, True, False)
rows = nat.vmat_analyze(ctx, b1, b2, p)
ctx.sync()
t0 = time.perf_counter()
for _ in range(iters):
    rows = nat.vmat_analyze(ctx, b1, b2, p)
ctx.sync()
dt = (time.perf_counter() - t0) / iters
print(f"vmat {n} pairs: {dt * 1e3:.3f} ms/batch -> {n / dt:.0f} pairs/s; {(f1.nbytes + f2.nbytes) / dt / 1e9:.0f} GB/s on the one-read bytes; ok {int((rows['status'] == 0).sum())}")
