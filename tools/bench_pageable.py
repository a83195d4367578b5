"""End-to-end PicketFence from an ordinary (pageable) numpy array vs page-locked frames: the staging ring's host copy is the limiter.
EPID_COPY_THREADS / EPID_COPY_NT select the staging variant (one process per variant: the copy pool is created once)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from pylinac_b200 import _native as nat
from pylinac_b200 import picketfence as pf
from oracle import synth

n = 512
ctx = nat.Context.default(0)
if os.environ.get("BIND"):       # what bench.py does before it allocates: pin the process to the GPU's NUMA node
    from pylinac_b200 import parallel as par
    print("bind:", par.bind_host_to_gpu(0), "affinity", len(os.sched_getaffinity(0)))
uniq = np.stack([synth.bench_pf_frame(i) for i in range(8)])
pinned = nat.pinned_empty((n, 1024, 1024), np.uint16)
for i in range(n):
    pinned[i] = uniq[i % 8]
pageable = np.array(pinned)
params = pf.make_params(2.56, (1024, 1024))
out = {}
for name, arr in (("pinned", pinned), ("pageable", pageable)):
    for _ in range(2):
        nat.pf_analyze(ctx, arr, params)
    t = []
    for _ in range(5):
        t0 = time.perf_counter()
        s, m = nat.pf_analyze(ctx, arr, params)
        t.append(time.perf_counter() - t0)
    out[name] = min(t)
    print(f"{name}: best {min(t) * 1e3:.2f} ms, median {sorted(t)[2] * 1e3:.2f} ms per 512 frames, ok {int((s['status'] == 0).sum())}")
print(f"bind={os.environ.get('BIND', '0')} threads={os.environ.get('EPID_COPY_THREADS', 'default')} nt={os.environ.get('EPID_COPY_NT', '1')}: pageable / pinned = {out['pinned'] / out['pageable']:.3f}")
