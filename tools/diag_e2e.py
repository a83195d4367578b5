"""GPU-box diagnostics for the end-to-end path: PCIe link, pinned H2D bandwidth, e2e variants."""
import os, subprocess, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from pylinac_b200 import _native as nat
from pylinac_b200 import picketfence as pf
from oracle import synth

print(subprocess.run(["nvidia-smi", "--query-gpu=index,pcie.link.gen.current,pcie.link.gen.max,pcie.link.width.current,pcie.link.width.max", "--format=csv"],
                     capture_output=True, text=True).stdout)
print("cores", len(os.sched_getaffinity(0)))
ctx = nat.Context.default(0)
uniq = np.stack([synth.bench_pf_frame(i) for i in range(8)])
n = 512
pinned = nat.pinned_empty((n, 1024, 1024), np.uint16)
for i in range(n):
    pinned[i] = uniq[i % 8]
pageable = np.array(pinned)
for name, arr in (("pinned", pinned), ("pageable", pageable)):
    for _ in range(2):
        t0 = time.perf_counter()
        b = nat.Batch.upload(ctx, arr)
        dt = time.perf_counter() - t0
        b.free()
    print(f"H2D {name}: {arr.nbytes / dt / 1e9:.1f} GB/s ({dt * 1e3:.1f} ms for {arr.nbytes / 1e6:.0f} MB)")
params = pf.make_params(2.56, (1024, 1024))
for cap in (1024, 512):
    for _ in range(2):
        nat.pf_analyze(ctx, pinned, params, meas_cap=cap)
    t0 = time.perf_counter()
    for _ in range(3):
        s, m = nat.pf_analyze(ctx, pinned, params, meas_cap=cap)
    dt = (time.perf_counter() - t0) / 3
    print(f"e2e pinned meas_cap={cap}: {dt * 1e3:.1f} ms/step, {n / dt:.0f} fps, status ok={int((s['status'] == 0).sum())}")
t0 = time.perf_counter()
s, m = nat.pf_analyze(ctx, pageable, params, meas_cap=512)
dt = time.perf_counter() - t0
print(f"e2e pageable: {dt * 1e3:.1f} ms/step, {n / dt:.0f} fps")
b = nat.Batch.upload(ctx, pinned)
for _ in range(2):
    nat.pf_analyze(ctx, b, params, meas_cap=512)
t0 = time.perf_counter()
s, m = nat.pf_analyze(ctx, b, params, meas_cap=512)
dt = time.perf_counter() - t0
print(f"device-resident incl. D2H of results: {dt * 1e3:.1f} ms/step, {n / dt:.0f} fps")
