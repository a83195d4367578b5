"""One analyze pass of each non-PF module on a small batch, for `ncu --metrics gpu__time_duration.sum` launch lists.
usage: python tools/prof_modules.py {wl|star|field} [frames]"""
import sys

import numpy as np

sys.path.insert(0, ".")
import bench  # noqa: E402  (frame generators)
from pylinac_b200 import _native as nat  # noqa: E402
from pylinac_b200 import field_analysis as fa, starshot as ss, winston_lutz as wlm  # noqa: E402

kind = sys.argv[1]
count = int(sys.argv[2]) if len(sys.argv) > 2 else 64
ctx = nat.Context.default()
uniq = np.stack([bench._gen_module_frames((kind, i)) for i in range(8)])
frames = np.concatenate([uniq] * (count // 8))
spec = {"wl": (lambda: wlm.make_params(2.56), nat.wl2d_analyze), "star": (lambda: ss.make_params(2.56), nat.starshot_analyze),
        "field": (lambda: fa.make_params(1 / 0.336), nat.field_analyze)}[kind]
b = nat.Batch.upload(ctx, frames)
rows = spec[1](ctx, b, spec[0]())          # warm-up (scratch allocation)
ctx.sync()
import time
t0 = time.perf_counter()
rows = spec[1](ctx, b, spec[0]())
ctx.sync()
print(kind, count, "frames", (time.perf_counter() - t0) * 1e3, "ms", "ok", int((rows["status"] == 0).sum()))
if kind == "star":
    print("iterations per frame (StarProfile constructions):", rows["iterations"][:8].tolist(), "mean", float(rows["iterations"].mean()))
