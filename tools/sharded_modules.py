"""Multi-GPU run of the per-image modules (BASELINE.json configs[2] and configs[4]): frames shard by index over
the ranks, every rank analyses its shard on its own GPU, one NCCL all-gather of the fixed-size result rows.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tools/sharded_modules.py

Prints per module: frames, world size, wall time of (H2D + analysis + gather) as max over ranks, and whether every gathered row
equals the row of its unique source frame computed on one GPU (bit for bit)."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist

from pylinac_b200 import _native as nat
from pylinac_b200 import field_analysis as fa
from pylinac_b200 import parallel as par
from pylinac_b200 import winston_lutz as wl
from tests.golden import field_cases, wl_cases


def main():
    dist.init_process_group(backend="gloo")
    world, rank, local = par.world_info()
    ctx = nat.Context.default(local % nat.device_count())
    par.bind_host_to_gpu(local % nat.device_count())
    par.init_comm(ctx, dist)

    def run(name, unique, n_total, analyze_rows, seed):
        order = np.random.default_rng(seed).integers(0, len(unique), n_total)
        ref = analyze_rows(unique)
        lo, hi = par.shard_range(n_total, world, rank)

        class Tiled:                       # only this rank's shard is ever materialised
            def __len__(self):
                return n_total

            def __getitem__(self, sl):
                return unique[order[sl]]

        par.analyze_sharded(analyze_rows, Tiled(), ctx=ctx)      # warm-up (scratch arenas, NCCL channels)
        dist.barrier()
        ctx.sync()
        t0 = time.perf_counter()
        rows = par.analyze_sharded(analyze_rows, Tiled(), ctx=ctx)
        ctx.sync()
        t = torch.tensor([time.perf_counter() - t0], dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        same = all(np.array_equal(rows[k], ref[k][order], equal_nan=True) if rows[k].dtype.kind == "f" else np.array_equal(rows[k], ref[k][order])
                   for k in rows.dtype.names)
        if rank == 0:
            print(f"{name}: {n_total} frames on {world} GPU(s), shard {hi - lo}; {float(t[0]) * 1e3:.1f} ms incl. tiling + H2D "
                  f"-> {n_total / float(t[0]):.0f} frames/s; gathered rows identical to the single-GPU rows: {same}", flush=True)
        return same

    ok = True
    names = ["g0", "g90", "g180", "g270", "couch45", "big_offset", "noisy", "field30", "fff"]
    uw = np.stack([wl_cases.case_frame(nm)[0] for nm in names])
    dpmm_wl = 1 / wl_cases.case_frame("g0")[1]
    ok &= run("Winston-Lutz 2-D (configs[2])", uw, 2048, lambda f: wl.analyze_batch(f, dpmm_wl).rows, 5)
    fnames = ["as1200_150", "as1200_offset", "fwhm_edges", "geometric"]
    uf = np.stack([field_cases.case_frame(nm)[0] for nm in fnames])
    dpmm_f = 1 / field_cases.case_frame("as1200_150")[1]
    ok &= run("FieldAnalysis (configs[4])", uf, 4096, lambda f: fa.analyze_batch(f, dpmm_f).rows, 6)
    dist.barrier()
    dist.destroy_process_group()
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
