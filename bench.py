#!/usr/bin/env python
"""Benchmark of the hot path named by BASELINE.json: EPID frames/s (1024x1024) through PicketFence.analyze().

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N ...            # the reference algorithm on the host CPU cores (oracle port)

A "step" is one pass of the whole PicketFence pipeline over one batch of synthetic frames (config.workload).
  value  : whole-job frames/s with the batch already resident in HBM (CUDA events, max over ranks)
  e2e    : the same metric through the public API `pylinac_b200.picketfence.analyze_batch` with HOST (pinned)
           frames -- chunked H2D copies and the D2H of the results are inside the timed region
  roofline: the frame-streaming kernel (k_pf_stream: ONE read of every frame through a TMA ring) vs the measured HBM
           copy bandwidth; its time comes from CUDA events around that kernel inside the timed region
  cpu_baseline: the oracle port (numpy/scipy restatement of the reference, bit-identical to it on the golden
           cases) on all host cores for a bounded sample of the same frames
"""
from __future__ import annotations

import argparse
import json
import multiprocessing as mp
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

FRAME_SHAPE = (1024, 1024)
DPMM = 2.56
PER_GPU_FRAMES = 512          # BASELINE.json configs[1]
METRIC = "EPID frames/sec (1024x1024) through PicketFence.analyze()"


def _gen_frame(i):
    from oracle import synth

    return synth.bench_pf_frame(i, FRAME_SHAPE)


def _oracle_one(i):
    import warnings

    from oracle import pf_oracle, synth

    a = synth.bench_pf_frame(i, FRAME_SHAPE)
    t0 = time.perf_counter()
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        r = pf_oracle.pf_analyze(a, DPMM)
    return time.perf_counter() - t0, r["n_meas"]


def host_cores() -> int:
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def cpu_reference_fps(n_frames: int, cores: int, start: int = 0):
    """Frames/s of the CPU restatement of the reference on `cores` processes (frame generation excluded)."""
    for k in ("OMP_NUM_THREADS", "MKL_NUM_THREADS", "OPENBLAS_NUM_THREADS"):
        os.environ[k] = "1"
    ctx = mp.get_context("fork")
    with ctx.Pool(cores) as pool:
        pool.map(_oracle_one, range(start, start + cores))           # warm the workers (imports)
        t0 = time.perf_counter()
        res = pool.map(_oracle_one, range(start, start + n_frames), chunksize=1)
        wall = time.perf_counter() - t0
    busy = sum(r[0] for r in res)
    # wall includes the generation of each frame inside the worker; the per-frame analysis time is measured
    # inside the worker, so throughput = frames / (sum of analysis time / cores)
    return n_frames / (busy / cores), wall, busy / n_frames


def generate_frames(n: int, start: int, cores: int) -> np.ndarray:
    ctx = mp.get_context("fork")
    with ctx.Pool(cores) as pool:
        frames = pool.map(_gen_frame, range(start, start + n), chunksize=4)
    return np.stack(frames)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""

    Q = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index: int):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            p = [x.strip() for x in ln.split(",")]
            if len(p) < 8:
                continue
            try:
                sm.append(float(p[1]))
                mx.append(float(p[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), p[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def dist_setup():
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        import torch.distributed as dist

        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group(backend="gloo", rank=rank, world_size=world)   # plumbing only: barrier + max(time)
        return dist, world, rank, local
    return None, 1, 0, 0


def run_reference(args):
    # under torchrun only rank 0 measures (the host cores are shared by all ranks); no process group is needed for that
    if int(os.environ.get("RANK", "0")) != 0:
        return 0
    cores = host_cores()
    sample = max(cores, min(4 * cores, 64))
    times = []
    for _ in range(args.warmup):
        cpu_reference_fps(cores, cores)
    fps_all = []
    for s in range(args.steps):
        fps, wall, per = cpu_reference_fps(sample, cores, start=1000 + s * sample)
        fps_all.append(fps)
        times.append(sample / fps * 1e3)
    fps = statistics.mean(fps_all)
    out = {
        "impl": "reference", "metric": METRIC, "value": fps, "unit": "frames/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": statistics.mean(times), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"PicketFence.analyze() on synthetic 1024x1024 MLC picket frames; each step = {sample} frames "
                               f"on {cores} host processes (bounded sample of the {PER_GPU_FRAMES}-frame batch)"},
        "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": cores, "kind": "port",
                         "sample": f"{sample} frames per step, analysis time only (frame generation excluded)"},
        "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(out))
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--frames", type=int, default=PER_GPU_FRAMES, help="frames per GPU per step")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    dist, world, rank, local = dist_setup()
    cores = host_cores()
    n = args.frames
    # ---- everything that forks happens BEFORE the CUDA context exists
    cpu_base = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        sample = max(cores, min(2 * cores, 32))
        fps, wall, per = cpu_reference_fps(sample, cores, start=2000)
        cpu_base = {"value": fps, "unit": "frames/s", "cores": cores, "kind": "port",
                    "sample": f"{sample} of the batch's frames on {cores} processes; {per * 1e3:.0f} ms/frame/core; analysis only"}
    gen_cores = max(1, cores // world)
    frames_np = generate_frames(n, start=rank * n, cores=gen_cores)

    from pylinac_b200 import _native as nat
    from pylinac_b200 import picketfence as pf

    ndev = nat.device_count()
    if ndev == 0:
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback)")
    ctx = nat.Context.default(local % ndev)
    from pylinac_b200 import parallel as par

    numa = par.bind_host_to_gpu(local % ndev)      # before the page-locked buffers exist: first touch puts them on the GPU's node
    params = pf.make_params(DPMM, FRAME_SHAPE)
    pinned = nat.pinned_empty(frames_np.shape, np.uint16)
    pinned[...] = frames_np
    del frames_np
    batch = nat.Batch.upload(ctx, pinned)
    H = FRAME_SHAPE[0] - 2 * params.crop_px
    W = FRAME_SHAPE[1] - 2 * params.crop_px

    if world > 1:
        import torch

        uid = torch.zeros(128, dtype=torch.uint8)
        if rank == 0:
            buf = np.zeros(128, np.uint8)
            nat.check(nat.lib().epid_comm_unique_id(buf.ctypes.data))
            uid = torch.from_numpy(buf)
        dist.broadcast(uid, src=0)
        nat.check(nat.lib().epid_comm_init(ctx.handle, world, rank, uid.numpy().ctypes.data))

    def barrier():
        if dist is not None:
            dist.barrier()

    # ---- warm-up (also grows the scratch arenas)
    nat.pf_bench(ctx, batch, params, max(args.warmup, 3))
    for _ in range(max(args.warmup, 3)):
        res = pf.analyze_batch(pinned, DPMM, meas_cap=1024)
        if world > 1:   # the first collective of a communicator sets up its channels: keep that out of the timed region
            allsum = np.empty(world * n, nat.PF_SUMMARY_DTYPE)
            nat.check(nat.lib().epid_gather_results(ctx.handle, res.summary.ctypes.data, res.summary.nbytes, allsum.ctypes.data))

    # ---- timed: device-resident
    clocks = ClockSampler(local % ndev)
    clocks.start()
    barrier()
    ctx.sync()
    l0 = ctx.launches()
    total_ms, stats_ms, launches = nat.pf_bench(ctx, batch, params, args.steps)
    ctx.sync()
    barrier()
    launches_timed = ctx.launches() - l0
    # ---- timed: end to end through the public API, host (pinned) frames in, host results out
    summ_bytes = nat.PF_SUMMARY_DTYPE.itemsize * n
    meas_cap = 1024
    meas_bytes = nat.PF_MEAS_DTYPE.itemsize * n * meas_cap
    barrier()
    ctx.sync()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        res = pf.analyze_batch(pinned, DPMM, meas_cap=meas_cap)
        if world > 1:
            allsum = np.empty(world * n, nat.PF_SUMMARY_DTYPE)
            nat.check(nat.lib().epid_gather_results(ctx.handle, res.summary.ctypes.data, res.summary.nbytes, allsum.ctypes.data))
    ctx.sync()
    e2e_s = time.perf_counter() - t0
    barrier()
    clk = clocks.stop()
    # ---- untimed: per-kernel device times (CUDA events between the kernels) for the roofline table
    stage_ms = nat.pf_bench_stages(ctx, batch, params, max(2, min(args.steps, 5))) if rank == 0 else {}
    assert all(int(s) == 0 for s in res.summary["status"]), "pipeline reported a failed frame"

    # ---- max over ranks
    if dist is not None:
        import torch

        t = torch.tensor([total_ms, e2e_s * 1e3, stats_ms], dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms, e2e_ms, stats_ms = (float(x) for x in t)
    else:
        e2e_ms = e2e_s * 1e3
    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        peak_src = "MEASURED_PEAKS.json hbm_gbs" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
        alg_bytes = n * H * W * 2                      # one read of every analysed frame view per launch
        ach = alg_bytes / (stats_ms / args.steps * 1e-3) / 1e9
        traffic = None
        try:
            tj = json.load(open(os.path.join(ROOT, "profiles", "stream_traffic.json")))
            # ncu dram__bytes_read + dram__bytes_write of one k_pf_stream launch, scaled from the captured batch size
            traffic = tj["dram_bytes_per_launch"] * n / tj["frames"]
        except Exception:
            pass
        frames_total = world * n * args.steps
        # per-kernel table: algorithmic bytes = the pixels the kernel has to read once (pilot: every 32nd row; windows: the
        # (leaf, picket) windows of PicketFence._get_mlc_window; tail / finalize: 1-D partial sums and window results)
        m0 = int(res.summary["n_meas"][0])
        widths = {int(params.leaf_num[i]): params.leaf_width_mm[i] * DPMM for i in range(params.n_leaves)}
        spacing = int(float(res.summary["picket_spacing_px"][0]))
        win_px = sum(int(widths[int(l)]) * spacing for l in res.meas["leaf_num"][0, :m0])
        alg = {"k_pf_init + k_pf_pilot": n * ((H + 31) // 32) * W * 2, "k_pf_stream": alg_bytes, "k_pf_windows_fast": n * win_px * 2}
        ktable = []
        step_ms = sum(stage_ms.values()) or 1.0
        for name, ms in stage_ms.items():
            if ms <= 0:
                continue
            ab = alg.get(name)
            ktable.append({"kernel": name, "ms": ms, "share": ms / step_ms, "algorithmic_bytes": ab,
                           "GBps": (ab / (ms * 1e-3) / 1e9) if ab else None, "frac": (ab / (ms * 1e-3) / 1e9 / peak) if ab else None})
        pipe_gbs = (n * args.steps / (total_ms * 1e-3)) * H * W * 2 / 1e9
        out = {
            "metric": METRIC, "value": frames_total / (total_ms * 1e-3), "unit": "frames/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u16 pixels / int exact sums / f64 profiles", "data": "synthetic",
            "config": {"workload": f"PicketFence.analyze() on a batch of {n} synthetic 1024x1024 MLC picket frames per GPU "
                                   "(BASELINE.json configs[1]); 10 pickets x 50 leaf pairs = 500 kisses per frame",
                       "frames_per_gpu": n, "parallelism": f"frames sharded over {world} GPU(s), no data-path collective",
                       "host_numa": numa,
                       "l2": f"batch = {n * FRAME_SHAPE[0] * FRAME_SHAPE[1] * 2 / 1e6:.0f} MB per GPU, larger than the 126 MB L2; no flush needed"},
            "e2e": {"value": frames_total / (e2e_ms * 1e-3), "unit": "frames/s", "h2d_bytes_per_step": int(n * FRAME_SHAPE[0] * FRAME_SHAPE[1] * 2),
                    "d2h_bytes_per_step": int(summ_bytes + meas_bytes), "ms_per_step": e2e_ms / args.steps,
                    "api": "pylinac_b200.picketfence.analyze_batch(host uint16 frames) -> per-frame results"
                           + (" + ncclAllGather of the summaries" if world > 1 else "")},
            "gpu_launches": int(launches_timed),
            "roofline": {"bound": "hbm", "kernel": "k_pf_stream<4>", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                         "traffic": traffic, "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": alg_bytes, "kernel_ms_per_launch": stats_ms / args.steps,
                         "kernel_share_of_step": stats_ms / total_ms,
                         "note": "k_pf_stream is the only kernel that reads whole frames from HBM; the (leaf, picket) window kernel is "
                                 "the longest kernel of the step but is bound by integer issue (sorting-network medians), not by HBM",
                         "pipeline": {"achieved": pipe_gbs, "frac": pipe_gbs / peak,
                                      "what": "device-resident frames/s per GPU x H*W*2 bytes (SURVEY.md 8(d) one-read metric) / peak"},
                         "kernels": ktable},
            "clocks": clk,
        }
        if cpu_base is not None:
            out["cpu_baseline"] = cpu_base
        print(json.dumps(out))
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
