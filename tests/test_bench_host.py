"""CPU: host-side pieces of bench.py that the driver's runs depend on -- the spawned-process isolation of the forking helpers (frame
generation, CPU baseline) and the reference arm on a tiny sample."""
import json
import subprocess
import sys

import numpy as np


def test_isolated_helpers_return_what_the_functions_return():
    import bench

    frames = bench._isolated(bench.generate_frames, 3, 5, 2, chunked=True)          # spawned interpreter -> forked pool -> chunks over a pipe
    assert frames.shape == (3, 1024, 1024) and frames.dtype == np.uint16
    np.testing.assert_array_equal(frames[1], bench._gen_frame(6))
    assert bench._isolated(max, 3, 7) == 7
    try:
        bench._isolated(int, "not a number")
    except RuntimeError as e:
        assert "ValueError" in str(e)
    else:
        raise AssertionError("a failure inside the isolated process must surface in the parent")


def test_reference_arm_prints_one_contract_line():
    out = subprocess.run([sys.executable, "bench.py", "--impl", "reference", "--steps", "1", "--warmup", "0", "--frames", "2"],
                         capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    d = json.loads(out.stdout.strip().splitlines()[-1])
    assert d["impl"] == "reference" and d["unit"] == "frames/s" and d["value"] > 0
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["cpu_baseline"]["kind"] == "port"
