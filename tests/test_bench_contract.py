"""CPU: the bench lines committed with the round-end profiles (profiles/r1m_bench.json, profiles/r2o_bench.json, produced by `python bench.py` on a B200)
carries every key of the benchmark contract, and the reference-arm line its own (task statement, section 4)."""
import json

REQUIRED = ["metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
            "dtype", "data", "config", "e2e", "gpu_launches", "roofline", "cpu_baseline", "clocks"]


import pytest


@pytest.mark.parametrize("path", ["profiles/r1m_bench.json", "profiles/r2o_bench.json"])
def test_bench_line_has_the_contract_keys(path):
    d = json.load(open(path))
    for k in REQUIRED:
        assert k in d, k
    assert d["higher_is_better"] is True and d["scaling"] == "weak" and d["vs_baseline"] is None and d["data"] == "synthetic"
    assert d["n_gpus"] == 1 and d["warmup"] >= 3 and d["gpu_launches"] > 0
    assert set(d["e2e"]) >= {"value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"}
    assert d["e2e"]["h2d_bytes_per_step"] == 512 * 1024 * 1024 * 2 and d["e2e"]["value"] < d["value"]
    r = d["roofline"]
    assert set(r) >= {"bound", "achieved", "peak", "unit", "frac", "traffic"} and r["bound"] == "hbm" and r["unit"] == "GB/s"
    assert abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9 and 0 < r["frac"] < 1
    if "k_pf_stream" in r["kernel"]:
        assert r["traffic"] >= r["algorithmic_bytes_per_launch"]        # the kernel that reads every frame: DRAM traffic >= the frames
    else:
        assert r["traffic"] > 0                                           # e.g. k_pf_win_medians re-reads only the window bands (~36 % of a frame)
    # per-kernel event marks inside the timed region: the shares add up to the step (the interval before an iteration's first mark,
    # i.e. its 1 KB constant upload, is not attributed to a kernel)
    assert abs(sum(k["share"] for k in r["kernels"]) - 1) < 5e-3
    assert set(d["cpu_baseline"]) >= {"value", "unit", "cores", "kind", "sample"} and d["cpu_baseline"]["kind"] in ("port", "reference")
    assert set(d["clocks"]) >= {"sm_mhz", "sm_max_mhz", "reasons"} and not d["clocks"]["reasons"]
    assert "workload" in d["config"] and "model" not in d["config"]


@pytest.mark.parametrize("path", ["profiles/r1m_bench_reference.json", "profiles/r2n_bench_reference.json"])
def test_reference_arm_line(path):
    d = json.load(open(path))
    assert d["impl"] == "reference" and d["metric"] == json.load(open("profiles/r1m_bench.json"))["metric"]
    assert d["e2e"]["value"] == d["value"] and d["e2e"]["h2d_bytes_per_step"] == 0 and d["cpu_baseline"]["value"] == d["value"]


def test_round2_bench_line_reports_the_fallback_and_module_workloads():
    d = json.load(open("profiles/r2o_bench.json"))
    m = d["config"]["mixed_noisy_5pct"]
    assert m["ratio_to_clean_step"] < 1.3 and m["frames_rerun_per_step"] == 25 and m["frames_exact_pipeline_per_step"] == 0
    assert d["roofline"]["kernel"] == max(d["roofline"]["kernels"], key=lambda k: k["share"])["kernel"]      # the dominant kernel by share
    assert set(d["modules"]) >= {"winston_lutz_2d", "starshot", "field_analysis", "vmat_drgs"}
    assert all(v["status_ok"] == v.get("frames", v.get("pairs")) for v in d["modules"].values())
    assert 0.5 < d["e2e_pageable"]["frac_of_pinned"] < 1.0
