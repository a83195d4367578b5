"""CPU: the bench line committed with the round-end profile (profiles/r1m_bench.json, produced by `python bench.py` on a B200)
carries every key of the benchmark contract, and the reference-arm line its own (task statement, section 4)."""
import json

REQUIRED = ["metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
            "dtype", "data", "config", "e2e", "gpu_launches", "roofline", "cpu_baseline", "clocks"]


def test_bench_line_has_the_contract_keys():
    d = json.load(open("profiles/r1m_bench.json"))
    for k in REQUIRED:
        assert k in d, k
    assert d["higher_is_better"] is True and d["scaling"] == "weak" and d["vs_baseline"] is None and d["data"] == "synthetic"
    assert d["n_gpus"] == 1 and d["warmup"] >= 3 and d["gpu_launches"] > 0
    assert set(d["e2e"]) >= {"value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"}
    assert d["e2e"]["h2d_bytes_per_step"] == 512 * 1024 * 1024 * 2 and d["e2e"]["value"] < d["value"]
    r = d["roofline"]
    assert set(r) >= {"bound", "achieved", "peak", "unit", "frac", "traffic"} and r["bound"] == "hbm" and r["unit"] == "GB/s"
    assert abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9 and 0 < r["frac"] < 1
    assert r["traffic"] >= r["algorithmic_bytes_per_launch"]            # DRAM traffic cannot be below the algorithmic bytes
    assert abs(sum(k["share"] for k in r["kernels"]) - 1) < 1e-6
    assert set(d["cpu_baseline"]) >= {"value", "unit", "cores", "kind", "sample"} and d["cpu_baseline"]["kind"] in ("port", "reference")
    assert set(d["clocks"]) >= {"sm_mhz", "sm_max_mhz", "reasons"} and not d["clocks"]["reasons"]
    assert "workload" in d["config"] and "model" not in d["config"]


def test_reference_arm_line():
    d = json.load(open("profiles/r1m_bench_reference.json"))
    assert d["impl"] == "reference" and d["metric"] == json.load(open("profiles/r1m_bench.json"))["metric"]
    assert d["e2e"]["value"] == d["value"] and d["e2e"]["h2d_bytes_per_step"] == 0 and d["cpu_baseline"]["value"] == d["value"]
