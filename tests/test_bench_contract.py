"""The committed bench records carry every key of the bench contract (CPU-only check of the JSON lines under
profiles/; the numbers themselves come from B200 runs)."""
import json
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _line(name):
    with open(os.path.join(ROOT, "profiles", name)) as f:
        return json.loads([l for l in f if l.startswith("{")][-1])


def test_c3_record_has_contract_keys():
    d = _line("r2_c3_1gpu_bench.json")
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "gpu_launches", "e2e", "clocks", "roofline", "cpu_baseline"):
        assert k in d, k
    assert d["metric"] == "base_x_permutation_fstats_per_sec" and d["n_gpus"] == 1 and d["warmup"] >= 3
    assert d["config"]["n"] == 60 and d["config"]["B"] == 1000 and d["config"]["N"] == 25000000  # BASELINE.json configs[2]
    assert "workload" in d["config"] and "l2" in d["config"]
    e = d["e2e"]
    assert e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0 and 0 < e["value"] < d["value"]
    r = d["roofline"]
    for k in ("bound", "achieved", "peak", "unit", "frac", "traffic"):
        assert k in r, k
    assert abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9 and r["traffic"] and r["kernel_ms_per_step"] < d["ms_per_step"]
    # value = units / time, units = N * (B + 1)
    assert abs(d["value"] - 25000000 * 1001 / (d["ms_per_step"] * 1e-3)) / d["value"] < 1e-6
    c = d["cpu_baseline"]
    assert c["kind"] in ("port", "reference") and c["cores"] >= 1 and "sample" in c
    assert d["clocks"]["sm_mhz"] and not set(d["clocks"]["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
    assert d["gpu_launches"] > 0


def test_reference_arm_record():
    d = _line("r2_c3_reference_arm.json")
    assert d["impl"] == "reference" and d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["value"] == d["value"]
    assert d["cpu_baseline"]["value"] == d["value"]


def test_scaling_records():
    one = _line("r2_c3_1gpu_bench.json")
    for name, n in (("r2_c3_2gpu_weak_bench.json", 2), ("r2_c3_4gpu_weak_bench.json", 4), ("r2_c3_8gpu_weak_bench.json", 8)):
        d = _line(name)
        assert d["n_gpus"] == n and d["scaling"] == "weak" and "multi_gpu" in d
        assert d["value"] > 0.8 * n * one["value"]  # weak scaling stays above 80 %
    g = _line("r2_c4_8gpu_bench.json")
    assert g["scaling"] == "strong" and g["n_gpus"] == 8 and g["config"]["N"] >= 300000000
