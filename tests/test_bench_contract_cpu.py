"""CPU: the committed bench lines (profiles/) carry every key of the measurement contract, and bench.py parses / refuses to
run its GPU arm without a device."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _line(name):
    return json.loads(open(os.path.join(ROOT, "profiles", name)).read().strip().splitlines()[-1])


def test_final_bench_line_has_the_contract_keys():
    d = _line("r01_bench_final.json")
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "e2e", "gpu_launches", "clocks", "roofline", "cpu_baseline"):
        assert k in d, k
    assert d["unit"] == "faces/s" and d["higher_is_better"] is True and d["scaling"] == "weak" and d["vs_baseline"] is None
    assert d["data"] == "synthetic" and "workload" in d["config"] and "model" not in d["config"]
    assert d["gpu_launches"] > 0 and d["value"] > 0 and abs(d["value"] - 8e3 / d["ms_per_step"]) < 1e-6 * d["value"]
    e = d["e2e"]
    assert e["unit"] == "faces/s" and e["h2d_bytes_per_step"] == 8 * 256 * 256 * 3 * 4 + 256 * 256 + 8 * 3 * 4
    assert e["d2h_bytes_per_step"] == 8 * 3 * 256 * 256 * 4 and e["value"] != d["value"]
    r = d["roofline"]
    assert r["bound"] in ("hbm", "tensor") and r["unit"] in ("GB/s", "TFLOP/s")
    assert abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9 and r["traffic"] is None or r["traffic"] > 0
    assert abs(r["achieved"] - r["algorithmic_bytes_per_face"] * 8 / (r["ms_per_launch"] * 1e-3) / 1e9) < 1e-6 * r["achieved"]
    c = d["cpu_baseline"]
    assert c["kind"] in ("port", "reference") and c["cores"] >= 1 and c["unit"] == "faces/s" and c["sample"]
    assert set(d["clocks"]) >= {"sm_mhz", "sm_max_mhz", "reasons"}
    assert not set(d["clocks"]["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}


def test_reference_arm_line():
    d = _line("r01_bench_reference_arm.json")
    assert d["impl"] == "reference" and d["metric"] == _line("r01_bench_final.json")["metric"] and d["unit"] == "faces/s"
    assert d["e2e"] == {"value": d["value"], "unit": "faces/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["cpu_baseline"]["value"] == d["value"] and d["cpu_baseline"]["kind"] == "port"


def test_bench_gpu_arm_refuses_to_run_without_a_device():
    import torch
    if torch.cuda.is_available():
        return
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1"], capture_output=True, text=True, timeout=300)
    assert r.returncode != 0 and "no CPU path" in (r.stderr + r.stdout)
