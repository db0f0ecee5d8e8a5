"""bench.py's JSON contract: one line, the keys the driver reads, on both arms."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BASE_KEYS = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
             "vs_baseline", "dtype", "data", "config", "e2e", "gpu_launches"}


def run_bench(*args, env=None):
    e = dict(os.environ)
    e.update(env or {})
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True, env=e, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, f"expected exactly one JSON line, got {len(lines)}"
    return json.loads(lines[0])


def test_reference_arm_line():
    """--impl reference: the reference's torch CPU path (oracle port) on the host cores; tiny sample here."""
    d = run_bench("--impl", "reference", "--steps", "1", "--warmup", "1", env={"URED_BENCH_CPU_SAMPLE": "2"})
    assert BASE_KEYS <= set(d) and d["impl"] == "reference"
    assert d["metric"] == "Chamfer+DCD fwd+bwd Gpair/s" and d["unit"] == "Gpair/s" and d["higher_is_better"] is True
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"] > 0
    assert d["e2e"] == {"value": d["value"], "unit": "Gpair/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"] and "model" not in d["config"] and d["vs_baseline"] is None


@pytest.mark.gpu
def test_gpu_arm_line():
    d = run_bench("--steps", "3", "--warmup", "3", "--no-cpu-baseline")
    assert BASE_KEYS | {"roofline", "clocks"} <= set(d) and "impl" not in d
    assert d["n_gpus"] == 1 and d["steps"] == 3 and d["scaling"] == "weak" and d["dtype"] == "f32" and d["value"] > 0
    assert d["gpu_launches"] == 4 * 3                       # pack, nn_tc_kernel, dcd_fwd_kernel, grad_gather_kernel per step
    r = d["roofline"]
    assert r["unit"] == "TFLOP/s" and 0.3 < r["frac"] < 1.2 and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
    e = d["e2e"]
    assert e["value"] > 0 and e["h2d_bytes_per_step"] == (640 + 64) * 2048 * 3 * 4 and e["d2h_bytes_per_step"] == 640 * 4
    assert "workload" in d["config"] and set(d["clocks"]) >= {"sm_mhz", "sm_max_mhz", "reasons"}
