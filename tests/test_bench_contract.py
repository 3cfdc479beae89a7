"""The driver's bench.py contract: one JSON line on stdout with the agreed keys, for the reference arm (CPU, runs
anywhere) and for our arm (GPU)."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(args, timeout):
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + args, capture_output=True, text=True, timeout=timeout,
                       cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, "stdout must carry exactly one JSON line, got %d" % len(lines)
    return json.loads(lines[0])


def test_reference_arm_line():
    d = _run(["--impl", "reference", "--steps", "1", "--warmup", "1"], 600)
    assert d["impl"] == "reference" and d["metric"] == "train_images_per_sec" and d["unit"] == "img/s"
    assert d["higher_is_better"] is True and d["value"] > 0 and d["steps"] == 1
    cb = d["cpu_baseline"]
    # "reference" when oracle/stage_ref.py has staged the unmodified model into baseline/_ref (it has wherever
    # /root/reference exists at build time), "port" otherwise
    staged = os.path.exists(os.path.join(ROOT, "baseline", "_ref", "augmented_cyclegan", "model.py"))
    assert cb["kind"] == ("reference" if staged else "port") and cb["cores"] >= 1 and cb["value"] == d["value"] and "batch" in cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "Augmented CycleGAN 64x64" in d["config"]["workload"] and "sample" in d["config"]


@pytest.mark.gpu
def test_our_arm_line():
    d = _run(["--steps", "3", "--warmup", "3", "--no-baselines"], 600)
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
              "dtype", "data", "config", "e2e", "gpu_launches", "clocks", "roofline"):
        assert k in d, k
    assert d["metric"] == "train_images_per_sec" and d["n_gpus"] == 1 and d["steps"] == 3 and d["warmup"] >= 3
    assert d["dtype"] == "bf16" and d["data"] == "synthetic" and d["scaling"] == "weak" and d["vs_baseline"] is None
    assert d["value"] > 0 and abs(d["value"] - 80 / (d["ms_per_step"] * 1e-3)) < 1e-6 * d["value"]
    assert d["gpu_launches"] >= 700 * 3                       # our kernels, counted by the library
    e = d["e2e"]
    assert 0 < e["value"] <= d["value"] * 1.05 and e["h2d_bytes_per_step"] == 2 * 80 * 3 * 64 * 64 * 4 and e["d2h_bytes_per_step"] > 0
    r = d["roofline"]
    assert r["bound"] == "tensor" and r["unit"] == "TFLOP/s" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
    assert 0 < r["frac"] < 1 and r["peak"] > 0 and (r["traffic"] is None or r["traffic"] > 0)
    assert "workload" in d["config"] and "model" not in d["config"]
    assert set(d["clocks"]) >= {"sm_mhz", "sm_max_mhz", "reasons"}
