"""bench.py contract on the CPU: the reference arm (`--impl reference`) times the oracle on the host cores and prints
one JSON line with the keys the driver reads; the GPU arm must refuse to run without a CUDA device instead of
falling back to anything."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(args, timeout=600):
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + args, cwd=ROOT, capture_output=True, text=True,
                          timeout=timeout)


def test_reference_arm_json_line():
    res = _run(["--impl", "reference", "--steps", "3", "--warmup", "1"])
    assert res.returncode == 0, res.stderr[-2000:]
    line = json.loads(res.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference"
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
                "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert key in line, key
    assert line["metric"] == "pdhg_iterations_per_sec_maxcut_n2000" and line["unit"] == "iterations/s"
    assert line["dtype"] == "f64" and line["data"] == "synthetic" and line["vs_baseline"] is None and line["higher_is_better"] is True
    assert line["steps"] == 3 and line["value"] > 0
    assert "workload" in line["config"] and "model" not in line["config"]
    cb = line["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == line["value"] and cb["sample"]
    e2e = line["e2e"]
    assert e2e["value"] == line["value"] and e2e["unit"] == line["unit"]
    assert e2e["h2d_bytes_per_step"] == 0 and e2e["d2h_bytes_per_step"] == 0


def test_gpu_arm_needs_a_gpu():
    import torch
    if torch.cuda.is_available():
        return
    res = _run(["--steps", "1", "--warmup", "1", "--no-cpu-baseline"], timeout=300)
    assert res.returncode != 0
    assert "reference" not in res.stdout            # no silent switch to the CPU arm
