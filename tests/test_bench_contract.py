"""bench.py's contract where it can be checked without a GPU: the reference arm (the CPU oracle timed through bench.py) prints one JSON
line with the keys the driver reads, and the product arm refuses to run without a CUDA device instead of falling back to the CPU."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(*args):
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True, cwd=ROOT, timeout=600)


def test_reference_arm_json_line(orc_mod):
    r = _run("--impl", "reference", "--config", "c1", "--steps", "2", "--warmup", "1")
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    b = json.loads(lines[0])
    assert b["impl"] == "reference" and b["metric"] == "Mrays/s" and b["unit"] == "Mrays/s" and b["higher_is_better"] is True
    assert b["n_gpus"] == 1 and b["steps"] == 2 and b["vs_baseline"] is None and b["data"] == "synthetic"
    assert b["warmup"] == 3  # W >= 3 is enforced (timing rules): the line reports the warm-up steps actually done
    assert b["value"] > 0 and b["ms_per_step"] > 0
    assert b["config"]["workload"].startswith("c1:")
    cb = b["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == b["value"] and cb["sample"]
    assert b["e2e"] == {"value": b["value"], "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    # rays per step x steps / time: the value is consistent with ms_per_step
    assert "gpu_launches" not in b or b["gpu_launches"] == 0


def test_product_arm_needs_a_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present: the product arm runs (bench.py itself is exercised by the driver)")
    r = _run("--config", "c1", "--steps", "1", "--warmup", "0")
    assert r.returncode != 0
    assert r.stdout.strip() == ""  # no JSON line, no number
    assert "no CPU fallback" in r.stderr
