"""The CPU-runnable part of bench.py's contract: the reference arm prints one JSON line with the keys the
driver reads, and without a GPU the product arm fails loudly instead of falling back."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(*args, timeout=600):
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True,
                          timeout=timeout, cwd=ROOT)


def test_reference_arm_line(built):
    p = _run("--impl", "reference", "--steps", "2", "--warmup", "1", "--beams", "8", "--azimuths", "256")
    assert p.returncode == 0, p.stderr[-2000:]
    line = json.loads(p.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "scans/s" and line["higher_is_better"] is True
    assert line["value"] > 0 and line["n_gpus"] == 1 and line["gpu_launches"] == 0
    cb = line["cpu_baseline"]
    assert cb["kind"] in ("reference", "port") and cb["cores"] >= 1 and cb["value"] == line["value"]
    e = line["e2e"]
    assert e["value"] == line["value"] and e["h2d_bytes_per_step"] == 0 and e["d2h_bytes_per_step"] == 0
    assert "workload" in line["config"] and "model" not in line["config"]


def test_reference_arm_other_ranks_exit_quietly(built):
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1",
                        "--beams", "8", "--azimuths", "256"], capture_output=True, text=True, timeout=300, cwd=ROOT, env=env)
    assert p.returncode == 0 and p.stdout.strip() == ""


def test_product_arm_needs_a_gpu(built):
    import torch
    if torch.cuda.is_available():
        import pytest
        pytest.skip("GPU present")
    p = _run("--steps", "1", "--warmup", "1", "--beams", "8", "--azimuths", "256", "--no-cpu-baseline")
    assert p.returncode != 0
    assert "CUDA" in (p.stderr + p.stdout)
