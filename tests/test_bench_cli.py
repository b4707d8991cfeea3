"""The bench.py contract that can be checked without a GPU: the reference arm's JSON line, and that the product
arm fails loudly (no CPU fallback) when there is no CUDA device."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(*args, timeout=600):
    env = dict(os.environ)
    env.pop("RANK", None)
    env.pop("WORLD_SIZE", None)
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True, timeout=timeout,
                          cwd=ROOT, env=env)


def test_reference_arm_line():
    p = _run("--impl", "reference", "--steps", "1", "--warmup", "0")
    assert p.returncode == 0, p.stderr[-2000:]
    line = json.loads(p.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference"
    assert line["metric"].startswith("Mrays/s") and line["unit"] == "Mrays/s"
    assert line["n_gpus"] == 1 and line["steps"] == 1 and line["higher_is_better"] is True
    assert line["vs_baseline"] is None and "AncientTemple.vox" in line["data"]
    assert "workload" in line["config"] and "configs[2]" in line["config"]["workload"]
    sys.path.insert(0, ROOT)
    import bench
    assert line["config"] == bench.CONFIG  # the two arms carry the same config dict (the driver compares them)
    assert line["value"] > 0 and line["ms_per_step"] > 0
    cb = line["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == line["value"] and cb["sample"]
    e2e = line["e2e"]
    assert e2e["value"] == line["value"] and e2e["unit"] == line["unit"]
    assert e2e["h2d_bytes_per_step"] == 0 and e2e["d2h_bytes_per_step"] == 0


def test_product_arm_needs_a_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    p = _run("--steps", "1", "--warmup", "1", "--no-cpu", timeout=300)
    assert p.returncode != 0 and not p.stdout.strip(), "bench.py must not produce a line without a GPU"
