"""bench.py's contract, the parts that can be checked without a GPU: the reference arm (`--impl reference`: the imported
unmodified reference on the host cores, oracle/_ref) prints exactly one JSON line with the driver's keys, and this
repo's own arm refuses to produce a number without a CUDA device (no CPU fallback)."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(*flags, timeout=600):
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *flags], capture_output=True, text=True,
                          timeout=timeout, cwd=ROOT)


def test_reference_arm_prints_one_contract_line():
    if not os.path.isdir(os.path.join(ROOT, "oracle", "_ref", "ngu")) and not os.path.isdir("/root/reference/ngu"):
        pytest.skip("neither the staged reference (oracle/_ref) nor /root/reference is present")
    res = _run("--impl", "reference", "--steps", "2", "--warmup", "3", "--actors-per-gpu", "8", "--capacity", "500")
    assert res.returncode == 0, res.stderr[-2000:]
    lines = [l for l in res.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, lines                       # ONE JSON line on stdout, everything else on stderr
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "episodic intrinsic-reward queries/sec" and d["unit"] == "queries/s"
    assert d["steps"] == 2 and d["warmup"] == 3 and d["n_gpus"] == 1
    assert d["higher_is_better"] is True and d["vs_baseline"] is None and d["dtype"] == "f32" and d["data"] == "synthetic"
    assert d["value"] > 0 and abs(d["value"] - 8 / (d["ms_per_step"] / 1e3)) <= 1e-6 * d["value"]
    cfg = d["config"]
    assert cfg["actors_per_gpu"] == 8 and cfg["capacity"] == 500 and cfg["k"] == 10 and "workload" in cfg
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("reference", "port") and cb["cores"] >= 1 and cb["value"] == d["value"] and cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["gpu_launches"] == 0


def test_own_arm_needs_a_gpu_and_says_so():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    res = _run("--steps", "2", "--warmup", "3", "--no-cpu-baseline", "--no-secondary", timeout=300)
    assert res.returncode != 0
    assert res.stdout.strip() == ""                     # no JSON line: nothing that could be mistaken for a measurement
    assert "no CPU fallback" in res.stderr
