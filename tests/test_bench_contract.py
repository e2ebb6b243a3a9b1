"""bench.py's contract, as far as it can be checked without a GPU.

* `--impl reference` (the CPU arm the driver runs beside ours) prints ONE JSON line with the contract's keys, the same
  metric / unit / config as our arm, and describes its sample.
* Our own arm refuses to run without a CUDA device: there is no CPU fallback behind the product path.
"""
import json
import os
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(*args, timeout=600):
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], cwd=ROOT, capture_output=True,
                          text=True, timeout=timeout)


def test_reference_arm_prints_one_contract_line():
    r = _run("--impl", "reference", "--steps", "1", "--warmup", "0")
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference"
    assert d["metric"] == "smpl_forward_bodies_per_sec" and d["unit"] == "bodies/s" and d["higher_is_better"] is True
    assert d["n_gpus"] == 1 and d["steps"] == 1 and d["vs_baseline"] is None and d["data"] == "synthetic"
    assert d["config"]["bodies_per_step_per_gpu"] == 4096 and "configs[2]" in d["config"]["workload"]
    assert d["value"] > 0 and abs(d["ms_per_step"] * d["value"] / 1e3 - 4096) < 1.0     # the WHOLE 4096-body batch per step
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and "4096 bodies per step" in cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": "bodies/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_own_arm_fails_loudly_without_a_gpu():
    r = _run("--steps", "1", "--warmup", "3", timeout=300)
    assert r.returncode != 0
    assert "no CPU fallback" in (r.stderr + r.stdout)
    assert not [l for l in r.stdout.splitlines() if l.startswith("{")]      # and prints no bench line
