"""bench.py's contract with the driver, as far as it can be checked without a GPU: the reference arm prints one JSON line
with the keys the driver reads, and the GPU arm refuses to run without a CUDA device (no CPU fallback)."""

import json
import os
import subprocess
import sys

import pytest

import oracle

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run_bench(*arguments):
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + list(arguments), capture_output=True, text=True, cwd=ROOT, timeout=600)


def test_reference_arm_prints_the_contract_line():
    proc = run_bench("--impl", "reference", "--steps", "1", "--warmup", "0", "--cpu-nprod", "40", "--cpu-threads", "2")
    assert proc.returncode == 0, proc.stderr
    lines = [line for line in proc.stdout.splitlines() if line.startswith("{")]
    assert len(lines) == 1
    line = json.loads(lines[0])
    assert line["impl"] == "reference" and line["metric"] == "mc_trial_moves_per_s" and line["unit"] == "moves/s"
    assert line["higher_is_better"] is True and line["n_gpus"] == 1 and line["steps"] == 1 and line["gpu_launches"] == 0
    assert line["value"] > 0 and line["ms_per_step"] > 0
    assert line["cpu_baseline"]["kind"] == ("reference" if oracle.have_ref() else "port")
    assert line["cpu_baseline"]["cores"] == 2 and line["cpu_baseline"]["value"] == line["value"]
    assert line["e2e"] == {"value": line["value"], "unit": "moves/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "lysozyme" in line["config"]["workload"]


def test_reference_arm_ranks_other_than_zero_do_nothing():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    proc = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0"],
                          capture_output=True, text=True, cwd=ROOT, env=env, timeout=120)
    assert proc.returncode == 0 and proc.stdout.strip() == ""


def test_gpu_arm_has_no_cpu_fallback():
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present: the refusal path cannot be exercised")
    proc = run_bench("--steps", "1", "--warmup", "0")
    assert proc.returncode != 0
    assert "CUDA" in (proc.stderr + proc.stdout)
