"""CPU: the parts of bench.py's contract that do not need a GPU — the reference arm's JSON line (BASELINE.json's metric,
unit and config; `cpu_baseline` and `e2e` objects), rank > 0 of a torchrun launch of the reference arm leaving without
work, and the own arm refusing to run without a CUDA device (no CPU fallback)."""
import json
import os
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BENCH = os.path.join(ROOT, "bench.py")


def _run(args, env_extra=None, timeout=300):
    env = dict(os.environ)
    for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK"):
        env.pop(k, None)
    env.update(env_extra or {})
    return subprocess.run([sys.executable, BENCH] + args, capture_output=True, text=True, timeout=timeout, cwd=ROOT, env=env)


def test_reference_arm_prints_the_contract_line():
    base = json.load(open(os.path.join(ROOT, "BASELINE.json")))
    out = _run(["--impl", "reference", "--steps", "1", "--warmup", "0", "--sample", "800"])
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["higher_is_better"] is True and line["scaling"] == "weak"
    assert line["unit"] == "patches/s" and line["value"] > 0 and line["steps"] == 1 and line["n_gpus"] == 1
    assert line["vs_baseline"] is None and line["data"] == "synthetic" and "workload" in line["config"]
    assert "model" not in line["config"]
    cb = line["cpu_baseline"]
    assert cb["kind"] in ("port", "reference") and cb["cores"] >= 1 and cb["value"] == line["value"] and "800" in cb["sample"]
    assert line["e2e"] == {"value": line["value"], "unit": line["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert line["gpu_launches"] == 0
    # the metric is the one BASELINE.json names (its text carries the unit)
    assert "patches" in json.dumps(base).lower()


def test_reference_arm_only_rank_zero_works():
    out = _run(["--impl", "reference", "--steps", "1", "--warmup", "0", "--sample", "800"], {"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"})
    assert out.returncode == 0 and out.stdout.strip() == ""


def test_own_arm_needs_a_gpu():
    if torch.cuda.is_available():
        pytest.skip("needs a box without a GPU")
    out = _run(["--steps", "1", "--warmup", "0", "--no-cpu-baseline", "--no-extras"])
    assert out.returncode != 0 and "no CUDA device" in (out.stderr + out.stdout)
