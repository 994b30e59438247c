"""CPU: bench.py's reference arm prints ONE JSON line with the contract's keys, and the product arm refuses to run
without a GPU (no CPU fallback)."""
import json
import os
import subprocess
import sys

import pytest
import torch

from conftest import ROOT

BENCH = os.path.join(ROOT, "bench.py")


def _run(*args, timeout=600):
    env = dict(os.environ, OMP_NUM_THREADS=str(min(8, os.cpu_count() or 1)))
    return subprocess.run([sys.executable, BENCH, *args], capture_output=True, text=True, timeout=timeout, cwd=ROOT, env=env)


@pytest.mark.parametrize("workload", ["render", "train"])
def test_reference_arm_line(workload):
    p = _run("--impl", "reference", "--gpus", "1", "--steps", "1", "--warmup", "0", "--workload", workload)
    assert p.returncode == 0, p.stderr[-2000:]
    lines = [ln for ln in p.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, lines                                    # exactly one line on stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["n_gpus"] == 1 and d["steps"] == 1 and d["warmup"] == 0
    assert d["unit"] == "rays/s" and d["higher_is_better"] is True and d["vs_baseline"] is None and d["data"] == "synthetic"
    assert d["value"] > 0 and d["ms_per_step"] > 0 and d["dtype"] == "f32" and d["scaling"] == "weak"
    assert isinstance(d["config"]["workload"], str) and "model" not in d["config"]
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and isinstance(cb["sample"], str)
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    if workload == "render":
        assert d["metric"].startswith("rays/sec (64+128 samples, 2x SS)")
        assert d["config"]["rays_per_step_per_gpu"] == 160000 and d["config"]["supersampling"] == 2


def test_reference_arm_nonzero_ranks_exit_without_work():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    p = subprocess.run([sys.executable, BENCH, "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0"],
                       capture_output=True, text=True, timeout=120, cwd=ROOT, env=env)
    assert p.returncode == 0 and p.stdout.strip() == ""


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU behaviour")
def test_product_arm_fails_loudly_without_a_gpu():
    p = _run("--steps", "1", timeout=120)
    assert p.returncode != 0 and p.stdout.strip() == ""
    assert "no GPU visible" in p.stderr and "no CPU fallback" in p.stderr
