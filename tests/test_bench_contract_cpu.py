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


def test_render_roofline_object():
    """The roofline object of the product arm (pure function): algorithmic FLOP of the step's two k_tc_pass launches over
    the event-timed step, against the sustained measured peak; the fine pass timed alone next to it."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("bench_module", BENCH)
    b = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(b)
    pk = {"bf16_tflops": 1652.1, "bf16_tflops_sustained": 1386.6, "hbm_gbs": 6550.7}
    r = b.render_roofline("bf16x3", 160000, 10, 708.297, 47.049, pk, "measured", 2.0)   # the numbers of profiles/r01_bench_bf16x3.json
    assert r["bound"] == "tensor" and r["unit"] == "TFLOP/s" and r["peak"] == 1386.6 and r["launches_per_step"] == 2
    assert r["flop_per_launch"] * 2 == 160000 * 192 * 1186816                          # 227.87 MFLOP per ray
    assert r["achieved"] == pytest.approx(r["flop_per_launch"] / (r["ms_per_launch"] / 1e3) / 1e12)
    assert r["achieved"] == pytest.approx(514.74, rel=1e-3) and r["frac"] == pytest.approx(r["achieved"] / 1386.6)
    # `traffic` is never a hard-coded constant: it is this round's committed ncu capture (profiles/r02_ncu_traffic.json) or null
    assert r["issued_frac"] == pytest.approx(3 * r["frac"]) and r["traffic"] == b.ncu_traffic("k_tc_pass", "mean_bytes_per_launch")
    assert (r["traffic"] is None) == (r["traffic_source"] is None)
    assert r["algorithmic_bytes_per_launch"] == 160000 * 564
    f = r["fine_pass_alone"]
    assert f["flop_per_launch"] == 160000 * 128 * 1186816 and f["achieved"] == pytest.approx(516.6, rel=1e-3)
    assert f["frac_vs_burst_peak"] == pytest.approx(f["achieved"] / 1652.1) and f["frac_vs_sustained_peak"] == pytest.approx(f["achieved"] / 1386.6)
    # one launch per frame (round 2): the same step FLOP in one launch; traffic from the frame variant's capture
    o = b.render_roofline("bf16x3", 160000, 10, 708.297, 47.049, pk, "measured", 1.0)
    assert o["launches_per_step"] == 1 and o["flop_per_launch"] == 160000 * 192 * 1186816 and o["ms_per_launch"] == pytest.approx(70.8297)
    assert o["achieved"] == pytest.approx(r["achieved"]) and "one launch" in o["kernel"]
    assert o["traffic"] == b.ncu_traffic("k_tc_pass", "frame_bytes_per_launch") and o["algorithmic_bytes_per_launch"] == 160000 * 80
    s = b.render_roofline("fp32_simt", 160000, 2, 2400.0, 800.0, {"bf16_tflops": 1590.0}, "fallback", 7.0)   # no sustained figure
    assert s["kernel"] == "k_simt_mlp" and s["traffic"] is None and s["peak"] == 1590.0 and s["issued_frac"] == pytest.approx(s["frac"])
    import json
    json.dumps(r), json.dumps(s)


def test_train_roofline_object():
    import importlib.util
    import json
    spec = importlib.util.spec_from_file_location("bench_module2", BENCH)
    b = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(b)
    assert b.TRAIN_BYTES_PER_POINT == 44084
    pk = {"bf16_tflops": 1652.1, "bf16_tflops_sustained": 1386.6, "hbm_gbs": 6550.7}
    r = b.train_roofline(2048, 20, 20 * 4.8617, pk, "measured")                      # profiles/r01_bench_train_1gpu.json
    assert r["bound"] == "hbm" and r["unit"] == "GB/s" and r["peak"] == 6550.7 and r["traffic"] is None
    # the dW partials (written + read back, independent of the batch) are part of the step's bytes
    assert r["dw_partial_bytes_per_step"] == b.dw_partial_bytes(2048) and 0.7e9 < r["dw_partial_bytes_per_step"] < 0.9e9
    assert r["bytes_per_step"] == 2048 * 192 * 44084 + r["dw_partial_bytes_per_step"]
    assert r["achieved"] == pytest.approx(r["bytes_per_step"] / 4.8617e-3 / 1e9)
    assert r["frac"] == pytest.approx(r["achieved"] / 6550.7)
    t = r["tensor"]
    assert t["flop_per_step"] == 2048 * 192 * 1186816 * 3 and t["frac"] == pytest.approx(t["achieved"] / 1386.6) and t["issued_frac"] == pytest.approx(3 * t["frac"])
    json.dumps(r)
