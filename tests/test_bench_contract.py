"""bench.py's output contract, checked on the arm that needs no GPU (--impl reference: the C port of the reference
on the host cores): one JSON line with the keys the driver reads."""
import json
import os
import subprocess
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_contract_line():
    p = subprocess.run([sys.executable, os.path.join(REPO, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                        "--cpu-members-per-core", "1", "--no-python-reference"], capture_output=True, text=True, timeout=600, cwd=REPO)
    assert p.returncode == 0, p.stderr[-2000:]
    lines = [ln for ln in p.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"].startswith("Preissmann node-steps/s") and d["unit"] == "node-steps/s"
    assert d["higher_is_better"] is True and d["scaling"] == "weak" and d["dtype"] == "f64" and d["vs_baseline"] is None
    assert d["n_gpus"] == 1 and d["steps"] == 1 and d["value"] > 0 and d["ms_per_step"] > 0
    assert "workload" in d["config"] and "model" not in d["config"]
    cb, e2e = d["cpu_baseline"], d["e2e"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and cb["sample"]
    assert e2e["value"] == d["value"] and e2e["h2d_bytes_per_step"] == 0 and e2e["d2h_bytes_per_step"] == 0
    assert d["gpu_launches"] == 0 and cb["per_core"] * cb["cores"] == cb["value"]


import pytest


def test_python_reference_leg_runs_the_staged_reference():
    """cpu_baseline kind "reference": two members of the calibration grid through the unmodified Python reference staged by
    oracle/stage_reference.py (skipped where it is not staged, e.g. a fresh clone without /root/reference)."""
    sys.path.insert(0, REPO)
    sys.path.insert(0, os.path.join(REPO, "oracle"))
    import ref_harness

    if not ref_harness.reference_available():
        pytest.skip("reference not staged")
    import bench

    r = bench.python_reference_sample(65536, 2, 121, 32)
    assert r["kind"] == "reference" and r["cores"] == 2 and 20 < r["per_core"] < 5000      # ~100 node-steps/s per core


@pytest.mark.gpu
def test_gpu_arm_prints_one_contract_line():
    p = subprocess.run([sys.executable, os.path.join(REPO, "bench.py"), "--members", "512", "--steps", "2", "--warmup", "3",
                        "--cpu-members-per-core", "1", "--no-python-reference", "--config5-nodes", "2001",
                        "--config5-members", "8"], capture_output=True, text=True, timeout=900, cwd=REPO)
    assert p.returncode == 0, p.stderr[-2000:]
    lines = [ln for ln in p.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert "impl" not in d or d["impl"] != "reference"
    assert d["unit"] == "node-steps/s" and d["n_gpus"] == 1 and d["steps"] == 2 and d["warmup"] == 3 and d["value"] > 0
    assert d["higher_is_better"] is True and d["scaling"] == "weak" and d["dtype"] == "f64" and d["data"] == "synthetic"
    assert d["gpu_launches"] == 2 * 3                       # GVF + Newton + objective kernels per step
    e2e, roof, cb, clk = d["e2e"], d["roofline"], d["cpu_baseline"], d["clocks"]
    assert e2e["value"] > 0 and e2e["h2d_bytes_per_step"] == 512 * 8 and e2e["d2h_bytes_per_step"] > 0
    assert roof["unit"] == "TFLOP/s" and 0 < roof["frac"] < 1 and abs(roof["frac"] - roof["achieved"] / roof["peak"]) < 1e-12
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] > 0
    assert "sm_mhz" in clk and "reasons" in clk
    par = d["parity"]
    assert par["iterations_equal"] is True and par["max_rel_depth"] < 1e-9 and par["max_rel_flow"] < 1e-9
    # the blocks beside the contract's keys
    assert d["e2e_full"]["d2h_bytes_per_step"] == 512 * (12 + 32 * 4 + 2 * 33 * 8) and d["e2e_full"]["value"] > 0
    st = d["strong"]
    assert st["members_total"] == 65536 and st["value"] > 0
    c5 = d["config5"]
    assert c5["roofline"]["bound"] == "hbm" and 0 < c5["roofline"]["frac"] < 1 and c5["newton_trips"] > 0
    assert c5["parity"]["iterations_equal"] is True and c5["parity"]["max_rel_depth"] < 1e-9 and c5["host_setup_s"] < 5
    sr = d["single_runs"]
    assert sr["example"]["newton_iterations"] == 87 and sr["akbari_firoozi"]["newton_iterations"] == 67
    assert sr["gerd_roseires"]["newton_iterations"] == 5610 and sr["gerd_roseires"]["run_seconds"] < 60
