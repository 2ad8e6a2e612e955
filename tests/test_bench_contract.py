"""bench.py's JSON contract.  The reference arm needs no GPU, so its line is checked here on the
CPU; the GPU arm's line is checked by the gpu-marked test below."""
import json
import os
import subprocess
import sys

import pytest

REPO = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
BASE_KEYS = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
             "vs_baseline", "dtype", "data", "config", "e2e", "gpu_launches"}


def run_bench(*args, timeout=600):
    p = subprocess.run([sys.executable, os.path.join(REPO, "bench.py"), *args], capture_output=True, text=True,
                       timeout=timeout, cwd=REPO)
    assert p.returncode == 0, p.stderr[-3000:]
    lines = [l for l in p.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1, p.stdout
    return json.loads(lines[0])


def test_reference_arm_line():
    d = run_bench("--impl", "reference", "--n", "400000", "--steps", "2", "--warmup", "1")
    assert BASE_KEYS <= set(d) and d["impl"] == "reference"
    assert d["metric"] == "fp64 objective+gradient evals/s (N obs)" and d["unit"] == "evals/s"
    assert d["dtype"] == "f64" and d["vs_baseline"] is None and d["higher_is_better"] is True
    assert d["e2e"] == {"value": d["value"], "unit": "evals/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("reference", "port") and cb["cores"] >= 1 and cb["value"] == d["value"]
    assert "workload" in d["config"] and "model" not in d["config"]
    assert 0 < d["value"] < 1e3 and abs(d["ms_per_step"] - 1e3 / d["value"]) < 1e-6 * d["ms_per_step"]


def test_reference_arm_other_ranks_exit_quietly():
    env = dict(os.environ, RANK="1", LOCAL_RANK="1", WORLD_SIZE="2")
    p = subprocess.run([sys.executable, os.path.join(REPO, "bench.py"), "--impl", "reference", "--gpus", "2"],
                       capture_output=True, text=True, timeout=60, env=env, cwd=REPO)
    assert p.returncode == 0 and p.stdout.strip() == ""


@pytest.mark.gpu
def test_gpu_arm_line():
    d = run_bench("--n", "2000000", "--steps", "5", "--warmup", "3", "--config5-items", "1000000")
    assert BASE_KEYS | {"roofline", "cpu_baseline", "clocks"} <= set(d) and "impl" not in d
    assert d["n_gpus"] == 1 and d["steps"] == 5 and d["warmup"] == 3 and d["gpu_launches"] == 5
    r = d["roofline"]
    assert r["bound"] == "hbm" and r["unit"] == "GB/s" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-12
    assert 0.3 < r["frac"] < 3.0 and r["kernel_launches_timed"] == 5
    assert d["e2e"]["h2d_bytes_per_step"] == 800 and d["e2e"]["d2h_bytes_per_step"] == 832
    assert d["launch"]["launches_per_eval"] == 1 and d["launch"]["tape_tier"] == "hbm"
    assert d["parity_check"]["status"] in ("ok", "no golden")
    rows = d["extra"]["configs"]                             # every other BASELINE config, one row each
    assert len(rows) >= 8 and not [r for r in rows if "error" in r], rows
    for r in rows:
        assert r["parity_rel_err"] is not None and r["parity_rel_err"] <= 1e-12, r
        assert r["roofline"]["frac"] > 0
    assert {r["tape_tier"] for r in rows if r["kernel"] in ("linreg2_p", "regression_linear", "regression_logistic")} == {"register"}
    assert len(d["extra"]["config5_at_1e8"]) == 3 and all(r["parity_rel_err"] <= 1e-12 for r in d["extra"]["config5_at_1e8"])
    assert 0.5 * d["value"] < d["e2e"]["value"] <= 1.05 * d["value"]
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["value"] < d["value"] / 100
    ocl = d["cpu_baseline"]["opencl_on_this_gpu"]            # context: the unmodified OpenCL path on this GPU, if it has one
    assert "unavailable" in ocl or 0 < ocl["value"] < d["value"]
    assert set(d["clocks"]) >= {"sm_mhz", "sm_max_mhz", "reasons"}
    assert abs(d["value"] - 1e3 / d["ms_per_step"]) < 1e-9 * d["value"]
