"""The bench.py contract: one JSON line on stdout with the keys the driver reads.  The reference arm runs on the CPU
(here: a small sample); the GPU arm is checked on the B200 at a small size, parity block included."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BASE_KEYS = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
             "dtype", "data", "config", "e2e", "gpu_launches", "cpu_baseline"}


def run_bench(*args, timeout=900):
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True, timeout=timeout, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, lines                       # exactly one line on stdout
    return json.loads(lines[0])


@pytest.mark.skipif(not os.path.exists(os.path.join(ROOT, "oracle", "_ref", "ref_fmm")), reason="oracle/_ref not built")
def test_reference_arm_line():
    d = run_bench("--impl", "reference", "--steps", "1", "--warmup", "0", "--cpu-sample-side", "24")
    assert BASE_KEYS <= set(d), BASE_KEYS - set(d)
    assert d["impl"] == "reference" and d["gpu_launches"] == 0 and d["higher_is_better"] is True
    assert d["unit"] == "particles/s" and d["value"] > 0 and d["dtype"] == "f64"
    # the arm states the sample it really ran, and what it is a sample of
    assert d["config"]["npart"] == 24 ** 3 and d["config"]["nside"] == 24 and "sample_of" in d["config"]
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("reference", "port") and cb["cores"] >= 1 and cb["value"] == d["value"] and cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": "particles/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


@pytest.mark.gpu
def test_gpu_arm_line_small():
    d = run_bench("--npart-side", "64", "--steps", "2", "--warmup", "3", "--cpu-sample-side", "32")
    assert BASE_KEYS <= set(d), BASE_KEYS - set(d)
    assert d["n_gpus"] == 1 and d["gpu_launches"] > 0 and d["value"] > 0 and d["dtype"] == "f32"
    rf = d["roofline"]
    assert {"bound", "achieved", "peak", "unit", "frac", "traffic"} <= set(rf) and 0 < rf["frac"] < 1
    assert d["e2e"]["h2d_bytes_per_step"] == 64 ** 3 * 24 and d["e2e"]["d2h_bytes_per_step"] == 64 ** 3 * 24
    assert d["e2e"]["max_abs_diff_vs_device_step_rank0"] == 0.0
    assert d["cpu_baseline"]["kind"] == "reference" and d["cpu_baseline"]["value"] > 0
    p = d["parity"]
    assert p["pass"] and p["fp64_rms"] <= 1e-6 and p["fp32_rms"] <= 1e-4 and p["nleaf_equal"] and p["nint_equal"], p
    assert d["momentum_residual"] < 1e-6
    assert set(d["clocks"]) >= {"sm_mhz", "sm_max_mhz", "reasons"}


@pytest.mark.skipif(not os.path.exists(os.path.join(ROOT, "oracle", "_ref", "ref_fmm")), reason="oracle/_ref not built")
def test_reference_arm_rank_count_guard(monkeypatch):
    """The unmodified reference does not return when a domain is thinner than twice the cut-off radius (its LET exchange
    assumes one meeting per peer through the periodic wrap: 32^3 / NSIDE 32 at 16 ranks never ends).  The CPU arm halves
    the rank count until src/domains.c's boxes are wide enough, and says how many ranks it really used."""
    import importlib
    sys.path.insert(0, ROOT)
    monkeypatch.setattr(sys, "argv", ["bench.py", "--cpu-sample-side", "32"])
    bench = importlib.import_module("bench")
    args = bench.parse()
    r = bench.reference_cpu_run(args, 32, 16)
    assert r["kind"] == "reference" and r["cores"] == 8 and r["n"] == 32 ** 3 and r["pps"] > 0
