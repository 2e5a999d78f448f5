"""The C-ABI driven from plain C (examples/force_step.c, compiled with gcc against include/pn2gpu.h): one Mode B
force step of every 8th demo particle; the accelerations it writes are bit-identical to the Python binding's (same
library, same inputs) and within tolerance of the unmodified reference's golden accelerations."""
import os
import subprocess

import numpy as np
import pytest

from conftest import load_golden, rms_rel

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "examples", "force_step")


@pytest.mark.parametrize("mode,tol", [("fp64", 1e-9), ("fp32", 1e-4)])
def test_c_program_matches_python_binding(pn2, small_pos, tmp_path, mode, tol):
    if not os.path.exists(EXE):
        subprocess.run(["make", "-C", os.path.join(ROOT, "examples")], check=True)
    g = load_golden("small_t04_np1.npz")
    n = len(small_pos)
    box, nside, mass = float(g["box"]), int(g["nside"]), float(g["mass"])
    pin, pout = tmp_path / "pos.f64", tmp_path / "acc.f64"
    small_pos.astype(np.float64).tofile(pin)
    r = subprocess.run([EXE, str(pin), str(pout), str(n), repr(box), str(nside), repr(mass), mode], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    print(r.stdout.strip())
    acc_c = np.fromfile(pout, np.float64).reshape(n, 3)
    prm = pn2.make_params(box, nside, n, mass, maxleaf=8, theta=0.4, precision=pn2.FP64 if mode == "fp64" else pn2.FP32)
    acc_py = pn2.Context(prm).force_step(small_pos)
    np.testing.assert_array_equal(acc_c, acc_py)
    assert rms_rel(acc_c, g["acc"]) < tol
