"""The NCCL transport of the LET exchange on >= 2 GPUs of one box (skipped on a single-GPU box; the same packed
blocks are exchanged device-to-device in tests/test_gpu_multirank.py)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("nranks", [2, 4, 8])
def test_nccl_let_exchange(nranks):
    import torch
    if torch.cuda.device_count() < nranks:
        pytest.skip(f"needs {nranks} GPUs")
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={nranks}", "--master-addr", "127.0.0.1",
                        "--master-port", str(29700 + nranks), os.path.join(ROOT, "tests", "nccl_worker.py")],
                       env=env, capture_output=True, text=True, timeout=600)
    print(r.stdout[-3000:])
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "NCCL_WORKER_OK" in r.stdout


def test_two_devices_one_thread(pn2=None):
    """One host thread driving two contexts on two devices through the host-pointer call (pn2_force_step): the staging
    buffers belong to the context (they were thread-local statics once, on whichever device came first)."""
    import numpy as np
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import sys
    sys.path.insert(0, os.path.join(ROOT, "photons-2.0_b200"))
    import pn2gpu
    pos = np.load(os.path.join(ROOT, "tests", "golden", "demo_pos_f32.npy")).astype(np.float64)
    g = np.load(os.path.join(ROOT, "tests", "golden", "demo_ns32_np1.npz"))
    prm = pn2gpu.make_params(float(g["box"]), 32, len(pos), float(g["mass"]), precision=pn2gpu.FP64)
    c0, c1 = pn2gpu.Context(prm, device=0), pn2gpu.Context(prm, device=1)
    a0 = c0.force_step(pos)                       # device 0 first, the larger set
    sub = pos[::2].copy()
    a1 = c1.force_step(sub)                       # then device 1 with FEWER particles: no regrowth of any buffer
    a1b = c0.force_step(sub)
    a0b = c1.force_step(pos)
    assert np.array_equal(a1, a1b) and np.array_equal(a0, a0b)
    err = float(np.sqrt(((a0 - g["acc"]) ** 2).sum() / (g["acc"] ** 2).sum()))
    assert err < 2e-9, err
    c0.close(); c1.close()
