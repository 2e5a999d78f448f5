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
