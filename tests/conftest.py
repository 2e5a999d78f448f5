import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")
PKG = os.path.join(ROOT, "photons-2.0_b200")
if PKG not in sys.path:
    sys.path.insert(0, PKG)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def rms_rel(a, b):
    """rms relative error of vector field a against reference b: sqrt(sum|a-b|^2 / sum|b|^2)."""
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    return float(np.sqrt(((a - b) ** 2).sum() / (b ** 2).sum()))


@pytest.fixture(scope="session")
def demo_pos():
    """Positions of the reference's demo/ic_lcdm.gdt2 (float32 values, returned as float64)."""
    return np.load(os.path.join(GOLDEN, "demo_pos_f32.npy")).astype(np.float64)


@pytest.fixture(scope="session")
def small_pos(demo_pos):
    return demo_pos[::8].copy()


def load_golden(name):
    return np.load(os.path.join(GOLDEN, name))


@pytest.fixture(scope="session")
def oracle():
    from oracle import pn_oracle
    pn_oracle.lib()
    return pn_oracle


@pytest.fixture(scope="session")
def pn2():
    """The product's host binding (photons-2.0_b200/pn2gpu.py over libpn2gpu.so)."""
    import pn2gpu
    return pn2gpu


def image_shifts(box):
    """The 26 periodic displacements in the reference's order (src/fmm.c:1028-1037)."""
    return [((q // 9 - 1) * box, ((q // 3) % 3 - 1) * box, (q % 3 - 1) * box) for q in range(27) if q != 13]
