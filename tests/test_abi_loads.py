"""CPU-side checks of the boundary: libpn2gpu.so loads and exports every symbol include/pn2gpu.h declares,
the struct views match the reference's sizes, and without a GPU the product fails loudly (no fallback)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    txt = open(os.path.join(ROOT, "include", "pn2gpu.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(pn2_[a-z0-9_]+)\s*\(", txt)))


def test_exports_every_declared_symbol(pn2):
    if not os.path.exists(pn2.LIB_PATH):
        pn2.build_library()
    L = pn2.lib()
    syms = header_symbols()
    assert len(syms) >= 28
    for s in syms:
        assert hasattr(L, s), f"libpn2gpu.so does not export {s}"
    assert sorted(pn2.EXPORTS) == syms


def test_struct_sizes_match_reference(pn2):
    # sizes probed from the reference build: Body 96, Pack 376, Node 392, RemoteNode 224, RemoteBody 32
    assert pn2.BODY.itemsize == 96 and pn2.PACK.itemsize == 376 and pn2.NODE.itemsize == 392
    assert pn2.RNODE.itemsize == 224 and pn2.RBODY.itemsize == 32
    assert C.sizeof(pn2.Params) == 64 and C.sizeof(pn2.Domain) == 56


def test_no_cpu_fallback(pn2):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(pn2.Pn2Error) as e:
        pn2.Context(pn2.make_params(100.0, 8, 512, 1.0))
    assert "no CUDA device" in str(e.value)


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "photons-2.0_b200")
    for dp, _, fs in os.walk(pkg):
        for f in fs:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".c", ".cpp")):
                txt = open(os.path.join(dp, f)).read()
                assert "pn_oracle" not in txt and "pn_ref" not in txt and "libpn_oracle" not in txt, f
