"""Drop-in proof (Mode A, C host side): the UNMODIFIED reference (its own tree build, dual-tree walks, LET exchange over
the MPI shim and driver sequence) with every task batch routed through photons-2.0_b200/host/pn2_fmm_glue.c to the
C-ABI of libpn2gpu.so -- oracle/_ref/ref_fmm_gpu, built by `make -C oracle ref` where /root/reference exists.
Its accelerations are compared with the all-CPU reference's golden vectors."""
import numpy as np
import pytest

from conftest import load_golden, rms_rel

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("nside,nranks,precision,tol", [(32, 1, "fp64", 1e-11), (32, 1, "fp32", 3e-5), (16, 1, "fp64", 1e-11),
                                                         (32, 2, "fp64", 1e-11), (32, 2, "fp32", 3e-5)])
def test_reference_with_gpu_batches(demo_pos, nside, nranks, precision, tol):
    import os
    from oracle import pn_ref
    if not os.path.exists(pn_ref.REF_EXE_GPU):
        pytest.skip("oracle/_ref/ref_fmm_gpu not built (needs /root/reference at build time)")
    g = load_golden(f"demo_ns{nside}_np{nranks}.npz")
    ranks = pn_ref.run_reference(demo_pos, float(g["box"]), nside, float(g["mass"]), maxleaf=8, theta=0.4, nranks=nranks, capture=0,
                                 gpu=True, env_extra={"PN2_PRECISION": precision}, timeout=900)
    acc = pn_ref.gather_acc(ranks, len(demo_pos))
    err = rms_rel(acc, g["acc"])
    print(f"reference driver + B200 batches, NSIDE {nside} NP={nranks} {precision}: rms rel err vs all-CPU reference {err:.3e}; "
          f"step {max(r['timing']['total'] for r in ranks):.2f} s")
    assert err < tol
    assert int(sum(r["idxP2P"] for r in ranks)) == int(g["idxP2P"].sum())
