"""Multi-rank Mode B on the GPU: domain decomposition (src/domains.c rule), LET prune / pack / exchange / unpack,
walks against received trees.  Ranks are contexts of one process exchanging device-to-device
(pn2_exchange_local), so this runs on a single GPU; the NCCL transport moves the same packed blocks
(tests/test_nccl_two_gpus.py, bench.py --gpus N).

Checked against (1) the oracle's multi-rank evaluation on Mode-B trees (reference-pinned walkers / LET pack),
(2) the UNMODIFIED reference's NP = 2, 3, 4 and 8 golden accelerations of the demo IC (src/remotes.c:684-751,
src/fmm.c:1021-1045 at the rank counts BASELINE.md's runs used)."""
import numpy as np
import pytest

from conftest import load_golden
from modeb_check import oracle_step_multirank, rms_rel

pytestmark = pytest.mark.gpu
TOL = {0: 1e-6, 1: 1e-4, 2: 1e-6}        # 0 = PN2_FP64 (table kernel), 1 = PN2_FP32, 2 = PN2_FP64_LIBM
MODES = (2, 0, 1)


def split(pn2, pos, nranks, box):
    import domains
    doms = domains.domain_boxes(nranks, box)
    owner = domains.domain_of(pos, nranks, box)
    idx = [np.nonzero(owner == r)[0] for r in range(nranks)]
    return doms, idx


def run_local_ranks(pn2, prm_o, precision, pos, nranks):
    doms, idx = split(pn2, pos, nranks, prm_o.box)
    ctxs = [pn2.Context(pn2.Params(prm_o.box, prm_o.rs, prm_o.cutoff, prm_o.soft, prm_o.theta, prm_o.mass, prm_o.maxleaf,
                                   prm_o.periodic, prm_o.longshort, precision)) for _ in range(nranks)]
    accs = pn2.force_step_local_ranks(ctxs, [pos[i] for i in idx], doms)
    acc = np.zeros_like(pos)
    for r in range(nranks):
        acc[idx[r]] = accs[r]
    infos = [c.step_info() for c in ctxs]
    for c in ctxs:
        c.close()
    return acc, infos, doms, idx


@pytest.mark.parametrize("tag,nranks", [("t04", 2), ("t12", 2), ("t12", 3), ("t04", 4)])
def test_small_vs_oracle(pn2, oracle, small_pos, tag, nranks):
    g = load_golden(f"small_{tag}_np1.npz")
    box = float(g["box"])
    prm_o = oracle.make_params(box, int(g["nside"]), len(small_pos), float(g["mass"]), theta=float(g["theta"]))
    doms, idx = split(pn2, small_pos, nranks, box)
    boxes = [(list(d.lo), list(d.hi)) for d in doms]
    ref_accs, ref_nint = oracle_step_multirank(oracle, [small_pos[i] for i in idx], boxes, [d.direct0 for d in doms], prm_o)
    ref = np.zeros_like(small_pos)
    for r in range(nranks):
        ref[idx[r]] = ref_accs[r]
    for precision in MODES:
        acc, infos, _, _ = run_local_ranks(pn2, prm_o, precision, small_pos, nranks)
        err = rms_rel(acc, ref)
        print(f"small {tag} NP={nranks} precision {precision}: rms rel err vs oracle {err:.3e}; LET cells {[i['n_let_nodes'] for i in infos]}")
        assert [i["n_interactions"] for i in infos] == ref_nint        # identical lists -> identical counts
        assert err < TOL[precision] and err < {2: 1e-11, 0: 2e-9, 1: 3e-5}[precision]
        if tag == "t12":
            assert sum(i["n_m2l_pairs"] for i in infos) > 1000


@pytest.mark.parametrize("nside,nranks", [(32, 2), (32, 3), (32, 4), (32, 8), (16, 2), (16, 4)])
def test_demo_vs_reference_golden(pn2, oracle, demo_pos, nside, nranks):
    """The reference itself at NP = 2, 3, 4, 8 (tests/golden/demo_ns*_np*.npz): same domain rule, same trees' leaf sets,
    same lists -> FP64 mode agrees to rounding, FP32 mode within 1e-4."""
    g = load_golden(f"demo_ns{nside}_np{nranks}.npz")
    prm_o = oracle.make_params(float(g["box"]), nside, len(demo_pos), float(g["mass"]), maxleaf=8, theta=0.4)
    for precision in MODES:
        acc, infos, _, _ = run_local_ranks(pn2, prm_o, precision, demo_pos, nranks)
        err = rms_rel(acc, g["acc"])
        print(f"demo nside {nside} NP={nranks} precision {precision}: rms rel err vs reference golden {err:.3e}")
        assert sum(i["n_interactions"] for i in infos) == int(g["nint_local"].sum() + g["p2p_count_remote"].sum())
        assert err < TOL[precision] and err < {2: 1e-9, 0: 2e-9, 1: 3e-5}[precision]
