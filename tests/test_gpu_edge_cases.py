"""Edge cases of the Mode B step on the device: an empty rank in a multi-rank step, the loud failures (leaves too wide
for the FP32 tile layout; a domain box that differs from the table the peers hold), MAXLEAF 1 and 2."""
import numpy as np
import pytest

from conftest import load_golden
from modeb_check import oracle_step_on_tree, rms_rel

pytestmark = pytest.mark.gpu


def test_empty_rank(pn2, oracle, small_pos):
    """Two ranks, all particles in the first rank's domain: the empty rank packs and receives nothing, the other one's
    forces are those of its own tree in its own (half) box plus the 26 images -- the oracle's one-rank evaluation there."""
    import domains
    g = load_golden("small_t04_np1.npz")
    box = float(g["box"])
    sub = small_pos[small_pos[:, 0] < 0.5 * box].copy()
    doms = domains.domain_boxes(2, box)
    owner = domains.domain_of(sub, 2, box)
    assert (owner == 0).all()
    prm_o = oracle.make_params(box, int(g["nside"]), len(sub), float(g["mass"]), theta=0.4)
    lo, hi = np.array(list(doms[0].lo)), np.array(list(doms[0].hi))
    tb = oracle.TreeB(sub, 8, list(lo), list(hi), direct0=doms[0].direct0)
    ref = oracle_step_on_tree(oracle, tb, prm_o, 0.5 * (lo + hi), hi - lo)
    ref_acc = np.zeros_like(ref["acc"])
    ref_acc[tb.ids] = ref["acc"]
    for precision, tol in ((pn2.FP64, 2e-9), (pn2.FP32, 3e-5)):
        ctxs = [pn2.Context(pn2.Params(prm_o.box, prm_o.rs, prm_o.cutoff, prm_o.soft, prm_o.theta, prm_o.mass, 8, 1, 1, precision)) for _ in range(2)]
        accs = pn2.force_step_local_ranks(ctxs, [sub, np.zeros((0, 3))], doms)
        assert accs[1].shape == (0, 3)
        err = rms_rel(accs[0], ref_acc)
        print(f"empty rank, precision {precision}: rms rel err {err:.2e}")
        assert err < tol and ctxs[0].step_info()["n_interactions"] == ref["nint"]
        for c in ctxs:
            c.close()


def test_wide_leaves_fail_loudly_in_fp32(pn2):
    """64 particles in the whole box at NSIDE 64: leaves are ~20 rs wide, far beyond what the FP32 tile layout of the
    long/short split can pad safely -- the step must refuse (and FP64 mode must work)."""
    rng = np.random.default_rng(4)
    pos = rng.random((64, 3)) * 1000.0
    ctx = pn2.Context(pn2.make_params(1000.0, 64, 64, 1.0, precision=pn2.FP32))
    with pytest.raises(pn2.Pn2Error, match="PN2_FP64"):
        ctx.force_step(pos)
    ctx.close()
    ctx = pn2.Context(pn2.make_params(1000.0, 64, 64, 1.0, precision=pn2.FP64))
    assert np.isfinite(ctx.force_step(pos)).all()
    ctx.close()


def test_stale_domain_table_is_refused(pn2, small_pos):
    import torch
    import domains
    doms = domains.domain_boxes(2, 100000.0)
    ctx = pn2.Context(pn2.make_params(100000.0, 24, len(small_pos), 2.5))
    ctx.set_comm(0, 2, doms, None)
    other = pn2.make_domain([0, 0, 0], [40000.0, 100000.0, 100000.0], doms[0].direct0)
    t = torch.from_numpy(small_pos[:100].copy()).cuda()
    with pytest.raises(pn2.Pn2Error, match="all_domains"):
        ctx.step_begin(t.data_ptr(), 100, other)
    ctx.close()


@pytest.mark.parametrize("maxleaf", [1, 2, 3])
def test_small_maxleaf(pn2, oracle, small_pos, maxleaf):
    """MAXLEAF 1..3: more leaves than n / 2 (the record capacity follows 2 n / MAXLEAF); tree bit-exact, forces within tolerance."""
    from test_gpu_mode_b import check_tree
    g = load_golden("small_t04_np1.npz")
    box = float(g["box"])
    pos = small_pos[:1500].copy()
    prm_o = oracle.make_params(box, int(g["nside"]), len(pos), float(g["mass"]), maxleaf=maxleaf, theta=0.4)
    tb = oracle.TreeB(pos, maxleaf, [0, 0, 0], [box] * 3)
    ref = oracle_step_on_tree(oracle, tb, prm_o, np.array([0.5 * box] * 3), np.array([box] * 3))
    ref_acc = np.zeros_like(ref["acc"])
    ref_acc[tb.ids] = ref["acc"]
    ctx = pn2.Context(pn2.Params(prm_o.box, prm_o.rs, prm_o.cutoff, prm_o.soft, prm_o.theta, prm_o.mass, maxleaf, 1, 1, pn2.FP64))
    acc = ctx.force_step(pos)
    check_tree(ctx, tb)
    assert rms_rel(acc, ref_acc) < 2e-9
    ctx.close()
