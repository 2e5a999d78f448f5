"""Pins the CPU oracle (oracle/port, our restatement) to golden vectors produced by the UNMODIFIED
reference (tests/golden/make_golden.py): accelerations, counters, tree arrays, interaction lists,
multipoles, local expansions and the pruned LET trees.  Runs on CPU, everywhere."""
import hashlib

import numpy as np
import pytest

from conftest import load_golden, rms_rel


def _hash(s, t):
    h = hashlib.sha256()
    h.update(np.ascontiguousarray(s, np.int32).tobytes())
    h.update(np.ascontiguousarray(t, np.int32).tobytes())
    return h.hexdigest()


@pytest.mark.parametrize("name,nranks", [("demo_ns32_np1", 1), ("demo_ns32_np2", 2), ("demo_ns32_np3", 3), ("demo_ns32_np4", 4),
                                         ("demo_ns32_np8", 8), ("demo_ns16_np1", 1), ("demo_ns16_np2", 2), ("demo_ns16_np4", 4)])
def test_force_matches_reference(oracle, demo_pos, name, nranks):
    g = load_golden(name + ".npz")
    prm = oracle.make_params(float(g["box"]), int(g["nside"]), len(demo_pos), float(g["mass"]), maxleaf=int(g["maxleaf"]),
                             theta=float(g["theta"]))
    acc, cnt = oracle.force(demo_pos, prm, nranks)
    assert cnt["p2p_pairs"] == int(g["idxP2P"].sum())
    assert cnt["m2l_pairs"] == int(g["idxM2L"].sum())
    assert cnt["int_local"] == int(g["nint_local"].sum())
    assert cnt["int_remote"] == int(g["p2p_count_remote"].sum())
    assert cnt["m2l_calls"] == int(g["walk_m2l_count"].sum())
    assert cnt["leaves"] == int((g["last_leaf"] - g["first_leaf"]).sum())
    # same algorithm, same lists, libm-level differences only
    assert rms_rel(acc, g["acc"]) < 1e-12


def test_demo_known_answers(oracle, demo_pos):
    """Known answers of SURVEY.md 8c (demo lcdm_g2.run, NP=1)."""
    g = load_golden("demo_ns32_np1.npz")
    assert int(g["last_leaf"][0] - g["first_leaf"][0]) == 4452
    assert int(g["idxP2P"][0]) == 1105228 and int(g["nint_local"][0]) == 63249566
    assert int(g["p2p_count_remote"][0]) == 25022637 and int(g["walk_m2l_count"][0]) == 32
    assert abs(np.sqrt((g["acc"] ** 2).sum(1).mean()) - 7.2293412492e-06) < 1e-15
    box = float(g["box"])
    t = oracle.Tree(demo_pos, 8, [0, 0, 0], [box] * 3)
    assert t.nleaf == 4452 and t.nnode == 4451
    np.testing.assert_array_equal(t.leaves()["npart"], g["leaf_npart"])
    np.testing.assert_array_equal(t.pos[0], g["first_pos"])
    prm = oracle.make_params(box, 32, len(demo_pos), float(g["mass"]))
    ps, pt, ms, mt = t.walk_local(prm)
    assert _hash(ps, pt) == str(g["p2p_hash"]) and _hash(ms, mt) == str(g["m2l_hash"])


@pytest.mark.parametrize("tag", ["t04", "t12"])
def test_small_tree_lists_operators(oracle, small_pos, tag):
    """Tree arrays, lists, M (after P2M+M2M) bit/rounding-exact against the reference dump, NP=1."""
    g = load_golden(f"small_{tag}_np1.npz")
    box = float(g["box"])
    prm = oracle.make_params(box, int(g["nside"]), len(small_pos), float(g["mass"]), theta=float(g["theta"]))
    t = oracle.Tree(small_pos, 8, [0, 0, 0], [box] * 3)
    np.testing.assert_array_equal(t.ids, g["r0_ids"])
    lf, nd = t.leaves(), t.nodes()
    for f in ("npart", "ipart", "width", "center"):
        np.testing.assert_array_equal(lf[f], g[f"r0_leaf_{f}"])
    for f in ("npart", "son", "split", "width", "center"):
        np.testing.assert_array_equal(nd[f], g[f"r0_node_{f}"])
    t.upward(float(g["mass"]))
    scale = np.abs(g["r0_node_M"]).max(0) + 1e-300
    assert np.abs((t.leaves()["M"] - g["r0_leaf_M"]) / scale).max() < 1e-13
    assert np.abs((t.nodes()["M"] - g["r0_node_M"]) / scale).max() < 1e-13
    ps, pt, ms, mt = t.walk_local(prm)
    for a, k in ((ps, "p2p_s"), (pt, "p2p_t"), (ms, "m2l_s"), (mt, "m2l_t")):
        np.testing.assert_array_equal(a, g[f"r0_{k}"])
    if tag == "t12":
        assert len(ms) > 0


@pytest.mark.parametrize("tag", ["t04", "t12"])
def test_small_let_and_remote_lists(oracle, small_pos, tag):
    """prepare_sendtree2 / remote walks against every tree rank 0 received in the reference NP=2 run."""
    g = load_golden(f"small_{tag}_np2.npz")
    let = load_golden(f"small_{tag}_let_np2.npz")
    box = float(g["box"])
    prm = oracle.make_params(box, int(g["nside"]), len(small_pos), float(g["mass"]), theta=float(g["theta"]))
    dc, dw, dstart, splits = oracle.domain_boxes(2, box)
    trees = []
    for r in range(2):
        # the reference's post-build particle order is a fixed point of the in-place partition
        # (src/fmm.c:60-72), so feeding it back reproduces the rank's tree without replaying
        # domain_decomposition's permutation
        p = small_pos[g[f"r{r}_ids"]]
        bl, br = dc[r] - 0.5 * dw[r], dc[r] + 0.5 * dw[r]
        t = oracle.Tree(p, 8, bl, br, direct0=int(dstart[r]))
        np.testing.assert_array_equal(t.ids, np.arange(len(p)))
        for f in ("npart", "ipart", "width", "center"):
            np.testing.assert_array_equal(t.leaves()[f], g[f"r{r}_leaf_{f}"])
        t.upward(float(g["mass"]))
        trees.append(t)
    # rank 0 receives, in call order: shift 0 from rank 1; then 26 shifts x (idx 0: itself, idx 1: rank 1)
    shifts = [(0.0, 0.0, 0.0)] + [((q // 9 - 1) * box, ((q // 3) % 3 - 1) * box, (q % 3 - 1) * box)
                                  for q in range(27) if q != 13]
    calls = [(1, shifts[0])] + [(s, sh) for sh in shifts[1:] for s in (0, 1)]
    assert int(let["ncap"]) == len(calls)
    l0 = dc[0] - 0.5 * dw[0]
    r0 = dc[0] + 0.5 * dw[0]
    tc, tw = 0.5 * (r0 + l0), r0 - l0
    nm2l = 0
    for i, (sender, sh) in enumerate(calls):
        lt = trees[sender].let_pack(prm, tc, tw, sh)
        a = lt.arrays()
        np.testing.assert_array_equal(a["npart"], let[f"c{i}_tree_npart"])
        np.testing.assert_array_equal(a["son"], let[f"c{i}_tree_son"])
        np.testing.assert_array_equal(a["width"], let[f"c{i}_tree_width"])
        np.testing.assert_array_equal(a["center"], let[f"c{i}_tree_center"])
        np.testing.assert_array_equal(a["body"], let[f"c{i}_body"])
        ps, pt, ms, mt = trees[0].walk_remote(lt, prm)
        for x, k in ((ps, "p2p_s"), (pt, "p2p_t"), (ms, "m2l_s"), (mt, "m2l_t")):
            np.testing.assert_array_equal(x, let[f"c{i}_{k}"])
        nm2l += len(ms)
    if tag == "t12":
        assert nm2l > 0


@pytest.mark.parametrize("tag,nranks", [("t04", 1), ("t12", 1), ("t04", 2), ("t12", 2)])
def test_small_force_and_L(oracle, small_pos, tag, nranks):
    g = load_golden(f"small_{tag}_np{nranks}.npz")
    prm = oracle.make_params(float(g["box"]), int(g["nside"]), len(small_pos), float(g["mass"]), theta=float(g["theta"]))
    acc, cnt = oracle.force(small_pos, prm, nranks)
    assert cnt["p2p_pairs"] == int(g["idxP2P"].sum()) and cnt["m2l_pairs"] == int(g["idxM2L"].sum())
    assert cnt["int_remote"] == int(g["p2p_count_remote"].sum()) and cnt["m2l_calls"] == int(g["walk_m2l_count"].sum())
    assert rms_rel(acc, g["acc"]) < 1e-12


def test_operator_identities(oracle):
    """Self-consistency of the operator restatements: M2M then M2L+L2L+L2P equals direct P2M->M2L->L2P
    within truncation-free identities (translation operators compose exactly for polynomials)."""
    import ctypes as C
    L = oracle.lib()
    rng = np.random.default_rng(3)
    pos = rng.random((8, 3))
    dp = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))
    c1 = np.array([0.5, 0.5, 0.5])
    c2 = np.array([0.7, 0.2, 0.4])
    M1, M2, M2b = np.zeros(20), np.zeros(20), np.zeros(20)
    L.pno_p2m(dp(pos), 0, 8, dp(c1), 1.5, dp(M1))
    L.pno_p2m(dp(pos), 0, 8, dp(c2), 1.5, dp(M2))
    d = c2 - c1
    L.pno_m2m(d[0], d[1], d[2], dp(M1), dp(M2b))
    np.testing.assert_allclose(M2b, M2, rtol=1e-12, atol=1e-14)   # M2M is exact for order <= 3
    # L2L exactness: shifting a cubic Taylor expansion is exact
    Lc = rng.random(20)
    La, Lb = np.zeros(20), np.zeros(20)
    s1, s2 = np.array([0.1, -0.2, 0.05]), np.array([-0.3, 0.1, 0.2])
    L.pno_l2l(s1[0], s1[1], s1[2], dp(Lc), dp(La))
    L.pno_l2l(s2[0], s2[1], s2[2], dp(La.copy()), dp(Lb))
    Ld = np.zeros(20)
    s = s1 + s2
    L.pno_l2l(s[0], s[1], s[2], dp(Lc), dp(Ld))
    np.testing.assert_allclose(Lb, Ld, rtol=1e-12, atol=1e-14)


@pytest.mark.parametrize("nranks", [1, 2])
def test_merger_open_newtonian(oracle, nranks):
    """configs[1]: demo/ic_merger.gdt2 through the reference built without -DPERIODIC_CONDITION -DLONGSHORT
    (plain 1/r^2 P2P, G = 1/r M2L, no images): pins the oracle's non-periodic / Newtonian branches."""
    import os
    from conftest import GOLDEN
    g = load_golden(f"merger_open_np{nranks}.npz")
    pos = np.load(os.path.join(GOLDEN, "merger_pos_f32.npy")).astype(np.float64) + float(g["shift"])
    prm = oracle.make_params(float(g["box"]), int(g["nside"]), len(pos), float(g["mass"]), maxleaf=8, theta=0.4, periodic=0, longshort=0)
    assert prm.soft == float(g["soft"])
    acc, cnt = oracle.force(pos, prm, nranks)
    assert cnt["p2p_pairs"] == int(g["idxP2P"].sum()) and cnt["m2l_pairs"] == int(g["idxM2L"].sum())
    assert cnt["int_local"] == int(g["nint_local"].sum()) and cnt["int_remote"] == int(g["p2p_count_remote"].sum())
    assert cnt["m2l_calls"] == int(g["walk_m2l_count"].sum())
    assert rms_rel(acc, g["acc"]) < 1e-12


def test_device_tree_restatement_has_the_reference_leaf_sets(oracle, demo_pos):
    """The Mode B tree (integer-mean, level-synchronous: oracle.TreeB restates the DEVICE builder) partitions the
    demo IC and the merger IC into the same leaf particle sets as the reference's build_kdtree."""
    import os
    from conftest import GOLDEN
    merger = np.load(os.path.join(GOLDEN, "merger_pos_f32.npy")).astype(np.float64) + 200.0
    for pos, box in ((demo_pos, 100000.0), (merger, 400.0)):
        ta = oracle.Tree(pos, 8, [0, 0, 0], [box] * 3)
        tb = oracle.TreeB(pos, 8, [0, 0, 0], [box] * 3)
        la, lb = ta.leaves(), tb.leaves()
        sa = set(tuple(sorted(ta.ids[i:i + n])) for i, n in zip(la["ipart"], la["npart"]))
        sb = set(tuple(sorted(tb.ids[i:i + n])) for i, n in zip(lb["ipart"], lb["npart"]))
        assert sa == sb
        # every particle lies inside its leaf's box
        c, w = lb["center"], lb["width"]
        k = np.repeat(np.arange(tb.nleaf), lb["npart"])
        order = np.concatenate([np.arange(i, i + n) for i, n in zip(lb["ipart"], lb["npart"])])
        assert (np.abs(tb.pos[order] - c[k]) <= 0.5 * w[k] + 1e-9).all()
