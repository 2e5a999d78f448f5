"""Shared by __graft_entry__.smoke() and tests: one small invocation of the hot path vs the oracle."""
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def rms_rel(a, b):
    return float(np.sqrt(((a - b) ** 2).sum() / (b ** 2).sum()))


def mode_a_small(pn2, oracle, precision):
    g = np.load(os.path.join(GOLDEN, "small_t12_np1.npz"))
    pos = np.load(os.path.join(GOLDEN, "demo_pos_f32.npy")).astype(np.float64)[::8].copy()
    box = float(g["box"])
    prm_o = oracle.make_params(box, int(g["nside"]), len(pos), float(g["mass"]), theta=float(g["theta"]))
    t = oracle.Tree(pos, 8, [0, 0, 0], [box] * 3)
    t.upward(prm_o.mass)
    lf, nd = t.leaves(), t.nodes()
    leaf = np.zeros(t.nleaf, pn2.PACK)
    btree = np.zeros(t.nnode, pn2.NODE)
    for f in ("npart", "ipart", "width", "center"):
        leaf[f] = lf[f]
    for f in ("npart", "son", "split", "width", "center"):
        btree[f] = nd[f]
    ctx = pn2.Context(pn2.Params(box, prm_o.rs, prm_o.cutoff, prm_o.soft, prm_o.theta, prm_o.mass, 8, 1, 1, precision))
    ps, pt, ms, mt = t.walk_local(prm_o)
    remotes = []
    tc, tw = np.array([0.5 * box] * 3), np.array([box] * 3)
    for q in range(27):
        if q == 13:
            continue
        sh = ((q // 9 - 1) * box, ((q // 3) % 3 - 1) * box, (q % 3 - 1) * box)
        lt = t.let_pack(prm_o, tc, tw, sh)
        a = lt.arrays()
        rt = np.zeros(lt.nnode, pn2.RNODE)
        for f in ("npart", "son", "width", "center", "M"):
            rt[f] = a[f]
        rb = np.zeros(lt.nbody, pn2.RBODY)
        rb["pos"] = a["body"]
        remotes.append((rt, rb) + (lambda w: ((w[0], w[1]), (w[2], w[3])))(t.walk_remote(lt, prm_o)))
    acc = pn2.short_range_force_mode_a(ctx, t.pos, leaf, t.first_leaf, btree, t.first_node, (ps, pt), (ms, mt), remotes)
    out = np.zeros_like(acc)
    out[t.ids] = acc
    launches = ctx.launch_count()
    ctx.close()
    return rms_rel(out, g["acc"]), launches


def mode_b_small(pn2, oracle, precision):
    """Device-built tree + lists + images + operators on N = 4096, against the oracle evaluated on the device's own
    tree (bit-exact tree check included) and against the reference's golden accelerations."""
    from modeb_check import oracle_step_on_tree
    g = np.load(os.path.join(GOLDEN, "small_t12_np1.npz"))
    pos = np.load(os.path.join(GOLDEN, "demo_pos_f32.npy")).astype(np.float64)[::8].copy()
    box = float(g["box"])
    prm_o = oracle.make_params(box, int(g["nside"]), len(pos), float(g["mass"]), theta=float(g["theta"]))
    ctx = pn2.Context(pn2.Params(box, prm_o.rs, prm_o.cutoff, prm_o.soft, prm_o.theta, prm_o.mass, 8, 1, 1, precision))
    acc = ctx.force_step(pos)
    tb = oracle.TreeB(pos, 8, [0, 0, 0], [box] * 3)
    assert np.array_equal(ctx.get_order(), tb.ids), "device tree order differs from the oracle restatement"
    ref = oracle_step_on_tree(oracle, tb, prm_o, np.array([0.5 * box] * 3), np.array([box] * 3))
    ref_acc = np.zeros_like(acc)
    ref_acc[tb.ids] = ref["acc"]
    info = ctx.step_info()
    assert info["n_interactions"] == ref["nint"], "device lists differ from the oracle's"
    launches = ctx.launch_count()
    ctx.close()
    return rms_rel(acc, ref_acc), rms_rel(acc, g["acc"]), launches


def run(pn2, oracle, np_, verbose=False):
    for precision, tol in ((pn2.FP64, 1e-6), (pn2.FP32, 1e-4)):
        e1, e2, launches = mode_b_small(pn2, oracle, precision)
        if verbose:
            print(f"smoke: Mode B N=4096 precision={'FP64' if precision == 0 else 'FP32'} rms rel err vs oracle {e1:.3e}, "
                  f"vs reference golden {e2:.3e} ({launches} kernel launches)")
        assert e1 < tol and e2 < tol, (precision, e1, e2)
    for precision, tol in ((pn2.FP64, 1e-6), (pn2.FP32, 1e-4)):
        err, launches = mode_a_small(pn2, oracle, precision)
        if verbose:
            print(f"smoke: Mode A N=4096 precision={'FP64' if precision == 0 else 'FP32'} rms rel err vs reference = {err:.3e}"
                  f" ({launches} kernel launches)")
        assert err < tol, (precision, err)
