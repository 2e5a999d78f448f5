"""Checker helpers for Mode B (device-built tree and lists): the oracle evaluated on the device's own tree.
TEST INFRASTRUCTURE (uses oracle/); shared by tests/ and __graft_entry__.smoke()."""
import numpy as np

IMG = 27


def image_shifts(box):
    return [((q // 9 - 1) * box, ((q // 3) % 3 - 1) * box, (q % 3 - 1) * box) for q in range(27) if q != 13]


def oracle_step_on_tree(oracle, t, prm, tc, tw, want_lists=False):
    """NP=1 force evaluation with the oracle's (reference-pinned) walkers and operators on tree t
    (oracle.Tree or oracle.TreeB): local walk + 26 displaced pruned self-images (src/fmm.c:1028-1045).
    Returns acc in tree order and, optionally, the pair sets keyed like the device lists."""
    acc = np.zeros((t.n, 3))
    t.upward(prm.mass)
    ps, pt, ms, mt = t.walk_local(prm)
    t.eval_p2p(prm, ps, pt, acc)
    t.eval_m2l(prm, ms, mt)
    n0 = t.first_leaf          # cell id = oracle id - n
    p2p = [np.stack([pt - n0, ps - n0], 1).astype(np.int64)] if want_lists else None
    m2l = [np.stack([mt - n0, ms - n0], 1).astype(np.int64)] if want_lists else None
    nint = int((t.leaves()["npart"][ps - n0].astype(np.int64) * t.leaves()["npart"][pt - n0]).sum()
               - t.leaves()["npart"][pt[ps == pt] - n0].sum())
    if prm.periodic:
        lf_np = t.leaves()["npart"]
        for k, sh in enumerate(image_shifts(prm.box)):
            lt = t.let_pack(prm, tc, tw, sh)
            rps, rpt, rms, rmt = t.walk_remote(lt, prm)
            t.eval_p2p_remote(lt, prm, rps, rpt, acc)
            t.eval_m2l_remote(lt, prm, rms, rmt)
            a = lt.arrays()
            nint += int((a["npart"][rps].astype(np.int64) * lf_np[rpt - n0]).sum())
            if want_lists:
                org = a["origin"].astype(np.int64) - n0
                p2p.append(np.stack([rpt - n0, org[rps] | ((k + 1) << IMG)], 1).astype(np.int64))
                m2l.append(np.stack([rmt - n0, org[rms] | ((k + 1) << IMG)], 1).astype(np.int64))
    t.downward(acc)
    out = {"acc": acc, "nint": nint}
    if want_lists:
        out["p2p"] = np.concatenate(p2p)
        out["m2l"] = np.concatenate(m2l)
    return out


def csr_to_pairs(sink, off, src):
    cnt = np.diff(off)
    s = np.repeat(sink.astype(np.int64), cnt)
    return np.stack([s, src.astype(np.int64)], 1)


def sort_pairs(p):
    if len(p) == 0:
        return p.reshape(0, 2)
    k = np.lexsort((p[:, 1], p[:, 0]))
    return p[k]


def rms_rel(a, b):
    return float(np.sqrt(((a - b) ** 2).sum() / (b ** 2).sum()))


def oracle_step_multirank(oracle, pos_by_rank, boxes, direct0s, prm):
    """Multi-rank force evaluation with the oracle's reference-pinned walkers / operators on Mode-B trees:
    per rank local walk; then for every displacement (27 if periodic) and every sender the tree of the sender
    pruned against the receiver's box (src/fmm.c:1021-1045, src/remotes.c:684-751).  boxes[r] = (lo, hi).
    Returns per-rank acc (input order of pos_by_rank[r]) and interaction counts."""
    P = len(pos_by_rank)
    trees, accs, nints = [], [], []
    for r in range(P):
        t = oracle.TreeB(pos_by_rank[r], prm.maxleaf, boxes[r][0], boxes[r][1], direct0=direct0s[r])
        t.upward(prm.mass)
        trees.append(t)
    for r in range(P):
        t = trees[r]
        acc = np.zeros((t.n, 3))
        n0 = t.first_leaf
        lf_np = t.leaves()["npart"]
        ps, pt, ms, mt = t.walk_local(prm)
        t.eval_p2p(prm, ps, pt, acc)
        t.eval_m2l(prm, ms, mt)
        nint = int((lf_np[ps - n0].astype(np.int64) * lf_np[pt - n0]).sum() - lf_np[pt[ps == pt] - n0].sum())
        lo, hi = np.asarray(boxes[r][0], float), np.asarray(boxes[r][1], float)
        tc, tw = 0.5 * (hi + lo), hi - lo
        shifts = [(0.0, 0.0, 0.0)] + (image_shifts(prm.box) if prm.periodic else [])
        for k, sh in enumerate(shifts):
            for s in range(P):
                if k == 0 and s == r:
                    continue
                lt = trees[s].let_pack(prm, tc, tw, sh)
                rps, rpt, rms, rmt = t.walk_remote(lt, prm)
                t.eval_p2p_remote(lt, prm, rps, rpt, acc)
                t.eval_m2l_remote(lt, prm, rms, rmt)
                nint += int((lt.arrays()["npart"][rps].astype(np.int64) * lf_np[rpt - n0]).sum())
        t.downward(acc)
        out = np.zeros_like(acc)
        out[t.ids] = acc
        accs.append(out)
        nints.append(nint)
    return accs, nints
