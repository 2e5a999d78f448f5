"""Domain decomposition on the device (src/domains.c:163-375) through the C-ABI:
pn2_domain_owner_device against the UNMODIFIED reference's prepare_body_inOrderOf_domain (golden owners under the
initial and the load-adjusted splits, bit-exact), and pn2_migrate_begin / _exchange_local / _result with all ranks
as contexts of one process: every rank ends up with exactly the records the reference would deliver to it, intact,
in blocks by source rank.  (The NCCL transport of the same blocks is covered by tests/test_nccl_two_gpus.py.)"""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def golden():
    g = np.load(os.path.join(ROOT, "tests", "golden", "domains_golden.npz"))
    pos = np.load(os.path.join(ROOT, "tests", "golden", "demo_pos_f32.npy")).astype(np.float64)[::int(g["pos_stride"])]
    return g, pos


def ctx_of(pn2):
    return pn2.Context(pn2.make_params(100000.0, 32, 32768, 1.0))


def test_owner_matches_reference(pn2, golden):
    import torch
    g, pos = golden
    ctx = ctx_of(pn2)
    d_pos = torch.from_numpy(pos).cuda()
    d_own = torch.empty(len(pos), dtype=torch.int32, device="cuda")
    for ci in range(int(g["ncase"])):
        P = int(g[f"P{ci}"])
        for k in range(len(g[f"loads{ci}"])):
            ctx.domain_owner_device(d_pos.data_ptr(), 3, len(pos), g[f"splits{ci}"][k], P, d_own.data_ptr())
            np.testing.assert_array_equal(d_own.cpu().numpy(), g[f"owner{ci}"][k])
    # records with a stride (the reference's 96-byte Body)
    body = torch.zeros((len(pos), 12), dtype=torch.float64, device="cuda")
    body[:, :3] = d_pos
    ctx.domain_owner_device(body.data_ptr(), 12, len(pos), g["splits4"][0], 8, d_own.data_ptr())
    np.testing.assert_array_equal(d_own.cpu().numpy(), g["owner4"][0])


@pytest.mark.parametrize("case,step", [(0, 0), (1, 2), (2, 1), (3, 0), (4, 0), (4, 2)])
def test_migrate_local_ranks(pn2, golden, case, step):
    import torch
    import domains
    g, pos = golden
    P = int(g[f"P{case}"])
    splits, owner = g[f"splits{case}"][step], g[f"owner{case}"][step]
    n = len(pos)
    rec = np.zeros((n, 12))
    rec[:, :3] = pos
    rec[:, 3:] = np.random.default_rng(7).standard_normal((n, 9))
    rec[:, 6] = np.arange(n)                                   # tag
    doms = domains.boxes_from_splits(P, 100000.0, splits)
    ctxs, held = [], []
    for r in range(P):
        c = ctx_of(pn2)
        c.set_comm(r, P, domains._domains_from(P, *doms), None)
        ctxs.append(c)
        mine = rec[r::P].copy()                                # any initial distribution
        held.append(torch.from_numpy(mine).cuda())
    sc = [ctxs[r].migrate_begin(held[r].data_ptr(), 12, held[r].shape[0], splits) for r in range(P)]
    for r in range(P):
        np.testing.assert_array_equal(sc[r], np.bincount(owner[r::P], minlength=P))
    pn2.migrate_exchange_local(ctxs)
    total = 0
    for r in range(P):
        ptr, m, rc = ctxs[r].migrate_result()
        np.testing.assert_array_equal(rc, [sc[s][r] for s in range(P)])
        assert m == int((owner == r).sum())
        total += m
        o = ctxs[r].migrate_fetch(12)
        assert o.shape[0] == m
        tags = o[:, 6].astype(np.int64)
        np.testing.assert_array_equal(np.sort(tags), np.flatnonzero(owner == r))      # exactly the reference's set
        np.testing.assert_array_equal(o, rec[tags])                                     # records intact
        # blocks by source rank, inside a block the sender's order
        np.testing.assert_array_equal(tags % P, np.repeat(np.arange(P), rc))
        for s in range(P):
            blk = tags[rc[:s].sum():rc[:s + 1].sum()]
            assert np.all(np.diff(blk) > 0)
    assert total == n


def test_single_rank_and_empty(pn2, golden):
    import torch
    g, pos = golden
    ctx = ctx_of(pn2)
    d = torch.from_numpy(pos).cuda()
    ptr, m = ctx.migrate_device(d.data_ptr(), 3, len(pos), np.zeros(1))
    assert m == len(pos) and ptr
    ptr, m = ctx.migrate_device(0, 3, 0, np.zeros(1))
    assert m == 0
    with pytest.raises(pn2.Pn2Error):
        ctx.migrate_begin(d.data_ptr(), 2, len(pos), np.zeros(1))      # a record holds at least the position
