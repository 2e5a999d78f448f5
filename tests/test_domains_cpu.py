"""CPU tests of the host-side multi-rank logic: the domain geometry of photons-2.0_b200/domains.py against the oracle
(src/domains.c:399-472, src/toptree.c:150-181, src/initial.c:199-223), and a world_size-2 gloo run of the
partitioning + id-broadcast plumbing the multi-GPU bench uses."""
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("P", [1, 2, 3, 4, 5, 8])
def test_domain_boxes_and_owner_match_oracle(pn2, oracle, demo_pos, P):
    import domains
    box = 100000.0
    dc, dw, dstart, splits = oracle.domain_boxes(P, box)
    doms = domains.domain_boxes(P, box)
    for r in range(P):
        np.testing.assert_array_equal(np.array(doms[r].lo), dc[r] - 0.5 * dw[r])
        np.testing.assert_array_equal(np.array(doms[r].hi), dc[r] + 0.5 * dw[r])
        assert doms[r].direct0 == dstart[r]
    import ctypes as C
    L = oracle.lib()
    pos = demo_pos[::16]
    own = domains.domain_of(pos, P, box)
    ref = np.array([L.pno_domain_of(pos[i].ctypes.data_as(C.POINTER(C.c_double)), P, splits.ctypes.data_as(C.POINTER(C.c_double)))
                    for i in range(len(pos))])
    np.testing.assert_array_equal(own, ref)
    import torch
    own_t = domains.domain_of(torch.from_numpy(pos), P, box).numpy()
    np.testing.assert_array_equal(own_t, ref)


def test_gloo_two_ranks_partition():
    """world_size 2 over gloo: both ranks generate the same synthetic set, keep their own domain, and agree on counts."""
    code = r'''
import os, sys
sys.path.insert(0, os.path.join(ROOT, "photons-2.0_b200"))
import torch, torch.distributed as dist
import numpy as np
import domains, synthetic
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
pos = synthetic.lcdm_like(16, device="cpu")
own = domains.domain_of(pos, world, synthetic.BOX)
mine = pos[own == rank]
doms = domains.domain_boxes(world, synthetic.BOX)
lo, hi = np.array(doms[rank].lo), np.array(doms[rank].hi)
assert bool(((mine.numpy() >= lo) & (mine.numpy() <= hi)).all())
cnt = torch.tensor([mine.shape[0]])
dist.all_reduce(cnt)
assert int(cnt) == 16 ** 3, int(cnt)
# the 128-byte id broadcast used for the NCCL communicator
t = torch.arange(128, dtype=torch.uint8) if rank == 0 else torch.zeros(128, dtype=torch.uint8)
dist.broadcast(t, 0)
assert int(t.sum()) == sum(range(128))
dist.destroy_process_group()
print("ok", rank)
'''.replace("ROOT", repr(ROOT))
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29613")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
                        "--master-port", "29613", "-c", code] if False else
                       [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
                        "--master-port", "29613", "--no-python", sys.executable, "-c", code],
                       env=env, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert r.stdout.count("ok") == 2


def test_adaptive_splits_match_reference_golden():
    """DomainTree.update (measure_domain_runtime + determine_split_domtree, src/domains.c:21-38, 86-160) and the owner
    rule under the adjusted splits, against the unmodified reference (tests/golden/make_domain_golden.py)."""
    import domains
    g = np.load(os.path.join(ROOT, "tests", "golden", "domains_golden.npz"))
    pos = np.load(os.path.join(ROOT, "tests", "golden", "demo_pos_f32.npy")).astype(np.float64)[::int(g["pos_stride"])]
    for ci in range(int(g["ncase"])):
        P = int(g[f"P{ci}"])
        dt = domains.DomainTree(P, float(g["box"]))
        np.testing.assert_array_equal(dt.splits[:P - 1], g[f"split0_{ci}"][:P - 1])
        for k, load in enumerate(g[f"loads{ci}"]):
            dt.update(load)
            np.testing.assert_array_equal(dt.splits[:P - 1], g[f"splits{ci}"][k][:P - 1])       # bit-exact doubles
            own = dt.owner(pos)
            np.testing.assert_array_equal(own, g[f"owner{ci}"][k])
            np.testing.assert_array_equal(np.bincount(own, minlength=P), g[f"sendcount{ci}"][k])
            # every rank's particles lie inside its box
            doms = dt.boxes()
            for r in range(P):
                q = pos[own == r]
                assert np.all(q >= np.array(doms[r].lo)) and np.all(q <= np.array(doms[r].hi))


def test_gloo_two_ranks_adaptive_decomposition():
    """world_size 2 over gloo: the per-step decomposition loop of src/photoNs.c:277-283 + src/domains.c:384-394 --
    every rank reports its load (here: its particle count, standing in for idxP2P + idxM2L), the loads are gathered
    (MPI_Allgather in measure_domain_runtime), every rank updates the same DomainTree, particles change owner; the
    ranks must agree on the splits bit for bit, the set stays partitioned, and the imbalance shrinks."""
    code = r'''
import os, sys
sys.path.insert(0, os.path.join(ROOT, "photons-2.0_b200"))
import torch, torch.distributed as dist
import numpy as np
import domains
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
pos = np.load(os.path.join(ROOT, "tests", "golden", "demo_pos_f32.npy")).astype(np.float64)      # clustered demo IC
dt = domains.DomainTree(world, 100000.0)
imb = []
for step in range(6):
    own = dt.owner(pos)
    mine = int((own == rank).sum())
    cnt = torch.tensor([float(mine)], dtype=torch.float64)
    allc = [torch.zeros(1, dtype=torch.float64) for _ in range(world)]
    dist.all_gather(allc, cnt)
    counts = np.array([float(c) for c in allc])
    assert counts.sum() == len(pos)
    imb.append(counts.max() / counts.mean())
    load = counts * world / (counts.sum() + 0.0001)               # DTIME_FRACTION, src/photoNs.c:283
    dt.update(load)
    s = torch.from_numpy(dt.splits.copy())
    ref = s.clone()
    dist.broadcast(ref, 0)
    assert torch.equal(s, ref)                                     # same doubles on every rank
    lo, hi = np.array(dt.boxes()[rank].lo), np.array(dt.boxes()[rank].hi)
    q = pos[dt.owner(pos) == rank]
    assert bool(((q >= lo) & (q <= hi)).all())
assert imb[-1] < imb[0] or imb[0] < 1.001, imb
dist.destroy_process_group()
print("ok", rank, ["%.4f" % x for x in imb])
'''.replace("ROOT", repr(ROOT))
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29614")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
                        "--master-port", "29614", "--no-python", sys.executable, "-c", code],
                       env=env, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert r.stdout.count("ok") == 2
    print(r.stdout[-300:])
