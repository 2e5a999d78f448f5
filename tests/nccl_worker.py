"""Worker of tests/test_nccl_two_gpus.py (one process per GPU under torch.distributed.run): a Mode B force step of
the reference's demo IC over NCCL, compared on rank 0 with the UNMODIFIED reference's golden accelerations at the
same number of ranks."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "photons-2.0_b200"))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402
import domains  # noqa: E402
import pn2gpu  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    g = np.load(os.path.join(ROOT, "tests", "golden", f"demo_ns32_np{world}.npz"))
    pos = np.load(os.path.join(ROOT, "tests", "golden", "demo_pos_f32.npy")).astype(np.float64)
    box = float(g["box"])
    doms = domains.domain_boxes(world, box)
    owner = domains.domain_of(pos, world, box)
    idx = np.nonzero(owner == rank)[0]
    worst = 0.0
    for precision, tol in ((pn2gpu.FP64, 1e-9), (pn2gpu.FP32, 3e-5)):
        prm = pn2gpu.make_params(box, 32, len(pos), float(g["mass"]), maxleaf=8, theta=0.4, precision=precision)
        ctx = pn2gpu.Context(prm, device=local)
        ctx.set_comm_torch(rank, world, doms)
        dpos = torch.from_numpy(pos[idx]).cuda()
        dacc = torch.zeros_like(dpos)
        for _ in range(2):                                  # twice: buffers are reused across steps
            ctx.force_step_device(dpos.data_ptr(), len(idx), dacc.data_ptr(), doms[rank])
            ctx.sync()
        ref = torch.from_numpy(g["acc"][idx]).cuda()
        num = ((dacc - ref) ** 2).sum()
        den = (ref ** 2).sum()
        t = torch.stack([num, den])
        dist.all_reduce(t)
        err = float(torch.sqrt(t[0] / t[1]))
        nint = torch.tensor([ctx.step_info()["n_interactions"]], device="cuda")
        dist.all_reduce(nint)
        if rank == 0:
            print(f"NCCL NP={world} precision {precision}: rms rel err vs reference golden {err:.3e}, interactions {int(nint)}", flush=True)
            assert int(nint) == int(g["nint_local"].sum() + g["p2p_count_remote"].sum())
            assert err < tol, err
        worst = max(worst, err)
        ctx.close()
    # ---- domain decomposition over NCCL (pn2_migrate_*, src/domains.c:268-375) against the reference's golden owners ----
    gd = np.load(os.path.join(ROOT, "tests", "golden", "domains_golden.npz"))
    case = {2: 0, 4: 2, 8: 4}.get(world)
    if case is not None:
        sub = pos[::int(gd["pos_stride"])]
        n = len(sub)
        splits, own = gd[f"splits{case}"][0], gd[f"owner{case}"][0]
        rec = np.zeros((n, 12))
        rec[:, :3] = sub
        rec[:, 6] = np.arange(n)
        prm = pn2gpu.make_params(box, 32, len(pos), float(g["mass"]))
        ctx = pn2gpu.Context(prm, device=local)
        ctx.set_comm_torch(rank, world, doms)
        held = torch.from_numpy(rec[rank::world].copy()).cuda()
        for _ in range(2):
            ptr, m = ctx.migrate_device(held.data_ptr(), 12, held.shape[0], splits)
            out = ctx.migrate_fetch(12)
        tags = out[:, 6].astype(np.int64)
        assert m == int((own == rank).sum()), (m, int((own == rank).sum()))
        assert np.array_equal(np.sort(tags), np.flatnonzero(own == rank))
        assert np.array_equal(out, rec[tags])
        cnt = torch.tensor([m], device="cuda")
        dist.all_reduce(cnt)
        assert int(cnt) == n
        if rank == 0:
            print(f"NCCL NP={world} migration: {n} records of 96 bytes, every rank holds the reference's set", flush=True)
        ctx.close()
    # ---- PM long-range force with the mesh all-reduced over NCCL (pn2_pm_force_device) against the reference's one-rank result ----
    gp = np.load(os.path.join(ROOT, "tests", "golden", "pm_demo_ns32.npz"))
    prm = pn2gpu.make_params(box, 32, len(pos), float(gp["mass"]), precision=pn2gpu.FP64)
    ctx = pn2gpu.Context(prm, device=local)
    ctx.set_comm_torch(rank, world, doms)
    dpos = torch.from_numpy(pos[idx]).cuda()
    dacc = torch.zeros_like(dpos)
    for _ in range(2):
        ctx.pm_force_device(dpos.data_ptr(), len(idx), 32, dacc.data_ptr())
        ctx.sync()
    ref = torch.from_numpy(gp["acc_pm"][idx]).cuda()
    t = torch.stack([((dacc - ref) ** 2).sum(), (ref ** 2).sum()])
    dist.all_reduce(t)
    err = float(torch.sqrt(t[0] / t[1]))
    if rank == 0:
        print(f"NCCL NP={world} PM force (ncclAllReduce of the mesh): rms rel err vs the reference {err:.3e}", flush=True)
        assert err < 1e-10, err
    ctx.close()
    dist.destroy_process_group()
    if rank == 0:
        print("NCCL_WORKER_OK", flush=True)


if __name__ == "__main__":
    main()
