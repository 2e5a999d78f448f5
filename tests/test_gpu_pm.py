"""The particle-mesh long-range force on the device (pn2_pm_*: csrc/pn2_pm.cu) against the reference's
PM path: src/partmesh.c compiled unmodified around the restated convolution of src/conv.f90 (tests/golden/pm_*.npz,
tests/golden/make_pm_golden.py).  FP64 on both sides; the device sums its CIC deposits with atomics, so agreement is
to rounding (~1e-13), not bit for bit."""
import math
import os

import numpy as np
import pytest
from scipy.special import erfc

from conftest import load_golden, rms_rel

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def ctx_for(pn2, g, n):
    prm = pn2.make_params(float(g["box"]), int(g["nside"]), n, float(g["mass"]), precision=pn2.FP64)
    return pn2.Context(prm)


@pytest.mark.parametrize("name,stride", [("pm_demo_ns32.npz", 1), ("pm_small_ns24.npz", 8)])
def test_pm_force_vs_reference(pn2, demo_pos, name, stride):
    import torch
    g = load_golden(name)
    pos = demo_pos[::stride].copy()
    nside, n = int(g["nside"]), len(pos)
    ctx = ctx_for(pn2, g, n)
    dpos = torch.from_numpy(pos).cuda()
    dacc = torch.zeros_like(dpos)
    # density after the deposit, potential and force after the convolution
    ctx.pm_begin(dpos.data_ptr(), n, nside)
    dens = ctx.pm_mesh(nside)
    ctx.pm_finish(dacc.data_ptr())
    ctx.sync()
    pot = ctx.pm_mesh(nside)
    acc = dacc.cpu().numpy()
    # the deposit kernel leaves the (NSIDE / BOX)^3 renormalisation of src/partmesh.c:168-178 to the Green kernel
    renorm = (nside / float(g["box"])) ** 3
    e_d = np.abs(dens * renorm - g["density"]).max() / np.abs(g["density"]).max()
    e_p = np.abs(pot - g["potential"]).max() / np.abs(g["potential"]).max()
    e_a = rms_rel(acc, g["acc_pm"])
    print(f"{name}: density {e_d:.2e} potential {e_p:.2e} acc_pm rms rel {e_a:.2e}; timings {ctx.pm_timings()}")
    assert e_d < 1e-12 and e_p < 1e-11 and e_a < 1e-10
    # the one-call entry point gives the same
    dacc2 = torch.zeros_like(dpos)
    ctx.pm_force_device(dpos.data_ptr(), n, nside, dacc2.data_ptr())
    ctx.sync()
    assert rms_rel(dacc2.cpu().numpy(), g["acc_pm"]) < 1e-10
    ctx.close()


@pytest.mark.parametrize("nranks", [2, 8])
def test_pm_multirank_mesh_reduction(pn2, demo_pos, nranks):
    """Particles split over the reference's domains, one context per rank, meshes summed by pn2_pm_reduce_local (the
    in-process stand-in of the ncclAllReduce): the force of every particle equals the one-rank reference result."""
    import torch
    import domains
    g = load_golden("pm_demo_ns32.npz")
    nside, box = int(g["nside"]), float(g["box"])
    owner = domains.domain_of(demo_pos, nranks, box)
    idx = [np.nonzero(owner == r)[0] for r in range(nranks)]
    doms = domains.domain_boxes(nranks, box)
    ctxs = [ctx_for(pn2, g, len(demo_pos)) for _ in range(nranks)]
    dpos, dacc = [], []
    for r in range(nranks):
        ctxs[r].set_comm(r, nranks, doms, None)
        t = torch.from_numpy(demo_pos[idx[r]]).cuda()
        dpos.append(t)
        dacc.append(torch.zeros_like(t))
        ctxs[r].pm_begin(t.data_ptr(), t.shape[0], nside)
    pn2.pm_reduce_local(ctxs)
    acc = np.zeros_like(demo_pos)
    for r in range(nranks):
        ctxs[r].pm_finish(dacc[r].data_ptr())
        ctxs[r].sync()
        acc[idx[r]] = dacc[r].cpu().numpy()
    err = rms_rel(acc, g["acc_pm"])
    print(f"PM NP={nranks}: acc_pm rms rel err vs the one-rank reference {err:.2e}")
    assert err < 1e-10
    for c in ctxs:
        c.close()


def test_pm_plus_short_range_is_newtonian(pn2):
    """Both halves of the force split on the device: Mode B short-range step + PM of a particle pair = m / d^2."""
    g = load_golden("pm_pair_ns64.npz")
    pos, box, nside, mass = g["pos"], float(g["box"]), int(g["nside"]), float(g["mass"])
    prm = pn2.make_params(box, nside, 2, mass, soft=0.0, precision=pn2.FP64)
    ctx = pn2.Context(prm)
    a_pm = ctx.pm_force(pos, nside)
    a_sr = ctx.force_step(pos)
    assert rms_rel(a_pm, g["acc_pm"]) < 1e-10
    d = pos[1] - pos[0]
    r = math.sqrt((d ** 2).sum())
    u = r / (2 * prm.rs)
    short = mass / r ** 2 * (erfc(u) + 2 * u / math.sqrt(math.pi) * math.exp(-u * u))
    assert np.abs(a_sr[0] - short * d / r).max() < 1e-8 * short        # table-driven g(u): |err| < 2e-10
    newton = mass / r ** 2 * d / r
    err = math.sqrt((((a_pm + a_sr)[0] - newton) ** 2).sum()) / math.sqrt((newton ** 2).sum())
    print("pair on the device: |PM + short - Newton| / |Newton| =", err)
    assert err < 0.03
    ctx.close()


def test_full_step_on_device_records(pn2, demo_pos):
    """The whole force evaluation of src/photoNs.c:97-116 (PM thread + fmm_*) and the KDK update (:150-196, 254-268) on
    device-resident Body records: pn2_pm_force_records + pn2_force_step_records + pn2_kick_device + pn2_drift_device,
    against the same step driven from host arrays (golden PM force path, numpy restatement of the KDK loops)."""
    import torch
    import cosmology
    import snapshot
    gs = np.load(os.path.join(ROOT, "tests", "golden", "snapshot_golden.npz"))
    g = load_golden("pm_small_ns24.npz")
    pos = demo_pos[::8].copy()
    n, box, nside = len(pos), float(g["box"]), int(g["nside"])
    rng = np.random.default_rng(5)
    vel = rng.standard_normal((n, 3)) * 50.0
    grav = 43007.105732
    prm = pn2.make_params(box, nside, n, float(g["mass"]), precision=pn2.FP64)
    hctx, dctx = pn2.Context(prm), pn2.Context(prm)
    hb = snapshot.to_body(pos, vel)
    hb[:, 3:6] = hctx.force_step(hb[:, 0:3])
    hb[:, 9:12] = hctx.pm_force(np.ascontiguousarray(hb[:, 0:3]), nside)
    assert rms_rel(hb[:, 9:12], g["acc_pm"]) < 1e-10
    db = torch.from_numpy(snapshot.to_body(pos, vel)).cuda()
    dctx.force_step_records(db.data_ptr(), 12, n)
    dctx.pm_force_records(db.data_ptr(), 12, n, nside)
    a_init, dloga = 1.0 / 50.0, 0.02
    for loop in range(2):
        dkh, dd = cosmology.step_factors(loop, dloga, a_init, float(gs["OmegaM0"]), float(gs["OmegaX0"]), grav)
        hb[:, 6:9] += hb[:, 9:12] * dkh
        hb[:, 6:9] += hb[:, 3:6] * dkh
        hb[:, 0:3] += hb[:, 6:9] * dd
        p = hb[:, 0:3]
        while (p < 0.0).any():
            p[p < 0.0] += box
        while (p >= box).any():
            p[p >= box] -= box
        hb[:, 3:6] = hctx.force_step(np.ascontiguousarray(hb[:, 0:3]))
        hb[:, 9:12] = hctx.pm_force(np.ascontiguousarray(hb[:, 0:3]), nside)
        hb[:, 6:9] += hb[:, 3:6] * dkh
        hb[:, 6:9] += hb[:, 9:12] * dkh
        dctx.kick_device(db.data_ptr(), n, dkh, True)
        dctx.drift_device(db.data_ptr(), n, dd, box)
        dctx.force_step_records(db.data_ptr(), 12, n)
        dctx.pm_force_records(db.data_ptr(), 12, n, nside)
        dctx.kick_device(db.data_ptr(), n, dkh, False)
    dctx.sync()
    out = db.cpu().numpy()
    # the PM deposit sums with atomics (order varies run to run): rounding-level agreement, not bit identity
    assert rms_rel(out[:, 3:6], hb[:, 3:6]) < 1e-8
    assert rms_rel(out[:, 9:12], hb[:, 9:12]) < 1e-10
    assert np.abs(out[:, 0:3] - hb[:, 0:3]).max() < 1e-9 * box and rms_rel(out[:, 6:9], hb[:, 6:9]) < 1e-10
    assert np.abs(out[:, 9:12]).max() > 0
    hctx.close(); dctx.close()
