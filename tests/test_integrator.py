"""KDK integrator: the host factors of photons-2.0_b200/cosmology.py bit for bit against the
unmodified reference's kick_loga / drift_loga (CPU), and the device kick / drift kernels bit for bit against a numpy
restatement of the reference's loops (src/photoNs.c:150-196, 254-268) on Body records (GPU)."""
import os

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_step_factors_match_reference():
    import cosmology
    rows = np.load(os.path.join(ROOT, "tests", "golden", "integrator_golden.npz"))["rows"]
    for om, ox, li, lf, kick, drift in rows:
        assert cosmology.kick_loga(li, lf, om, ox) == kick
        assert cosmology.drift_loga(li, lf, om, ox) == drift


def kdk_numpy(body, dkh, dd, box):
    """The reference's opening half kick, drift with periodic wrap, closing half kick; mul and add rounded separately."""
    b = body.copy()
    b[:, 6:9] += b[:, 9:12] * dkh                 # src/photoNs.c:158-162
    b[:, 6:9] += b[:, 3:6] * dkh                  # :165-169
    b[:, 0:3] += b[:, 6:9] * dd                   # :171-175
    p = b[:, 0:3]
    while (p < 0.0).any():                        # :177-195
        p[p < 0.0] += box
    while (p >= box).any():
        p[p >= box] -= box
    b[:, 6:9] += b[:, 3:6] * dkh                  # :257-261
    b[:, 6:9] += b[:, 9:12] * dkh                 # :264-268
    return b


@pytest.mark.gpu
def test_device_kdk_bit_exact(pn2):
    import torch
    rng = np.random.default_rng(11)
    n, box = 100003, 100000.0
    body = np.zeros((n, 12))
    body[:, 0:3] = rng.uniform(0, box, (n, 3))
    body[:, 3:6] = rng.standard_normal((n, 3)) * 1e-5
    body[:, 6:9] = rng.standard_normal((n, 3)) * 3e3          # fast enough that many particles cross the box faces
    body[:, 9:12] = rng.standard_normal((n, 3)) * 1e-5
    body[:5, 6] = [4e6, -4e6, 2.5e6, -1e-300, 0.0]            # several box lengths per step, denormal products
    dkh, dd = 0.5 * 0.1755554 * 43007.105732, 8.51355621
    ctx = pn2.Context(pn2.make_params(box, 32, n, 1.0))
    d = torch.from_numpy(body).cuda()
    ctx.kick_device(d.data_ptr(), n, dkh, True)
    ctx.drift_device(d.data_ptr(), n, dd, box)
    ctx.kick_device(d.data_ptr(), n, dkh, False)
    ctx.sync()
    out = d.cpu().numpy()
    ref = kdk_numpy(body, dkh, dd, box)
    np.testing.assert_array_equal(out, ref)
    assert (out[:, :3] >= 0).all() and (out[:, :3] < box).all()
    ctx.kick_device(0, 0, dkh, True)                           # empty set
    with pytest.raises(pn2.Pn2Error):
        ctx.drift_device(d.data_ptr(), n, dd, 0.0)


@pytest.mark.gpu
def test_device_resident_loop_matches_host_driven_loop(pn2):
    """force -> kick -> drift -> kick on device Body records (pn2_force_step_records, pn2_kick_device,
    pn2_drift_device; the particles never leave the GPU) against the same loop driven from host arrays through
    pn2_force_step + the numpy restatement of the reference's KDK loops: bit-identical after 3 steps (PM excluded:
    acc_pm = 0)."""
    import torch
    import cosmology
    import snapshot
    g = np.load(os.path.join(ROOT, "tests", "golden", "snapshot_golden.npz"))
    pos = np.load(os.path.join(ROOT, "tests", "golden", "demo_pos_f32.npy")).astype(np.float64)[::8].copy()
    n, box = len(pos), float(g["BOXSIZE"])
    rng = np.random.default_rng(3)
    vel = rng.standard_normal((n, 3)) * 50.0
    grav = 43007.105732
    prm = pn2.make_params(box, 24, n, float(g["MASSPART"]) * 8, precision=pn2.FP64)
    # host-driven reference loop
    hb = snapshot.to_body(pos, vel)
    hctx = pn2.Context(prm)
    hb[:, 3:6] = hctx.force_step(hb[:, 0:3])
    # device-resident loop
    dctx = pn2.Context(prm)
    db = torch.from_numpy(snapshot.to_body(pos, vel)).cuda()
    dctx.force_step_records(db.data_ptr(), 12, n)
    a_init, dloga = 1.0 / 50.0, 0.02
    for loop in range(3):
        dkh, dd = cosmology.step_factors(loop, dloga, a_init, float(g["OmegaM0"]), float(g["OmegaX0"]), grav)
        # host: src/photoNs.c:158-196, force, :257-268
        hb[:, 6:9] += hb[:, 9:12] * dkh
        hb[:, 6:9] += hb[:, 3:6] * dkh
        hb[:, 0:3] += hb[:, 6:9] * dd
        p = hb[:, 0:3]
        while (p < 0.0).any():
            p[p < 0.0] += box
        while (p >= box).any():
            p[p >= box] -= box
        hb[:, 3:6] = hctx.force_step(np.ascontiguousarray(hb[:, 0:3]))
        hb[:, 6:9] += hb[:, 3:6] * dkh
        hb[:, 6:9] += hb[:, 9:12] * dkh
        # device
        dctx.kick_device(db.data_ptr(), n, dkh, True)
        dctx.drift_device(db.data_ptr(), n, dd, box)
        dctx.force_step_records(db.data_ptr(), 12, n)
        dctx.kick_device(db.data_ptr(), n, dkh, False)
    dctx.sync()
    out = db.cpu().numpy()
    np.testing.assert_array_equal(out, hb)
    assert np.abs(out[:, 3:6]).max() > 0 and not out[:, 9:12].any()
