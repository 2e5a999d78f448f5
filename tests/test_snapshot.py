"""Gadget-2 snapshot reader / writer of photons-2.0_b200/snapshot.py against the UNMODIFIED
reference's read_Particle_Gadget2 / write_Particle_Gadget2 (tests/golden/make_snapshot_golden.py): a file written by
the reference is read back bit for bit, our writer reproduces its particle blocks and header fields byte for byte,
and (in the build container, where /root/reference exists) the demo IC is read like the reference reads it."""
import os

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")
DEMO = "/root/reference/demo/ic_lcdm.gdt2"


def test_reads_a_file_written_by_the_reference():
    import snapshot
    g = np.load(os.path.join(GOLD, "snapshot_golden.npz"))
    pos, vel, head = snapshot.read_gadget2(os.path.join(GOLD, "ref_written_256.gdt2"))
    assert pos.shape == (256, 3)
    np.testing.assert_array_equal(pos, g["pos_first"][:256])
    # the reference's writer divides by a^1.5 and rounds to float32, the reader multiplies again: one float32 rounding
    np.testing.assert_allclose(vel, g["vel_first"][:256], rtol=2e-7, atol=0)
    assert head["BoxSize"] == g["BOXSIZE"] and head["Omega0"] == g["OmegaM0"] and head["OmegaLambda"] == g["OmegaX0"]
    assert head["HubbleParam"] == g["Hubble0"] and head["redshift"] == g["InitialTime"]
    assert head["mass"][1] == g["MASSPART"] and int(head["npart"][1]) == 256 and int(head["npartTotal"][1]) == int(g["NPART_TOTAL"])


def test_writer_reproduces_the_reference_blocks(tmp_path):
    import snapshot
    g = np.load(os.path.join(GOLD, "snapshot_golden.npz"))
    out = tmp_path / "ours.gdt2"
    snapshot.write_gadget2(str(out), g["pos_first"][:256], g["vel_first"][:256], float(g["BOXSIZE"]), float(g["MASSPART"]),
                           float(g["InitialTime"]), float(g["OmegaM0"]), float(g["OmegaX0"]), float(g["Hubble0"]),
                           npart_total=int(g["NPART_TOTAL"]))
    ours = open(out, "rb").read()
    ref = open(os.path.join(GOLD, "ref_written_256.gdt2"), "rb").read()
    assert len(ours) == len(ref) == 4 + 256 + 4 + 2 * (4 + 256 * 12 + 4)
    # particle blocks byte for byte (the record markers of the reference are uninitialised ints)
    p0 = 4 + 256 + 4 + 4
    assert ours[p0:p0 + 3072] == ref[p0:p0 + 3072]
    v0 = p0 + 3072 + 4 + 4
    assert ours[v0:v0 + 3072] == ref[v0:v0 + 3072]
    # header: every field the reference sets (its flag_* and fill bytes are uninitialised stack)
    ho, hr = np.frombuffer(ours[4:260], snapshot.HEADER)[0], np.frombuffer(ref[4:260], snapshot.HEADER)[0]
    for k in ("npart", "mass", "time", "redshift", "npartTotal", "num_files", "BoxSize", "Omega0", "OmegaLambda", "HubbleParam"):
        np.testing.assert_array_equal(ho[k], hr[k])
    # round trip through our own reader
    pos, vel, _ = snapshot.read_gadget2(str(out), 10, 100)
    np.testing.assert_array_equal(pos, g["pos_first"][10:110])
    body = snapshot.to_body(pos, vel)
    assert body.shape == (100, 12) and np.array_equal(body[:, 6:9], vel) and not body[:, 3:6].any()
    with pytest.raises(ValueError):
        snapshot.read_gadget2(str(out), 200, 100)


@pytest.mark.skipif(not os.path.exists(DEMO), reason="needs /root/reference (build container only)")
def test_demo_ic_like_the_reference():
    import snapshot
    g = np.load(os.path.join(GOLD, "snapshot_golden.npz"))
    pos, vel, head = snapshot.read_gadget2(DEMO)
    np.testing.assert_array_equal(pos, np.load(os.path.join(GOLD, "demo_pos_f32.npy")).astype(np.float64))
    np.testing.assert_array_equal(vel[:512], g["vel_first"])                 # bit-identical to the reference's reader
    np.testing.assert_array_equal(vel.sum(axis=0), g["vel_sum"])
    p2, v2, _ = snapshot.read_gadget2(DEMO, int(g["sub_start"]), 300)
    np.testing.assert_array_equal(p2, g["sub_pos"])
    np.testing.assert_array_equal(v2, g["sub_vel"])
    assert snapshot.read_header(DEMO)["BoxSize"] == g["BOXSIZE"]


@pytest.mark.gpu
def test_snapshot_to_device_records_and_back(tmp_path):
    """The device path of SURVEY.md 8f.4: the reference-written file becomes device-resident Body records
    (pn2_snapshot_to_body_device) bit-identical to the host reader's, and the records written back from the device
    (pn2_body_to_snapshot_device) reproduce the reference writer's particle blocks byte for byte."""
    import pn2gpu
    import snapshot
    g = np.load(os.path.join(GOLD, "snapshot_golden.npz"))
    path = os.path.join(GOLD, "ref_written_256.gdt2")
    pos, vel, head = snapshot.read_gadget2(path)
    # 256 particles in the whole box: leaves far wider than the split scale -> FP64 mode (the FP32 tile layout refuses them)
    ctx = pn2gpu.Context(pn2gpu.make_params(float(g["BOXSIZE"]), 32, 256, float(g["MASSPART"]), precision=pn2gpu.FP64))
    body, head_d = snapshot.load_body_device(ctx, path)
    hb = body.cpu().numpy()
    np.testing.assert_array_equal(hb, snapshot.to_body(pos, vel))
    np.testing.assert_array_equal(hb[:, 0:3], g["pos_first"][:256])
    sub, _ = snapshot.load_body_device(ctx, path, 10, 100)
    np.testing.assert_array_equal(sub.cpu().numpy(), hb[10:110])
    # a force step straight on the loaded records
    ctx.force_step_records(body.data_ptr(), 12, 256)
    ctx.sync()
    assert np.abs(body.cpu().numpy()[:, 3:6]).max() > 0
    # back to a file
    import torch
    rec = torch.from_numpy(snapshot.to_body(g["pos_first"][:256], g["vel_first"][:256])).cuda()
    out = tmp_path / "dev.gdt2"
    snapshot.save_body_device(ctx, str(out), rec, float(g["BOXSIZE"]), float(g["MASSPART"]), float(g["InitialTime"]), float(g["OmegaM0"]),
                              float(g["OmegaX0"]), float(g["Hubble0"]), npart_total=int(g["NPART_TOTAL"]))
    ours, ref = open(out, "rb").read(), open(path, "rb").read()
    p0 = 4 + 256 + 4 + 4
    v0 = p0 + 3072 + 4 + 4
    assert len(ours) == len(ref) and ours[p0:p0 + 3072] == ref[p0:p0 + 3072] and ours[v0:v0 + 3072] == ref[v0:v0 + 3072]
    ctx.close()
