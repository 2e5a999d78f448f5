"""Gadget-2 snapshot files <-> particle arrays, following the reference's reader and writer
(src/snapshot.c:5-22 header, :211-293 read_Particle_Gadget2, :397-503 write_Particle_Gadget2):

    int32 256 | header (256 bytes) | int32 256
    int32 m   | float32 pos[N][3]  | int32 m          all six particle types back to back
    int32 m   | float32 vel[N][3]  | int32 m          file velocities = v / a^1.5 (gdt2unit, :261, :469)
    (ids, masses: not read by the reference)

Positions become float64 exactly (float32 values widened), velocities float64(float32) * a^1.5, as in the reference.
`to_body` lays the arrays out as the reference's Body records (pos, acc, vel, acc_pm: 12 doubles) for the device calls
that take records (pn2_migrate_*, pn2_kick_device, pn2_drift_device)."""
import numpy as np

HEADER = np.dtype([("npart", "<i4", 6), ("mass", "<f8", 6), ("time", "<f8"), ("redshift", "<f8"), ("flag_sfr", "<i4"),
                   ("flag_feedback", "<i4"), ("npartTotal", "<i4", 6), ("flag_cooling", "<i4"), ("num_files", "<i4"),
                   ("BoxSize", "<f8"), ("Omega0", "<f8"), ("OmegaLambda", "<f8"), ("HubbleParam", "<f8"),
                   ("fill", "V96")])
assert HEADER.itemsize == 256


def read_header(path):
    """src/snapshot.c:62-119: the 256-byte header between two record markers."""
    with open(path, "rb") as f:
        f.read(4)
        return np.frombuffer(f.read(256), HEADER)[0]


def read_gadget2(path, n_start=0, n_count=None):
    """(pos, vel, header) of particles [n_start, n_start + n_count) in file order (src/snapshot.c:211-293)."""
    with open(path, "rb") as f:
        f.read(4)
        head = np.frombuffer(f.read(256), HEADER)[0]
        f.read(4)
        n = int(head["npart"].sum())
        if n_count is None:
            n_count = n - n_start
        if n_start < 0 or n_count < 0 or n_start + n_count > n:
            raise ValueError(f"particles [{n_start}, {n_start + n_count}) outside the file's {n}")
        f.read(4)
        pos = np.frombuffer(f.read(12 * n), "<f4").reshape(n, 3)
        f.read(4)
        f.read(4)
        vel = np.frombuffer(f.read(12 * n), "<f4").reshape(n, 3)
    gdt2unit = (1.0 / (1.0 + float(head["redshift"]))) ** 1.5
    sl = slice(n_start, n_start + n_count)
    return pos[sl].astype(np.float64), vel[sl].astype(np.float64) * gdt2unit, head


def write_gadget2(path, pos, vel, box, mass, redshift, omega0, omega_lambda, hubble, npart_total=None):
    """All particles as type 1, as the reference writes them (src/snapshot.c:397-503); record markers carry the
    record length (the reference leaves them uninitialised; its reader ignores their value)."""
    pos = np.asarray(pos, np.float64)
    vel = np.asarray(vel, np.float64)
    n = pos.shape[0]
    head = np.zeros(1, HEADER)
    head["npart"][0, 1] = n
    head["mass"][0, 1] = mass
    head["npartTotal"][0, 1] = n if npart_total is None else npart_total
    head["time"] = 1.0 / (redshift + 1.0)
    head["redshift"] = redshift
    head["num_files"] = 1
    head["BoxSize"] = box
    head["Omega0"] = omega0
    head["OmegaLambda"] = omega_lambda
    head["HubbleParam"] = hubble
    gdt2unit = (1.0 / (1.0 + redshift)) ** 1.5
    p32 = pos.astype("<f4")
    v32 = (vel.astype(np.float32).astype(np.float64) / gdt2unit).astype("<f4")      # (float)v / gdt2unit, stored as float
    with open(path, "wb") as f:
        for payload in (head.tobytes(), p32.tobytes(), v32.tobytes()):
            m = np.array([len(payload)], "<i4").tobytes()
            f.write(m)
            f.write(payload)
            f.write(m)


def to_body(pos, vel=None):
    """(n, 12) float64 Body records (inc/typesdef.h:25-31): pos 0-2, acc 3-5, vel 6-8, acc_pm 9-11."""
    n = pos.shape[0]
    body = np.zeros((n, 12))
    body[:, 0:3] = pos
    if vel is not None:
        body[:, 6:9] = vel
    return body


def read_blocks(path, n_start=0, n_count=None):
    """The file's float32 blocks as they are (pos32, vel32, header, gdt2unit): what travels to the device."""
    with open(path, "rb") as f:
        f.read(4)
        head = np.frombuffer(f.read(256), HEADER)[0]
        f.read(4)
        n = int(head["npart"].sum())
        if n_count is None:
            n_count = n - n_start
        if n_start < 0 or n_count < 0 or n_start + n_count > n:
            raise ValueError(f"particles [{n_start}, {n_start + n_count}) outside the file's {n}")
        p0 = f.tell() + 4
        f.seek(p0 + 12 * n_start)
        pos = np.frombuffer(f.read(12 * n_count), "<f4").reshape(n_count, 3)
        f.seek(p0 + 12 * n + 8 + 12 * n_start)
        vel = np.frombuffer(f.read(12 * n_count), "<f4").reshape(n_count, 3)
    return pos, vel, head, (1.0 / (1.0 + float(head["redshift"]))) ** 1.5


def load_body_device(ctx, path, n_start=0, n_count=None):
    """Snapshot -> device-resident Body records (torch CUDA tensor (n, 12) float64) for pn2_force_step_records /
    pn2_pm_force_records / pn2_kick_device / pn2_drift_device / pn2_migrate_*: the float32 blocks are uploaded from
    pinned memory and widened / scaled on the device (pn2_snapshot_to_body_device).  Returns (body, header)."""
    import torch
    import pn2gpu
    pos, vel, head, unit = read_blocks(path, n_start, n_count)
    n = pos.shape[0]
    dev = torch.device("cuda", ctx.device_index())
    hp = torch.from_numpy(np.array(pos)).pin_memory()            # np.frombuffer views are read-only: copy
    hv = torch.from_numpy(np.array(vel)).pin_memory()
    dp, dv = hp.to(dev, non_blocking=True), hv.to(dev, non_blocking=True)
    body = torch.empty((n, 12), dtype=torch.float64, device=dev)
    torch.cuda.synchronize(dev)
    pn2gpu._ck(pn2gpu.lib().pn2_snapshot_to_body_device(ctx.h, dp.data_ptr(), dv.data_ptr(), n, unit, body.data_ptr()))
    ctx.sync()
    return body, head


def save_body_device(ctx, path, body, box, mass, redshift, omega0, omega_lambda, hubble, npart_total=None):
    """Device-resident Body records -> snapshot file: converted to the file's float32 blocks on the device
    (pn2_body_to_snapshot_device), one D2H copy of 24 bytes per particle, written like write_gadget2."""
    import torch
    import pn2gpu
    n = body.shape[0]
    unit = (1.0 / (1.0 + redshift)) ** 1.5
    dp = torch.empty((n, 3), dtype=torch.float32, device=body.device)
    dv = torch.empty((n, 3), dtype=torch.float32, device=body.device)
    torch.cuda.synchronize(body.device)
    pn2gpu._ck(pn2gpu.lib().pn2_body_to_snapshot_device(ctx.h, body.data_ptr(), n, unit, dp.data_ptr(), dv.data_ptr()))
    ctx.sync()
    head = np.zeros(1, HEADER)
    head["npart"][0, 1] = n
    head["mass"][0, 1] = mass
    head["npartTotal"][0, 1] = n if npart_total is None else npart_total
    head["time"] = 1.0 / (redshift + 1.0)
    head["redshift"] = redshift
    head["num_files"] = 1
    head["BoxSize"] = box
    head["Omega0"] = omega0
    head["OmegaLambda"] = omega_lambda
    head["HubbleParam"] = hubble
    with open(path, "wb") as f:
        for payload in (head.tobytes(), dp.cpu().numpy().tobytes(), dv.cpu().numpy().tobytes()):
            m = np.array([len(payload)], "<i4").tobytes()
            f.write(m)
            f.write(payload)
            f.write(m)
