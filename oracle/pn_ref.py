"""Driver for the compiled UNMODIFIED reference (oracle/_ref).  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this module; the product path never does.
"""
import os
import subprocess
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, "_ref")
REF_EXE = os.path.join(REF_DIR, "ref_fmm")
REF_LIB = os.path.join(REF_DIR, "libphotons_ref.so")
REF_EXE_OPEN = os.path.join(REF_DIR, "ref_fmm_open")   # built without -DPERIODIC_CONDITION -DLONGSHORT
REF_EXE_PM = os.path.join(REF_DIR, "ref_pm")          # src/partmesh.c unmodified + C restatement of conv.f90's convolution
REF_EXE_GPU = os.path.join(REF_DIR, "ref_fmm_gpu")   # the same reference with its task batches routed to libpn2gpu.so

NMULTI = 20
# struct layouts of the reference (inc/typesdef.h:25-57, inc/photoNs.h:177-189,285-291); sizes probed: SURVEY.md 8
BODY = np.dtype([("pos", "f8", 3), ("acc", "f8", 3), ("vel", "f8", 3), ("acc_pm", "f8", 3)])
PACK = np.dtype([("npart", "i4"), ("ipart", "i4"), ("width", "f8", 3), ("center", "f8", 3),
                 ("M", "f8", NMULTI), ("L", "f8", NMULTI)])
NODE = np.dtype([("updated", "i4"), ("npart", "i4"), ("son", "i4", 2), ("split", "f8"), ("width", "f8", 3),
                 ("center", "f8", 3), ("M", "f8", NMULTI), ("L", "f8", NMULTI)])
RNODE = np.dtype([("npart", "i4"), ("son", "i4", 2), ("pad", "i4"), ("width", "f8", 3), ("center", "f8", 3),
                  ("M", "f8", NMULTI)])
RBODY = np.dtype([("pos", "f8", 3), ("replenish", "f8")])
TOPNODE = np.dtype([("son", "i4", 2), ("split", "f8"), ("M", "f8", NMULTI), ("center", "f8", 3),
                    ("width", "f8", 3)])
assert BODY.itemsize == 96 and PACK.itemsize == 376 and NODE.itemsize == 392
assert RNODE.itemsize == 224 and RBODY.itemsize == 32 and TOPNODE.itemsize == 224

SCALARS = ["box", "rs", "cutoff", "soft", "theta", "mass", "maxleaf", "nside", "rank", "size", "npart",
           "first_leaf", "last_leaf", "first_node", "last_node", "idxP2P", "idxM2L", "p2p_count_remote",
           "walk_m2l_count", "this_domain", "direct_local_start", "mostleft", "nint_local", "repeat"]


def available():
    return os.path.exists(REF_EXE)


def read_gadget2_positions(path):
    """Positions (float64 copy of the float32 block) + header fields of a Gadget-2 snapshot.
    Wire format as read by the reference: src/snapshot.c:64-119 (header), :211-293 (blocks)."""
    with open(path, "rb") as f:
        raw = f.read()
    assert np.frombuffer(raw, "i4", 1, 0)[0] == 256
    npart = np.frombuffer(raw, "i4", 6, 4)
    mass = np.frombuffer(raw, "f8", 6, 28)
    box = np.frombuffer(raw, "f8", 1, 4 + 24 + 48 + 16 + 8 + 24 + 8)[0]
    n = int(npart.sum())
    off = 4 + 256 + 4
    nb = np.frombuffer(raw, "i4", 1, off)[0]
    assert nb == n * 12
    pos = np.frombuffer(raw, "f4", 3 * n, off + 4).reshape(n, 3).astype(np.float64)
    return pos, {"npart": npart.copy(), "mass": mass.copy(), "box": float(box)}


def _read_records(path):
    out = {}
    with open(path, "rb") as f:
        raw = f.read()
    off = 0
    while off < len(raw):
        name = raw[off:off + 24].split(b"\0")[0].decode()
        nb = int(np.frombuffer(raw, "i8", 1, off + 24)[0])
        out.setdefault(name, []).append(raw[off + 32:off + 32 + nb])
        off += 32 + nb
    return out


def _parse_rank(path):
    r = _read_records(path)
    d = {}
    sc = np.frombuffer(r["scalars"][0], "f8")
    for k, v in zip(SCALARS, sc):
        d[k] = float(v) if k in ("box", "rs", "cutoff", "soft", "theta", "mass") else int(v)
    d["timing"] = dict(zip(["construct", "prepare", "task", "ext", "total"], np.frombuffer(r["timing"][0], "f8")))
    d["part"] = np.frombuffer(r["part"][0], BODY).copy()
    d["toptree"] = np.frombuffer(r["toptree"][0], TOPNODE).copy()
    if "leaf" in r:
        d["leaf"] = np.frombuffer(r["leaf"][0], PACK).copy()
        d["btree"] = np.frombuffer(r["btree"][0], NODE).copy()
        for k in ("p2p_s", "p2p_t", "m2l_s", "m2l_t"):
            d[k] = np.frombuffer(r[k][0], "i4").copy()
    if "rcap_hdr" in r:
        hdr = np.frombuffer(r["rcap_hdr"][0], "i8").reshape(-1, 5)
        rc = []
        for i, h in enumerate(hdr):
            rc.append({"seq": int(h[0]),
                       "tree": np.frombuffer(r["rcap_tree"][i], RNODE).copy(),
                       "body": np.frombuffer(r["rcap_body"][i], RBODY).copy(),
                       "p2p_s": np.frombuffer(r["rcap_p2p_s"][i], "i4").copy(),
                       "p2p_t": np.frombuffer(r["rcap_p2p_t"][i], "i4").copy(),
                       "m2l_s": np.frombuffer(r["rcap_m2l_s"][i], "i4").copy(),
                       "m2l_t": np.frombuffer(r["rcap_m2l_t"][i], "i4").copy()})
        d["remote"] = rc
    return d


def run_reference(pos, box, nside, mass, maxleaf=8, theta=0.4, split=-1.0, soft=-1.0, nranks=1, capture=0,
                  repeat=1, workdir=None, timeout=3600, gpu=False, env_extra=None, open_newtonian=False):
    """Run one short-range force evaluation of the unmodified reference on `pos` (N x 3 float64, in
    [0, box)) at `nranks` ranks of the fork/socketpair mini-MPI.  Returns a list with one dict per rank."""
    exe = REF_EXE_GPU if gpu else (REF_EXE_OPEN if open_newtonian else REF_EXE)
    if not os.path.exists(exe):
        raise RuntimeError("oracle/_ref/ref_fmm is not built (run `make -C oracle ref` where /root/reference exists)")
    pos = np.ascontiguousarray(pos, dtype=np.float64)
    n = pos.shape[0]
    with tempfile.TemporaryDirectory(dir=workdir) as td:
        pfile = os.path.join(td, "params.txt")
        with open(pfile, "w") as f:
            f.write(f"NPART_TOTAL {n}\nBOXSIZE {float(box)!r}\nNSIDE {int(nside)}\nMAXLEAF {int(maxleaf)}\n"
                    f"OPENANGLE {float(theta)!r}\nSPLIT {float(split)!r}\nSOFT {float(soft)!r}\n"
                    f"MASS {float(mass)!r}\nCAPTURE {int(capture)}\nREPEAT {int(repeat)}\n")
        xfile = os.path.join(td, "pos.f64")
        pos.tofile(xfile)
        env = dict(os.environ, PN_SHIM_NP=str(nranks))
        env.update(env_extra or {})
        res = subprocess.run([exe, pfile, xfile, os.path.join(td, "out")], env=env, cwd=td,
                             capture_output=True, text=True, timeout=timeout)
        if res.returncode != 0:
            raise RuntimeError(f"ref_fmm failed rc={res.returncode}\n{res.stdout[-2000:]}\n{res.stderr[-2000:]}")
        ranks = [_parse_rank(os.path.join(td, f"out.r{r}.bin")) for r in range(nranks)]
    for d in ranks:
        d["stdout"] = res.stdout
    return ranks


def gather_acc(ranks, n):
    """Accelerations in input order (ids were carried in Body.vel[0] by the harness)."""
    acc = np.full((n, 3), np.nan)
    for d in ranks:
        ids = d["part"]["vel"][:, 0].astype(np.int64)
        acc[ids] = d["part"]["acc"]
    return acc


def run_reference_pm(pos, box, nside, mass, split=-1.0, workdir=None, timeout=3600):
    """The reference's particle-mesh force (partmesh_thread, src/partmesh.c:18-796, compiled unmodified; its Fortran
    convolution restated in C, oracle/ref_shim/ref_pm_harness.c) on one rank.  Returns {"acc_pm": (n, 3) in input
    order, "density", "potential": (nside,)*3 -- the convolution's input and output --, "sec"}."""
    if not os.path.exists(REF_EXE_PM):
        raise RuntimeError("oracle/_ref/ref_pm is not built (run `make -C oracle ref` where /root/reference exists)")
    pos = np.ascontiguousarray(pos, dtype=np.float64)
    n = pos.shape[0]
    with tempfile.TemporaryDirectory(dir=workdir) as td:
        pfile = os.path.join(td, "params.txt")
        with open(pfile, "w") as f:
            f.write(f"NPART_TOTAL {n}\nBOXSIZE {float(box)!r}\nNSIDE {int(nside)}\nMASS {float(mass)!r}\nSPLIT {float(split)!r}\n")
        xfile = os.path.join(td, "pos.f64")
        pos.tofile(xfile)
        res = subprocess.run([REF_EXE_PM, pfile, xfile, os.path.join(td, "out.bin")], env=dict(os.environ, PN_SHIM_NP="1"), cwd=td,
                             capture_output=True, text=True, timeout=timeout)
        if res.returncode != 0:
            raise RuntimeError(f"ref_pm failed rc={res.returncode}\n{res.stdout[-2000:]}\n{res.stderr[-2000:]}")
        r = _read_records(os.path.join(td, "out.bin"))
    return {"acc_pm": np.frombuffer(r["acc_pm"][0], "f8").reshape(n, 3).copy(),
            "density": np.frombuffer(r["density"][0], "f8").reshape(nside, nside, nside).copy(),
            "potential": np.frombuffer(r["potential"][0], "f8").reshape(nside, nside, nside).copy(),
            "sec": float(np.frombuffer(r["timing"][0], "f8")[0])}
