"""Generate tests/golden/* from the UNMODIFIED reference (oracle/_ref/ref_fmm, built from /root/reference).

Run in the build container only (needs /root/reference):  python tests/golden/make_golden.py
The fixtures are what pins the CPU oracle (oracle/port) and, through it, the CUDA path on the GPU box,
where /root/reference does not exist.

Fixtures
  demo_pos_f32.npy            positions of demo/ic_lcdm.gdt2 (float32 block of the Gadget-2 file, N=32768)
  demo_ns32_np{1,2,3,4,8}.npz, demo_ns16_np{1,2,4}.npz
                              accelerations in input order + counters of one short-range force evaluation
                              (demo/lcdm_g2.run parameters: MaxPackage 8, OPENANGLE 0.4; NSIDE 16 exercises M2L)
  small_{t04,t12}_np{1,2}.npz every 8th demo particle (N=4096), NSIDE 24, theta 0.4 / 1.2: accelerations, full tree
                              arrays (leaves, nodes, M, L after the step) and full local P2P/M2L lists per rank
  small_{t04,t12}_let_np2.npz every received pruned tree / ghost bodies / remote lists of the NP=2 run, rank 0
"""
import hashlib
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
from oracle import pn_ref  # noqa: E402

REF_DEMO = "/root/reference/demo/ic_lcdm.gdt2"


def pair_hash(s, t):
    h = hashlib.sha256()
    h.update(np.ascontiguousarray(s, np.int32).tobytes())
    h.update(np.ascontiguousarray(t, np.int32).tobytes())
    return h.hexdigest()


def counters(ranks):
    return {k: np.array([d[k] for d in ranks], np.int64) for k in
            ("npart", "first_leaf", "last_leaf", "first_node", "last_node", "idxP2P", "idxM2L", "p2p_count_remote",
             "walk_m2l_count", "nint_local")}


def main():
    pos, hd = pn_ref.read_gadget2_positions(REF_DEMO)
    box, mass = hd["box"], float(hd["mass"][1])
    np.save(os.path.join(HERE, "demo_pos_f32.npy"), pos.astype(np.float32))
    assert np.array_equal(pos.astype(np.float32).astype(np.float64), pos)
    for nside in (32, 16):
        for nranks in (1, 2, 3, 4, 8):
            if nside == 16 and nranks in (3, 8):
                continue
            r = pn_ref.run_reference(pos, box, nside, mass, maxleaf=8, theta=0.4, nranks=nranks, capture=1, timeout=600)
            acc = pn_ref.gather_acc(r, len(pos))
            c = counters(r)
            extra = {}
            if nranks == 1:
                d = r[0]
                extra = {"p2p_hash": pair_hash(d["p2p_s"], d["p2p_t"]), "m2l_hash": pair_hash(d["m2l_s"], d["m2l_t"]),
                         "leaf_npart": d["leaf"]["npart"].astype(np.int8), "first_pos": d["part"]["pos"][0],
                         "first_acc": d["part"]["acc"][0]}
            np.savez_compressed(os.path.join(HERE, f"demo_ns{nside}_np{nranks}.npz"), acc=acc, box=box, mass=mass,
                                nside=nside, maxleaf=8, theta=0.4, **c, **extra)
            print(f"demo nside={nside} np={nranks}: rms|acc|={np.sqrt((acc**2).sum(1).mean()):.10e}", {k: v.tolist() for k, v in c.items()})

    # small set: every 8th demo particle (N=4096).  NSIDE 24 keeps the cut-off small enough for the
    # reference's LET receive buffers (sized from the local NNODE, src/fmm.c:1004-1011; larger cut-offs
    # overrun them and the reference segfaults).  theta 1.2 forces local and remote M2L pairs.
    spos = pos[::8].copy()
    sbox, smass, snside = box, 2.5, 24
    for theta, tag in ((0.4, "t04"), (1.2, "t12")):
        for nranks in (1, 2):
            r = pn_ref.run_reference(spos, sbox, snside, smass, maxleaf=8, theta=theta, nranks=nranks, capture=2,
                                     timeout=120)
            acc = pn_ref.gather_acc(r, len(spos))
            out = {"acc": acc, "box": sbox, "mass": smass, "nside": snside, "maxleaf": 8, "theta": theta}
            out.update(counters(r))
            for k, d in enumerate(r):
                out[f"r{k}_ids"] = d["part"]["vel"][:, 0].astype(np.int64)
                for f in ("npart", "ipart", "width", "center", "M", "L"):
                    out[f"r{k}_leaf_{f}"] = d["leaf"][f]
                for f in ("npart", "son", "split", "width", "center", "M", "L"):
                    out[f"r{k}_node_{f}"] = d["btree"][f]
                for f in ("p2p_s", "p2p_t", "m2l_s", "m2l_t"):
                    out[f"r{k}_{f}"] = d[f]
            np.savez_compressed(os.path.join(HERE, f"small_{tag}_np{nranks}.npz"), **out)
            print(f"small {tag} np={nranks}: rms|acc|={np.sqrt((acc**2).sum(1).mean()):.10e}",
                  {k: v.tolist() for k, v in counters(r).items() if k.startswith(("idx", "walk", "p2p"))})
            if nranks == 2:
                let = {}
                rc = r[0]["remote"]
                let["ncap"] = len(rc)
                for i, c in enumerate(rc):
                    let[f"c{i}_seq"] = c["seq"]
                    for f in ("npart", "son", "width", "center", "M"):
                        let[f"c{i}_tree_{f}"] = c["tree"][f]
                    let[f"c{i}_body"] = c["body"]["pos"]
                    for f in ("p2p_s", "p2p_t", "m2l_s", "m2l_t"):
                        let[f"c{i}_{f}"] = c[f]
                np.savez_compressed(os.path.join(HERE, f"small_{tag}_let_np2.npz"), **let)
                print("   LET captures on rank 0:", len(rc), "remote m2l pairs:", sum(len(c["m2l_s"]) for c in rc))


def merger():
    """configs[1]: demo/ic_merger.gdt2 (60000 particles, BoxSize 0, positions in [-192, 192]) shifted by +200 into a
    BOXSIZE 400 box, run through the reference built WITHOUT -DPERIODIC_CONDITION -DLONGSHORT (oracle/_ref/ref_fmm_open):
    plain Newtonian P2P and G = 1/r M2L, no images.  All particles get MASSPART = head.mass[1] (src/snapshot.c:89)."""
    pos32, hd = pn_ref.read_gadget2_positions("/root/reference/demo/ic_merger.gdt2")
    np.save(os.path.join(HERE, "merger_pos_f32.npy"), pos32.astype(np.float32))
    assert np.array_equal(pos32.astype(np.float32).astype(np.float64), pos32)
    pos = pos32 + 200.0
    mass, box = float(hd["mass"][1]), 400.0
    for nranks in (1, 2):
        r = pn_ref.run_reference(pos, box, 8, mass, maxleaf=8, theta=0.4, nranks=nranks, capture=1, open_newtonian=True, timeout=900)
        acc = pn_ref.gather_acc(r, len(pos))
        np.savez_compressed(os.path.join(HERE, f"merger_open_np{nranks}.npz"), acc=acc, box=box, mass=mass, nside=8, maxleaf=8, theta=0.4,
                            shift=200.0, soft=r[0]["soft"], **counters(r))
        print(f"merger open np={nranks}: rms|acc|={np.sqrt((acc**2).sum(1).mean()):.10e}", {k: v.tolist() for k, v in counters(r).items() if k.startswith(("idx", "walk", "p2p", "nint"))})


if __name__ == "__main__" and len(sys.argv) > 1 and sys.argv[1] == "merger":
    merger()


if __name__ == "__main__" and len(sys.argv) == 1:
    main()
    merger()
