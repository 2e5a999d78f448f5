"""Generate tests/golden/domains_golden.npz from the UNMODIFIED reference's domain decomposition
(oracle/_ref/libphotons_ref.so = /root/reference/src/domains.c, initial.c compiled in place).

Run in the build container only (needs /root/reference):  python tests/golden/make_domain_golden.py

For every case (P ranks, a sequence of per-rank load fractions = DTIME_FRACTION of src/photoNs.c:283):
  split0     the splits of domain_initialize()                     (src/domains.c:432-470)
  splits[k]  after measure_domain_runtime(load k) + determine_split_domtree(), applied in sequence
             (src/domains.c:21-38, 86-160; the MPI_Allgather is replaced by writing time_node directly)
  owner[k]   the destination rank of every test position under splits[k], and sendcount[k], obtained by running
             prepare_body_inOrderOf_domain() (src/domains.c:268-296) on a Body array tagged with the input index
The test positions are every 16th particle of the demo IC (2048 positions).
"""
import ctypes as C
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.join(HERE, "..", "..")
LIB = os.path.join(ROOT, "oracle", "_ref", "libphotons_ref.so")


class DomainNode(C.Structure):      # inc/photoNs.h:300-307
    _fields_ = [("son", C.c_int * 2), ("time_node", C.c_double), ("time_left", C.c_double), ("time_right", C.c_double),
                ("split", C.c_double), ("split_previous", C.c_double)]


CASES = [
    (2, [[1.3, 0.7], [0.9, 1.1]]),
    (3, [[1.5, 0.9, 0.6], [1.0, 1.0, 1.0], [0.8, 1.3, 0.9]]),
    (4, [[1.2, 0.8, 1.1, 0.9], [0.7, 1.4, 1.0, 0.9]]),
    (5, [[1.6, 0.8, 0.9, 0.7, 1.0], [1.1, 1.0, 0.9, 1.2, 0.8]]),
    (8, [[1.4, 0.6, 1.1, 0.9, 1.3, 0.7, 1.0, 1.0], [0.9, 1.1, 1.0, 1.2, 0.8, 1.0, 0.9, 1.1], [1.0] * 8]),
    (6, [[1.3, 0.7, 1.1, 0.9, 1.2, 0.8], [0.9, 1.0, 1.1, 1.0, 0.8, 1.2]]),
    (7, [[0.6, 1.4, 1.0, 1.1, 0.9, 1.2, 0.8], [1.0] * 7]),
    (16, [[1.0 + 0.05 * ((3 * k) % 7 - 3) for k in range(16)], [1.0 + 0.04 * ((5 * k) % 9 - 4) for k in range(16)]]),
]


def main():
    L = C.CDLL(LIB)
    box = 100000.0
    pos = np.load(os.path.join(HERE, "demo_pos_f32.npy")).astype(np.float64)[::16].copy()
    n = len(pos)
    out = {"pos_stride": 16, "box": box, "ncase": len(CASES)}
    for ci, (P, loads) in enumerate(CASES):
        C.c_int.in_dll(L, "PROC_SIZE").value = P
        C.c_int.in_dll(L, "PROC_RANK").value = 0
        C.c_double.in_dll(L, "BOXSIZE").value = box
        L.reset_mem()
        L.setup_domain_index()
        L.domain_initialize()
        dt = C.POINTER(DomainNode).in_dll(L, "domtree")
        ml = C.c_int.in_dll(L, "mostleft").value
        get = lambda: np.array([dt[k].split for k in range(2 * P - 1)])
        split0 = get()
        splits, owners, counts = [], [], []
        L.fill_time_domtree.restype = C.c_double
        for load in loads:
            for r in range(P):                       # measure_domain_runtime without the Allgather
                idom = r + ml
                if idom > 2 * P - 2:
                    idom -= P
                dt[idom].time_node = load[r]
            L.fill_time_domtree(0)
            L.determine_split_domtree(P, 0, dt)
            splits.append(get())
            body = np.zeros((n, 12))                 # Body = pos, acc, vel, acc_pm (96 bytes)
            body[:, :3] = pos
            body[:, 6] = np.arange(n)                # vel[0] carries the input index
            send = (C.c_int * P)()
            L.prepare_body_inOrderOf_domain(0, body.ctypes.data_as(C.c_void_p), n, 0, send)
            sc = np.array(list(send))
            assert sc.sum() == n
            own = np.zeros(n, np.int32)
            own[body[:, 6].astype(np.int64)] = np.repeat(np.arange(P), sc)
            owners.append(own)
            counts.append(sc)
        out[f"P{ci}"] = P
        out[f"loads{ci}"] = np.array(loads)
        out[f"split0_{ci}"] = split0
        out[f"splits{ci}"] = np.array(splits)
        out[f"owner{ci}"] = np.array(owners)
        out[f"sendcount{ci}"] = np.array(counts)
        print(P, "split0", split0[:P - 1], "->", splits[-1][:P - 1], "sendcount", counts[-1])
    np.savez_compressed(os.path.join(HERE, "domains_golden.npz"), **out)


if __name__ == "__main__":
    main()
