"""Generate tests/golden/snapshot_golden.npz and tests/golden/ref_written_256.gdt2 with the UNMODIFIED reference's
Gadget-2 reader and writer (src/snapshot.c, oracle/_ref/libphotons_ref.so) on demo/ic_lcdm.gdt2.
Run in the build container only:  python tests/golden/make_snapshot_golden.py"""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
L = C.CDLL(os.path.join(HERE, "..", "..", "oracle", "_ref", "libphotons_ref.so"))
DEMO = b"/root/reference/demo/ic_lcdm.gdt2"
N = 32768
C.c_int.in_dll(L, "PROC_RANK").value = 1            # quiet
L.read_GadgetHeader(DEMO)
body = np.zeros((N, 12))
C.c_void_p.in_dll(L, "part").value = body.ctypes.data
L.read_Particle_Gadget2(DEMO, 0, N)
hdr = {k: C.c_double.in_dll(L, k).value for k in ("BOXSIZE", "OmegaM0", "OmegaX0", "Hubble0", "InitialTime", "MASSPART")}
hdr["NPART_TOTAL"] = C.c_long.in_dll(L, "NPART_TOTAL").value
print(hdr, body[0, :3], body[0, 6:9])
# the reference's writer on the first 256 particles (uses the statics its reader filled)
C.c_double.in_dll(L, "Redshift_Time").value = hdr["InitialTime"]
out = os.path.join(HERE, "ref_written_256.gdt2")
L.write_Particle_Gadget2(out.encode(), 0, 256)
np.savez_compressed(os.path.join(HERE, "snapshot_golden.npz"), vel_first=body[:512, 6:9], pos_first=body[:512, 0:3],
                    vel_sum=body[:, 6:9].sum(axis=0), vel_abs_sum=np.abs(body[:, 6:9]).sum(axis=0),
                    sub_start=1000, sub_pos=0, **hdr)
# a ranged read, as the reference's ranks do (n_start, n_count)
body2 = np.zeros((300, 12))
C.c_void_p.in_dll(L, "part").value = body2.ctypes.data
L.read_Particle_Gadget2(DEMO, 1000, 300)
d = dict(np.load(os.path.join(HERE, "snapshot_golden.npz")))
d["sub_pos"] = body2[:, 0:3]
d["sub_vel"] = body2[:, 6:9]
np.savez_compressed(os.path.join(HERE, "snapshot_golden.npz"), **d)
print(os.path.getsize(out), "bytes written by the reference")
