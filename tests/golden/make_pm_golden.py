"""Generate tests/golden/pm_*.npz from the reference's particle-mesh force: src/partmesh.c compiled UNMODIFIED
(partmesh_thread: CIC deposit, mesh exchange, 4-point gradient, CIC gather) around a C restatement of the Fortran
convolution of src/conv.f90:128-247 (2DECOMP&FFT is absent: oracle/ref_shim/ref_pm_harness.c, "parity unpinned" for
that one function).  Run in the build container only:  python tests/golden/make_pm_golden.py

  pm_demo_ns32.npz   demo/ic_lcdm.gdt2 (N = 32768), NSIDE 32 (demo/lcdm_g2.run): acc_pm in input order, the density mesh
                     handed to the convolution and the potential it returns
  pm_small_ns24.npz  every 8th demo particle, NSIDE 24 (not a power of two: mixed-radix path of the transforms)
  pm_pair_ns64.npz   two particles 5 length units apart in a BOX 100, NSIDE 64 mesh: the physics check
                     (PM + short-range = Newtonian)
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
from oracle import pn_ref  # noqa: E402


def main():
    pos, hd = pn_ref.read_gadget2_positions("/root/reference/demo/ic_lcdm.gdt2")
    box, mass = hd["box"], float(hd["mass"][1])
    r = pn_ref.run_reference_pm(pos, box, 32, mass)
    np.savez_compressed(os.path.join(HERE, "pm_demo_ns32.npz"), acc_pm=r["acc_pm"], density=r["density"], potential=r["potential"],
                        box=box, mass=mass, nside=32)
    print("demo ns32: rms |acc_pm| = %.10e, mean density %.6e" % (np.sqrt((r["acc_pm"] ** 2).sum(1).mean()), r["density"].mean()))
    sp = pos[::8].copy()
    r = pn_ref.run_reference_pm(sp, box, 24, 2.5)
    np.savez_compressed(os.path.join(HERE, "pm_small_ns24.npz"), acc_pm=r["acc_pm"], density=r["density"], potential=r["potential"],
                        box=box, mass=2.5, nside=24)
    print("small ns24: rms |acc_pm| = %.10e" % np.sqrt((r["acc_pm"] ** 2).sum(1).mean()))
    pp = np.array([[50.2, 50.3, 50.1], [50.2 + 3.0, 50.3 + 4.0, 50.1]])
    r = pn_ref.run_reference_pm(pp, 100.0, 64, 1.0)
    np.savez_compressed(os.path.join(HERE, "pm_pair_ns64.npz"), pos=pp, acc_pm=r["acc_pm"], box=100.0, mass=1.0, nside=64)
    print("pair:", r["acc_pm"])


if __name__ == "__main__":
    main()
