"""Generate tests/golden/integrator_golden.npz from the UNMODIFIED reference's kick_loga / drift_loga
(src/initial.c:639-683, oracle/_ref/libphotons_ref.so).  Run in the build container only:
    python tests/golden/make_integrator_golden.py"""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
L = C.CDLL(os.path.join(HERE, "..", "..", "oracle", "_ref", "libphotons_ref.so"))
L.kick_loga.restype = C.c_double
L.drift_loga.restype = C.c_double
L.kick_loga.argtypes = [C.c_double, C.c_double]
L.drift_loga.argtypes = [C.c_double, C.c_double]
rows = []
for om, ox in ((0.25, 0.75), (0.3, 0.7), (1.0, 0.0)):
    C.c_double.in_dll(L, "OmegaM0").value = om
    C.c_double.in_dll(L, "OmegaX0").value = ox
    for ai, af, nstep in ((0.02, 1.0, 64), (0.5, 1.0, 10), (1.0 / 50, 1.0 / 49, 1)):
        dloga = (np.log(af) - np.log(ai)) / nstep
        for loop in (0, nstep // 2, nstep - 1):
            li = loop * dloga + np.log(ai)
            lf = (loop + 1) * dloga + np.log(ai)
            rows.append((om, ox, li, lf, L.kick_loga(li, lf), L.drift_loga(li, lf)))
np.savez(os.path.join(HERE, "integrator_golden.npz"), rows=np.array(rows))
print(np.array(rows)[:3])
