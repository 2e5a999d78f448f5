"""The PM oracle's own consistency (CPU, golden files only): the C restatement of the reference's Fortran convolution
(src/conv.f90:128-247; oracle/ref_shim/ref_pm_harness.c, "parity unpinned": 2DECOMP&FFT is absent) is re-derived here
with numpy's FFT from the density mesh the UNMODIFIED src/partmesh.c deposited, and the force it leads to is checked
against physics: PM + short-range = Newtonian."""
import math

import numpy as np
from scipy.special import erfc

from conftest import load_golden


def green_numpy(density, box, rs):
    """conv.f90:172-220 with numpy: potential = IFFT(gf * FFT(density)), gf = pref exp(-k2 a) sinc^-4 / k2."""
    n = density.shape[0]
    pi_f = float(np.float32(3.1415926))                     # the reference's M_PI is a default-real literal (:143)
    ism = (2 * pi_f * rs / box) ** 2
    pref = box * box / (pi_f * n * n * n)
    l = np.fft.fftfreq(n, 1.0 / n)
    l[n // 2] = n // 2                                      # n > nhalf wraps, n == nhalf stays positive (:180-183)
    f = pi_f * l / n
    s = np.where(l == 0, 1.0, np.sin(f) / np.where(f == 0, 1.0, f))
    t = np.exp(-l * l * ism) / s ** 4
    k2 = l[:, None, None] ** 2 + l[None, :, None] ** 2 + l[None, None, :] ** 2
    gf = pref * t[:, None, None] * t[None, :, None] * t[None, None, :] / np.where(k2 == 0, 1.0, k2)
    gf[0, 0, 0] = pref
    return np.real(np.fft.ifftn(np.fft.fftn(density) * gf)) * n ** 3      # the reference's transforms are unnormalised


def test_convolution_restatement_against_numpy():
    for name in ("pm_demo_ns32.npz", "pm_small_ns24.npz"):
        g = load_golden(name)
        nside, box = int(g["nside"]), float(g["box"])
        pot = green_numpy(g["density"], box, 1.25 * box / nside)
        err = np.abs(pot - g["potential"]).max() / np.abs(g["potential"]).max()
        print(name, "potential: C restatement vs numpy", err)
        assert err < 1e-12


def test_deposit_conserves_mass():
    g = load_golden("pm_demo_ns32.npz")
    nside, box = int(g["nside"]), float(g["box"])
    total = g["density"].sum() * (box / nside) ** 3          # density = mass per cell volume (src/partmesh.c:168-178)
    assert abs(total / (float(g["mass"]) * len(g["acc_pm"])) - 1) < 1e-12


def test_pm_plus_short_range_is_newtonian():
    """Two particles 5 apart (2.6 rs): the mesh force of src/partmesh.c + the short-range factor of src/fmm.c:845-848
    add up to m / d^2 along the separation (mesh discretisation: a few per cent at this distance)."""
    g = load_golden("pm_pair_ns64.npz")
    pos, acc, box, nside = g["pos"], g["acc_pm"], float(g["box"]), int(g["nside"])
    rs = 1.25 * box / nside
    d = pos[1] - pos[0]
    r = math.sqrt((d ** 2).sum())
    u = r / (2 * rs)
    short = float(g["mass"]) / r ** 2 * (erfc(u) + 2 * u / math.sqrt(math.pi) * math.exp(-u * u))
    total = acc[0] + short * d / r
    newton = float(g["mass"]) / r ** 2 * d / r
    err = math.sqrt(((total - newton) ** 2).sum()) / math.sqrt((newton ** 2).sum())
    print("pair: |PM + short - Newton| / |Newton| =", err)
    assert err < 0.03
    assert np.abs(acc[0] + acc[1]).max() < 1e-12 * np.abs(acc).max() + 1e-15       # momentum
