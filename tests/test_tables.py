"""The committed function tables of the device kernels against scipy / mpmath (CPU; no GPU needed): the tables are what
replaces libm in the FP64 P2P kernel (pn2_gtab.h: g(u) of src/fmm.c:845-848) and in the M2L kernel (pn2_m2ltab.h: erfc(u) and
exp(-u^2)/sqrt(pi) of src/operator.c:294-307), evaluated here exactly as the device does (interval by round-to-nearest of
u / h - 1/2, Horner in float64), and the float32 polynomial of the FP32 P2P kernel (pn2_p2p.cuh, tools/fit_g.py)."""
import os
import re

import numpy as np
from scipy.special import erfc

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "photons-2.0_b200", "csrc")


def parse_header(path):
    txt = open(path).read()
    defs = {m.group(1): float(m.group(2)) for m in re.finditer(r"#define\s+(\w+)\s+([-+0-9.eE]+)\s*$", txt, re.M)}
    tabs = {}
    for m in re.finditer(r"const double (\w+)\[[^\]]*\]\[[^\]]*\]\s*=\s*\{(.*?)\n\};", txt, re.S):
        rows = re.findall(r"\{([^{}]*)\}", m.group(2))
        tabs[m.group(1)] = np.array([[float(x) for x in r.split(",")] for r in rows])
    return defs, tabs


def eval_table(C, invh, k_last, u):
    """C[j][k]: coefficient j of interval k; the device's evaluation (pn2_walk.cu p2p_interact_tab64, pn2_operators.cuh m2l_tab_eval)"""
    t = u * invh - 0.5
    k = np.rint(t)
    d = t - k
    kc = np.minimum(k.astype(int), k_last)
    v = C[-1, kc]
    for j in range(C.shape[0] - 2, -1, -1):
        v = v * d + C[j, kc]
    return v


def g_exact(u):
    return erfc(u) + 2 * u / np.sqrt(np.pi) * np.exp(-u * u)


def test_fp64_p2p_table():
    defs, tabs = parse_header(os.path.join(CSRC, "pn2_gtab.h"))
    C = tabs["PN2_GTAB"]
    K, deg = int(defs["PN2_GTAB_K"]), int(defs["PN2_GTAB_DEG"])
    assert C.shape == (deg + 1, K)
    u = np.linspace(0.0, 7.5, 200001)
    err = np.abs(eval_table(C, defs["PN2_GTAB_INVH"], K - 1, u) - np.where(u < 6.0, g_exact(u), 0.0))
    assert err.max() < 2e-10, err.max()                     # header: 1.9e-10; north_star's FP64 tolerance is 1e-6 on accelerations
    assert np.all(eval_table(C, defs["PN2_GTAB_INVH"], K - 1, np.array([6.01, 50.0, 1e4])) == 0.0)     # padding slots (at 1e4) / far pairs: exactly 0


def test_m2l_tables():
    defs, tabs = parse_header(os.path.join(CSRC, "pn2_m2ltab.h"))
    K, deg, kpad = int(defs["PN2_M2LTAB_K"]), int(defs["PN2_M2LTAB_DEG"]), int(defs["PN2_M2LTAB_KPAD"])
    u = np.linspace(0.0, 8.0, 200001)
    for name, fn in (("PN2_M2LTAB_E", erfc), ("PN2_M2LTAB_X", lambda x: np.exp(-x * x) / np.sqrt(np.pi))):
        C = tabs[name]
        assert C.shape == (deg + 1, kpad)
        err = np.abs(eval_table(C, defs["PN2_M2LTAB_INVH"], K - 1, u) - np.where(u < 6.4, fn(u), 0.0))
        assert err.max() < 5e-16, (name, err.max())         # the last bits of double (scipy's own erfc is good to ~1e-16)
        assert np.all(C[:, K - 1:] == 0.0)


def test_fp32_split_polynomial():
    """g(u) = exp(-u^2) (1 + u^2 R(u)) with the committed coefficients of the default degree, float32 Horner"""
    txt = open(os.path.join(CSRC, "pn2_p2p.cuh")).read()
    deg = int(re.search(r"#define PN2_RDEG (\d)", txt).group(1))
    m = re.search(r"#elif PN2_RDEG == %d\s*\n#define PN2_RCOEF \{([^}]*)\}" % deg, txt) or re.search(r"#if PN2_RDEG == %d\s*\n#define PN2_RCOEF \{([^}]*)\}" % deg, txt)
    c = np.array([float(x.strip().rstrip("f")) for x in m.group(1).split(",")], np.float32)
    assert len(c) == deg + 1
    u = np.linspace(0.0, 6.0, 60001).astype(np.float32)
    q = np.full_like(u, c[-1])
    for k in range(deg - 1, -1, -1):
        q = (q * u + c[k]).astype(np.float32)
    g = np.exp(-(u.astype(np.float64)) ** 2) * (1.0 + (u.astype(np.float64)) ** 2 * q.astype(np.float64))
    err = np.abs(g - g_exact(u.astype(np.float64))).max()
    bound = {8: 4e-7, 7: 1e-6, 6: 4e-6}[deg]
    assert err < bound, (deg, err)
