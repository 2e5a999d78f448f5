"""Weighted minimax fit of the PM/FMM force-split factor used by the FP32 P2P kernel.

    g(u) = erfc(u) + 2u/sqrt(pi) exp(-u^2)  (src/fmm.c:845-848)  =  exp(-u^2) * Q(u),
    Q(u) = erfcx(u) + 2u/sqrt(pi)

Q(u) = 1 + u^2 R(u) (exact at u = 0: erfcx(u) = 1 - 2u/sqrt(pi) + u^2 - ...), and R is fitted by a
polynomial of degree DEG in u with weight u^2 exp(-u^2) (so the ABSOLUTE error of g is
minimised; Lawson iteration), then checked with float32 Horner evaluation.  Prints a C initialiser.
"""
import sys
import numpy as np
from scipy.special import erfc, erfcx
from numpy.polynomial import chebyshev as Ch, polynomial as Po

DEG = int(sys.argv[1]) if len(sys.argv) > 1 else 8
U = 6.0


def g(u):
    return erfc(u) + 2 * u / np.sqrt(np.pi) * np.exp(-u * u)


def Q(u):
    """R(u) = (erfcx(u) + 2u/sqrt(pi) - 1) / u^2, series near 0"""
    u = np.asarray(u, np.float64)
    big = (erfcx(u) + 2 * u / np.sqrt(np.pi) - 1) / np.where(u > 0, u * u, 1)
    a = 4 / (3 * np.sqrt(np.pi))
    small = 1 - a * u + u * u / 2 - (8 / (15 * np.sqrt(np.pi))) * u ** 3 + u ** 4 / 6
    return np.where(u < 1e-2, small, big)


M = 8000
u = np.linspace(0, U, M)
t = 2 * u / U - 1
V = Ch.chebvander(t, DEG)
w0 = u * u * np.exp(-u * u)
lam = np.ones(M) / M
for it in range(200):
    W = np.sqrt(lam) * w0
    c = np.linalg.lstsq(V * W[:, None], Q(u) * W, rcond=None)[0]
    e = np.abs((V @ c - Q(u)) * w0)
    lam = lam * (e + 1e-300)
    lam /= lam.sum()
# Chebyshev on [0, U] -> power basis in u
p = Ch.cheb2poly(c)                     # in t
pu = np.zeros(DEG + 1)
for k, a in enumerate(p):               # t = 2u/U - 1
    pu[:k + 1] += a * Po.polypow([-1.0, 2.0 / U], k)
uu = np.linspace(0, 9.0, 900001)
q64 = 1 + uu * uu * Po.polyval(uu, pu)
print("deg", DEG, "max abs err of g (float64 eval): %.3e" % np.abs(q64 * np.exp(-uu * uu) - g(uu)).max())
# float32 Horner, float32 exp2
c32 = pu.astype(np.float32)
u32 = uu.astype(np.float32)
acc = np.full_like(u32, c32[-1])
for k in range(DEG - 1, -1, -1):
    acc = (acc.astype(np.float64) * u32 + c32[k]).astype(np.float32)   # fma-like: one rounding
e32 = np.exp2((u32 * u32 * np.float32(-1.4426950408889634)).astype(np.float32)).astype(np.float32)
q32 = (acc.astype(np.float64) * (u32 * u32).astype(np.float32) + 1).astype(np.float32)
g32 = (q32 * e32).astype(np.float32)
print("max abs err of g (float32 Horner): %.3e" % np.abs(g32.astype(np.float64) - g(u32.astype(np.float64))).max())
print("static const float PN2_R%d[%d] = {" % (DEG, DEG + 1) + ", ".join("%.9ef" % x for x in pu) + "};")
