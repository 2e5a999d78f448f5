"""Host-side time-step factors of the reference's KDK integrator (src/initial.c:639-683): Simpson integrals of
dt/a (kick) and dt/a^2 (drift) over log a with 128 blocks, H(a) = 0.1 sqrt(OmegaM0 / a^3 + OmegaX0) in the reference's
units.  Restated operation by operation (same summation order) and checked bit for bit against the reference's own
kick_loga / drift_loga (tests/golden/make_integrator_golden.py)."""
import math

NBLOCK = 128


def _simpson(loga_i, loga_f, omega_m, omega_x, power):
    dloga = (loga_f - loga_i) / NBLOCK
    a_f, a_i = math.exp(loga_f), math.exp(loga_i)

    def term(z1):
        h = 0.1 * math.sqrt(omega_m * z1 * z1 * z1 + omega_x)
        return dloga * z1 / h if power == 1 else dloga * z1 * z1 / h

    z1 = 1.0 / a_i
    t = term(z1)
    for n in range(1, NBLOCK):
        z1 = 1.0 / math.exp(loga_i + dloga * n)
        h = 0.1 * math.sqrt(omega_m * z1 * z1 * z1 + omega_x)
        w = 2.0 * (1 + n % 2) * dloga
        t += (w * z1 / h) if power == 1 else (w * z1 * z1 / h)
    z1 = 1.0 / a_f
    t += term(z1)
    return t / 3.0


def kick_loga(loga_i, loga_f, omega_m, omega_x):
    """src/initial.c:639-660"""
    return _simpson(loga_i, loga_f, omega_m, omega_x, 1)


def drift_loga(loga_i, loga_f, omega_m, omega_x):
    """src/initial.c:662-683"""
    return _simpson(loga_i, loga_f, omega_m, omega_x, 2)


def step_factors(loop, dloga, a_init, omega_m, omega_x, grav_const):
    """(dkh, dd) of step `loop` (src/photoNs.c:141-150): dkh = 0.5 * kick * GravConst, dd = drift."""
    loga_i = loop * dloga + math.log(a_init)
    loga_f = (loop + 1) * dloga + math.log(a_init)
    dk = kick_loga(loga_i, loga_f, omega_m, omega_x)
    dd = drift_loga(loga_i, loga_f, omega_m, omega_x)
    return 0.5 * dk * grav_const, dd
