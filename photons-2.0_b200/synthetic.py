"""Synthetic LambdaCDM-like particle sets (SURVEY.md 8d iii): a regular grid displaced by a Zel'dovich field
with a BBKS-shaped power spectrum, periodic box.  Generated with torch FFTs on whatever device is given
(the GPU for the bench, the CPU for the small reference-arm samples); torch is plumbing here.

Units follow the reference's demo (demo/lcdm_g2.run, demo/ic_lcdm.gdt2): BOX = 100000 kpc/h,
mass = Omega_m * 3 * 0.01 / (8 pi G) * BOX^3 / N  (src/initial.c:596), G = 43007.1.
"""
import math

import torch

BOX = 100000.0
OMEGA_M = 0.25
GRAV = 43007.105732


def particle_mass(n_total, box=BOX, omega_m=OMEGA_M):
    return omega_m * 3.0 * 0.01 / (8.0 * math.pi * GRAV) * box ** 3 / float(n_total)


def lcdm_like(nside, box=BOX, disp_rms=0.3, seed=12345, device="cpu", gamma=0.175, ns=0.96):
    """nside^3 particles: grid + displacement field psi = grad(inverse laplacian(delta)), delta Gaussian with
    P(k) ~ k^ns T_BBKS(k)^2; psi is scaled to an rms of `disp_rms` grid spacings per dimension
    (0.3: z~49-like; 2: clustered).  Returns float64 positions (N, 3) in [0, box) on `device`."""
    dev = torch.device(device)
    g = torch.Generator(device=dev)
    g.manual_seed(seed)
    n = nside
    white = torch.randn((n, n, n), generator=g, device=dev, dtype=torch.float32)
    wk = torch.fft.rfftn(white)
    del white
    kf = 2.0 * math.pi / (box / 1000.0)            # fundamental mode in h/Mpc (box in kpc/h)
    kx = torch.fft.fftfreq(n, d=1.0 / n, device=dev).to(torch.float32) * kf
    kz = torch.fft.rfftfreq(n, d=1.0 / n, device=dev).to(torch.float32) * kf
    k2 = kx[:, None, None] ** 2 + kx[None, :, None] ** 2 + kz[None, None, :] ** 2
    k = torch.sqrt(k2)
    qv = k / gamma
    T = torch.log(1.0 + 2.34 * qv) / (2.34 * qv + 1e-30) * (1.0 + 3.89 * qv + (16.1 * qv) ** 2 + (5.46 * qv) ** 3 + (6.71 * qv) ** 4) ** -0.25
    amp = torch.sqrt(k.clamp_min(1e-30) ** ns) * T
    amp[0, 0, 0] = 0.0
    k2[0, 0, 0] = 1.0
    dk = wk * amp / k2                               # -> psi_k = i k delta_k / k^2
    del wk, amp, T, qv, k
    spacing = box / n
    pos = torch.empty((n * n * n, 3), dtype=torch.float64, device=dev)
    grid1 = (torch.arange(n, device=dev, dtype=torch.float64) + 0.5) * spacing
    comps = [kx[:, None, None], kx[None, :, None], kz[None, None, :]]
    psis = []
    for d in range(3):
        psi = torch.fft.irfftn(1j * comps[d] * dk, s=(n, n, n))
        psis.append(psi)
    del dk
    rms = math.sqrt(sum(float((p.double() ** 2).mean()) for p in psis) / 3.0)
    scale = disp_rms * spacing / rms
    shape = [(n, 1, 1), (1, n, 1), (1, 1, n)]
    for d in range(3):
        x = grid1.reshape(shape[d]) + psis[d].double() * scale
        pos[:, d] = torch.remainder(x, box).reshape(-1)
        psis[d] = None
    # remainder can return box for tiny negative inputs
    pos = torch.where(pos >= box, pos - box, pos)
    return pos


def poisson(n_total, box=BOX, seed=378412, device="cpu"):
    """Uniform random (Poisson) particles: the reference's own `-2` initial condition (ic_uniform, src/initial.c:558-618:
    ran3 with seed 378412 + rank per coordinate).  Same distribution, torch's generator instead of ran3."""
    dev = torch.device(device)
    g = torch.Generator(device=dev)
    g.manual_seed(seed)
    pos = torch.rand((n_total, 3), generator=g, device=dev, dtype=torch.float64) * box
    return torch.where(pos >= box, pos - box, pos)
