"""Synthetic NBODY6++ snapshots for tests and bench (numpy, seeded).

Restates, not copies, the reference's initial-condition recipes:

* Plummer positions/velocities: Aarseth, Henon & Wielen sampling as in src/Main/setup.F:62-107
  (radius from the cumulative mass with rejection of r > 10, von-Neumann rejection of
  q^2 (1-q^2)^3.5 for the speed, scaling SX = 3*pi/16, SV = sqrt(M/SX), c.m. correction).
* Kroupa, Tout & Gilmore 1993 masses: src/Main/imf2.f:41
  m = 0.08 + (0.19 X^1.55 + 0.05 X^0.6) / (1 - X)^0.58, accepted inside [mlo, mhi], then
  normalised to total mass 1 (scale.F).
* Neighbour radius and regular step as FPOLY0 sets them (src/Main/fpoly0.F:51-56):
  RS_i = RS0 sqrt(1+r^2), STEPR_i = min(SMAX/8 sqrt(1+r^2), SMAX).

The RNG is numpy's PCG64, not the reference's ran2: snapshots are statistically, not bitwise,
the reference's.
"""
from __future__ import annotations

import numpy as np


def kroupa_masses(n: int, rng: np.random.Generator, mlo: float = 0.08, mhi: float = 100.0) -> np.ndarray:
    out = np.empty(n)
    filled = 0
    while filled < n:
        x = rng.random(int((n - filled) * 1.1) + 16)
        zm = 0.08 + (0.19 * x**1.55 + 0.05 * x**0.6) / (1.0 - x) ** 0.58
        zm = zm[(zm >= mlo) & (zm <= mhi)]
        k = min(zm.size, n - filled)
        out[filled:filled + k] = zm[:k]
        filled += k
    return out


def plummer(n: int, seed: int = 1, imf: str = "equal"):
    """Returns m[n], x[n,3], v[n,3] (float64) in N-body units (total mass 1, virial equilibrium)."""
    rng = np.random.default_rng(seed)
    m = np.full(n, 1.0 / n) if imf == "equal" else kroupa_masses(n, rng)
    m = m / m.sum()
    # radii (setup.F:63-68)
    r = np.empty(n)
    filled = 0
    while filled < n:
        a1 = rng.random(int((n - filled) * 1.05) + 16)
        a1 = a1[a1 >= 1.0e-10]
        ri = (a1 ** (-2.0 / 3.0) - 1.0) ** (-0.5)
        ri = ri[ri <= 10.0]
        k = min(ri.size, n - filled)
        r[filled:filled + k] = ri[:k]
        filled += k
    a2, a3 = rng.random(n), rng.random(n)
    x = np.empty((n, 3))
    x[:, 2] = (1.0 - 2.0 * a2) * r
    rxy = np.sqrt(np.maximum(r * r - x[:, 2] ** 2, 0.0))
    x[:, 0] = rxy * np.cos(2 * np.pi * a3)
    x[:, 1] = rxy * np.sin(2 * np.pi * a3)
    # speeds (setup.F:74-78)
    q = np.empty(n)
    filled = 0
    while filled < n:
        a4 = rng.random(int((n - filled) * 10.5) + 64)
        a5 = rng.random(a4.size)
        ok = 0.1 * a5 <= a4 * a4 * (1.0 - a4 * a4) ** 3.5
        a4 = a4[ok]
        k = min(a4.size, n - filled)
        q[filled:filled + k] = a4[:k]
        filled += k
    vmag = q * np.sqrt(2.0) / (1.0 + r * r) ** 0.25
    a6, a7 = rng.random(n), rng.random(n)
    v = np.empty((n, 3))
    v[:, 2] = (1.0 - 2.0 * a6) * vmag
    vxy = np.sqrt(np.maximum(vmag * vmag - v[:, 2] ** 2, 0.0))
    v[:, 0] = vxy * np.cos(2 * np.pi * a7)
    v[:, 1] = vxy * np.sin(2 * np.pi * a7)
    # c.m. frame and scaling (setup.F:94-104)
    zmass = m.sum()
    x -= (m[:, None] * x).sum(0) / zmass
    v -= (m[:, None] * v).sum(0) / zmass
    sx = 1.5 * 2 * np.pi / 16.0
    x *= sx
    v *= np.sqrt(zmass / sx)
    return m, x, v


def radii(x: np.ndarray, m: np.ndarray, rs0: float, smax: float = 0.125, m_flag: int = 0):
    """h2 = RS^2 (divided by the mean mass when m_flag=1, util_gpu.F:46-51) and dtr = STEPR."""
    ri2 = (x * x).sum(1)
    rs = rs0 * np.sqrt(1.0 + ri2)
    dtr = np.minimum(smax / 8.0 * np.sqrt(1.0 + ri2), smax)
    h2 = rs * rs
    if m_flag:
        h2 = h2 / m.mean()
    return h2, dtr


def rs0_for_nnb(n: int, nnb: float) -> float:
    """RS0 giving about `nnb` neighbours near the centre of a Plummer sphere in N-body units.

    Central density rho0 = 3 N / (4 pi a^3) with a = 3 pi / 16; nnb = 4/3 pi RS^3 rho0.
    """
    a = 3.0 * np.pi / 16.0
    return a * (nnb / n) ** (1.0 / 3.0)


def radii_nnb(x: np.ndarray, m: np.ndarray, nnb: float, smax: float = 0.125, m_flag: int = 0):
    """Neighbour spheres holding about `nnb` members everywhere: RS_i from the local Plummer density,
    n(r) = 3N/(4 pi a^3) (1 + r^2/a^2)^(-5/2), a = 3 pi/16 -- the state the integrator's RS control
    (NNBOPT, regcor_gpu.F:623-760) converges to, as opposed to FPOLY0's first guess RS0 sqrt(1+r^2).
    dtr as in fpoly0.F:53-56."""
    n = x.shape[0]
    a = 3.0 * np.pi / 16.0
    ri2 = (x * x).sum(1)
    rs = a * (nnb / n) ** (1.0 / 3.0) * (1.0 + ri2 / (a * a)) ** (5.0 / 6.0)
    # the integrator never lets a halo sphere reach back into the core (RS limits in regcor_gpu.F); without a bound the
    # r^(5/3) growth does exactly that at small N.  Not binding for r <= 10 at N = 10^6.
    rs = np.minimum(rs, 0.4 * np.sqrt(ri2 + a * a))
    dtr = np.minimum(smax / 8.0 * np.sqrt(1.0 + ri2), smax)
    h2 = rs * rs
    if m_flag:
        # criterion r^2 < h2 * m_j: a sphere of radius rs sqrt(m_j / <m>) per j, so the count scales with
        # <(m/<m>)^1.5>; RS is taken smaller by that factor^(1/3) to keep ~nnb members with an IMF
        w = float(((m / m.mean()) ** 1.5).mean())
        h2 = h2 / w ** (2.0 / 3.0) / m.mean()
    return h2, dtr
