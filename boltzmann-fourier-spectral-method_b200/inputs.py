"""Synthetic distributions on the reference's velocity grid.

Grid and constants are the drivers' (maxwell_bkw_fftw.cpp:54-71): S=5, R=2S, L=(3+sqrt 2)/2*S,
dv=2L/Nv, v_i=-L+dv/2+i*dv, Maxwell molecules gamma=0, b_gamma=1/(4 pi).
"""
import math

import numpy as np

from .quadratures import pi

S_SUPPORT = 5.0
R_SUPPORT = 2 * S_SUPPORT
L_DOMAIN = ((3 + math.sqrt(2)) / 2) * S_SUPPORT
GAMMA_MAXWELL = 0.0
B_GAMMA_MAXWELL = 1 / (4 * pi)


def velocity_axis(Nv):
    dv = 2 * L_DOMAIN / Nv
    return -L_DOMAIN + dv / 2 + np.arange(Nv) * dv, dv


def bkw(Nv, t=6.5):
    """BKW solution f(t) and its exact time derivative Q = df/dt (maxwell_bkw_fftw.cpp:74-99)."""
    v, _ = velocity_axis(Nv)
    K = 1 - math.exp(-t / 6)
    dK = math.exp(-t / 6) / 6
    r_sq = (v[:, None, None] ** 2 + v[None, :, None] ** 2 + v[None, None, :] ** 2)
    norm = 1 / (2 * (2 * pi * K) ** 1.5)
    f = np.exp(-r_sq / (2 * K)) * ((5 * K - 3) / K + (1 - K) / K ** 2 * r_sq) * norm
    Q = (-3 / (2 * K) + r_sq / (2 * K ** 2)) * f
    Q = Q + norm * np.exp(-r_sq / (2 * K)) * (3 / K ** 2 + (K - 2) / K ** 3 * r_sq)
    Q = Q * dK
    return np.ascontiguousarray(f), np.ascontiguousarray(Q)


def _shape(Nv):
    """Nv: one size (cubic grid) or a (Nvx, Nvy, Nvz) tuple."""
    return (Nv, Nv, Nv) if np.isscalar(Nv) else tuple(int(n) for n in Nv)


def maxmix(Nv, seed=1234):
    """Sum of four Maxwellians with seeded random density / temperature / drift; `Nv` may be a
    (Nvx, Nvy, Nvz) tuple (every axis spans the same [-L, L])."""
    rng = np.random.default_rng(seed)
    nx, ny, nz = _shape(Nv)
    vx, vy, vz = velocity_axis(nx)[0], velocity_axis(ny)[0], velocity_axis(nz)[0]
    f = np.zeros((nx, ny, nz))
    for _ in range(4):
        rho = rng.uniform(0.5, 1.5)
        T = rng.uniform(0.5, 1.5)
        u = rng.uniform(-2, 2, size=3)
        r_sq = ((vx[:, None, None] - u[0]) ** 2 + (vy[None, :, None] - u[1]) ** 2
                + (vz[None, None, :] - u[2]) ** 2)
        f += rho * (2 * pi * T) ** -1.5 * np.exp(-r_sq / (2 * T))
    return np.ascontiguousarray(f)


def noise(Nv, seed=12345):
    """i.i.d. U(0,1) samples: not band limited, exercises the Nyquist planes (parity only)."""
    rng = np.random.default_rng(seed)
    return np.ascontiguousarray(rng.random(_shape(Nv)))


def error_norms(Q, Q_exact, Nv):
    """L1, L2 (scaled by dv^3) and Linf errors as printed by maxwell_bkw_fftw.cpp:145-166."""
    _, dv = velocity_axis(Nv)
    d = np.abs(np.asarray(Q).ravel() - np.asarray(Q_exact).ravel())
    return float(d.sum() * dv ** 3), float(math.sqrt((d ** 2).sum() * dv ** 3)), float(d.max())
