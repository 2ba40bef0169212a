"""Time integration of the space-homogeneous Boltzmann equation df/dt = Q(f,f) on top of the
collision operator (BASELINE.json config 3: "BKW time integration to t_final ... error vs exact BKW
solution").  The reference has no time integrator -- its drivers evaluate Q once at t = 6.5
(maxwell_bkw_fftw.cpp:74-99) -- so this is the caller-side extension SURVEY.md section 8(f) ranks
first.  Classical RK4 with all stages resident where the operator's buffers live (device tensors
for the CUDA operator, NumPy arrays for a CPU stand-in); only `collide(Q, f)` is backend specific.
"""
import math

import numpy as np

from .inputs import velocity_axis
from .quadratures import pi


def bkw_exact(Nv, t):
    """Exact BKW solution f(t, v) on the reference grid (maxwell_bkw_fftw.cpp:74-91); needs
    t > 6 ln(5/2) ~ 5.498 for positivity."""
    v, _ = velocity_axis(Nv)
    K = 1 - math.exp(-t / 6)
    r_sq = v[:, None, None] ** 2 + v[None, :, None] ** 2 + v[None, None, :] ** 2
    f = np.exp(-r_sq / (2 * K)) * ((5 * K - 3) / K + (1 - K) / K ** 2 * r_sq) / (2 * (2 * pi * K) ** 1.5)
    return np.ascontiguousarray(f)


def rk4(collide, f, t0, t_final, dt, new_like, axpy):
    """Integrate f from t0 to t_final with classical RK4.

    collide(Q, f) -> Q : the collision operator (Q written in place)
    new_like(f)        : allocate a buffer like f
    axpy(out, a, x, y) : out <- a*x + y (elementwise, out may alias y)
    Returns (f(t_final), number of steps, number of operator evaluations)."""
    n_steps = max(1, int(round((t_final - t0) / dt)))
    h = (t_final - t0) / n_steps
    k = [new_like(f) for _ in range(4)]
    tmp = new_like(f)
    for _ in range(n_steps):
        collide(k[0], f)
        axpy(tmp, 0.5 * h, k[0], f)
        collide(k[1], tmp)
        axpy(tmp, 0.5 * h, k[1], f)
        collide(k[2], tmp)
        axpy(tmp, h, k[2], f)
        collide(k[3], tmp)
        axpy(f, h / 6, k[0], f)
        axpy(f, h / 3, k[1], f)
        axpy(f, h / 3, k[2], f)
        axpy(f, h / 6, k[3], f)
    return f, n_steps, 4 * n_steps


def rk4_torch(op, f_dev, t0, t_final, dt):
    """RK4 with a BoltzmannOperatorB200 and a flat float64 CUDA tensor (stages stay on the device)."""
    import torch

    def axpy(out, a, x, y):
        torch.add(y, x, alpha=a, out=out)

    return rk4(lambda Q, f: op(Q, f), f_dev, t0, t_final, dt, torch.empty_like, axpy)


def rk4_numpy(collide, f, t0, t_final, dt):
    """Same integrator on the host; `collide(f) -> Q` e.g. the CPU oracle."""
    def coll(Q, x):
        Q[...] = collide(x)
        return Q

    def axpy(out, a, x, y):
        np.add(y, a * x, out=out)

    return rk4(coll, f, t0, t_final, dt, np.empty_like, axpy)
