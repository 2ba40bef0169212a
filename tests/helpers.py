"""Shared helpers for the parity tests."""
import numpy as np

import bfsm_b200 as B

inp = B.inputs

#: tolerance stated by BASELINE.json's north_star: relative L-infinity error <= 1e-12 (fp64)
REL_LINF_TOL = 1e-12


def quadrature(n_r, n_s):
    gl = B.GaussLegendreQuadrature(n_r, 0.0, inp.R_SUPPORT)
    sd = B.SphericalDesign(n_s)
    return gl, sd


def oracle_args(gl, sd, gamma=inp.GAMMA_MAXWELL, b_gamma=inp.B_GAMMA_MAXWELL, L=inp.L_DOMAIN):
    return (gl.getNodes(), gl.getWeights(), sd.getx(), sd.gety(), sd.getz(), sd.getWeights(),
            gamma, b_gamma, L)


def make_input(kind, Nv, seed=0):
    if kind == "bkw":
        return inp.bkw(Nv)[0]
    if kind == "maxmix":
        return inp.maxmix(Nv, 1234 + seed)
    if kind == "noise":
        return inp.noise(Nv, 12345 + seed)
    raise ValueError(kind)


def rel_linf(a, b):
    a, b = np.asarray(a).ravel(), np.asarray(b).ravel()
    return float(np.abs(a - b).max() / np.abs(b).max())


def make_operator(Nv, n_r, n_s, **kw):
    """Nv: one size (cubic) or a (Nvx, Nvy, Nvz) tuple."""
    gl, sd = quadrature(n_r, n_s)
    nx, ny, nz = (Nv, Nv, Nv) if np.isscalar(Nv) else Nv
    op = B.BoltzmannOperatorB200(gl, sd, nx, ny, nz, inp.GAMMA_MAXWELL, inp.B_GAMMA_MAXWELL,
                                 inp.L_DOMAIN, **kw)
    op.initialize()
    return op, gl, sd
