#!/usr/bin/env python3
"""Generate tests/golden/*.npz|json from the UNMODIFIED reference CPU operator.

Run in the build container (needs /root/reference, compiled by oracle/Makefile into
oracle/_ref/libbfsm_ref.so):    python tests/golden/make_golden.py

Outputs
  reference_q_vectors.npz   Q = BoltzmannOperator<FFTW_Backend>()(f) for seeded inputs at small
                            configurations, plus the quadrature nodes/weights the reference built
                            (GaussLegendre.hpp / SphericalDesign.cpp), single OpenMP thread.
  reference_q_64cubed.npz   the same at 64^3 (2 radii x 12-point design), stored as every second point
                            per axis plus per-plane sums of the full array (`--only-64` writes just
                            this file)
  bkw_known_answers.json    (a) the L1/L2/Linf errors PUBLISHED in the reference's Results/
                            (file:line cited per entry), (b) the same norms recomputed here with the
                            reference operator at BASELINE.json's (Nv, N_r, N_sigma) combinations.
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import bfsm_b200 as B  # noqa: E402
from oracle import oracle as O  # noqa: E402

inp = B.inputs
HERE = os.path.dirname(os.path.abspath(__file__))

VECTOR_CASES = [  # (Nv, N_r, N_sigma, input kind)
    (16, 8, 6, "bkw"), (16, 8, 6, "maxmix"), (16, 8, 6, "noise"),
    (16, 4, 12, "noise"), (32, 2, 12, "maxmix"), (32, 2, 12, "noise"),
]

PUBLISHED = [  # Results/maxwell_bkw_fftw_atomics.txt, 1-thread runs (N_gl = Nv)
    {"Nv": 32, "N_r": 32, "N_sigma": 12, "L1": 1.54029638e-03, "L2": 1.01189917e-04,
     "Linf": 4.25120273e-05, "source": "Results/maxwell_bkw_fftw_atomics.txt:19-21"},
    {"Nv": 32, "N_r": 32, "N_sigma": 32, "L1": 1.17016592e-03, "L2": 9.52999355e-05,
     "Linf": 4.41548044e-05, "source": "Results/maxwell_bkw_fftw_atomics.txt:371-373"},
    {"Nv": 64, "N_r": 64, "N_sigma": 12, "L1": 8.91494353e-11, "L2": 8.30921744e-12,
     "Linf": 3.06852243e-12, "source": "Results/maxwell_bkw_fftw_atomics.txt:195-197"},
    {"Nv": 64, "N_r": 64, "N_sigma": 32, "L1": 8.89089820e-11, "L2": 8.31383827e-12,
     "Linf": 3.07321573e-12, "source": "Results/maxwell_bkw_fftw_atomics.txt:547-549"},
]

RECOMPUTE = [(16, 8, 6), (16, 16, 6), (32, 16, 32), (32, 16, 48), (32, 32, 12)]

#: 64^3 (the size the pipelined plane kernel serves): every second point per axis (1/8 of the grid,
#: 256 KB per case) plus per-x-plane sums of Q and Q^2 of the FULL array -> reference_q_64cubed.npz
SUBSAMPLED_CASES = [(64, 2, 12, "maxmix"), (64, 2, 12, "noise"),
                    # the flagship design itself: ALL 192 directions of ss019.192 with one radius (the
                    # reference's six batch arrays need 96*64^3*192 = 4.8 GB here, 154.6 GB with 32 radii)
                    (64, 1, 192, "maxmix"), (64, 1, 192, "noise")]
STRIDE = 2


def make_input(kind, Nv):
    if kind == "bkw":
        return inp.bkw(Nv)[0]
    if kind == "maxmix":
        return inp.maxmix(Nv, 1234)
    return inp.noise(Nv, 12345)


def main():
    out = {}
    for Nv, n_r, n_s, kind in ([] if "--only-64" in sys.argv else VECTOR_CASES):
        op = O.ReferenceOperator(Nv, n_r, n_s, inp.GAMMA_MAXWELL, inp.B_GAMMA_MAXWELL, inp.L_DOMAIN,
                                 a=0.0, b=inp.R_SUPPORT, threads=1)
        f = make_input(kind, Nv)
        key = f"Nv{Nv}_r{n_r}_s{n_s}_{kind}"
        out[key + "_Q"] = op(f)
        rho, w_r, sx, sy, sz, sw = op.quadrature()
        out[f"Nv{Nv}_r{n_r}_s{n_s}_rho"] = rho
        out[f"Nv{Nv}_r{n_r}_s{n_s}_wr"] = w_r
        op.close()
        print("vector", key, float(np.abs(out[key + "_Q"]).max()))
    if out:
        np.savez_compressed(os.path.join(HERE, "reference_q_vectors.npz"), **out)

    sub = {"stride": np.array(STRIDE)}
    for Nv, n_r, n_s, kind in SUBSAMPLED_CASES:
        op = O.ReferenceOperator(Nv, n_r, n_s, inp.GAMMA_MAXWELL, inp.B_GAMMA_MAXWELL, inp.L_DOMAIN,
                                 a=0.0, b=inp.R_SUPPORT, threads=(1 if n_s <= 12 else 0))
        Q = op(make_input(kind, Nv)).reshape(Nv, Nv, Nv)
        op.close()
        key = f"Nv{Nv}_r{n_r}_s{n_s}_{kind}"
        sub[key + "_Qsub"] = Q[::STRIDE, ::STRIDE, ::STRIDE].copy()
        sub[key + "_plane_sum"] = Q.sum(axis=(1, 2))
        sub[key + "_plane_sumsq"] = (Q * Q).sum(axis=(1, 2))
        sub[key + "_max"] = np.array(np.abs(Q).max())
        print("subsampled", key, float(np.abs(Q).max()))
    np.savez_compressed(os.path.join(HERE, "reference_q_64cubed.npz"), **sub)
    if "--only-64" in sys.argv:
        return

    recomputed = []
    for Nv, n_r, n_s in RECOMPUTE:
        op = O.ReferenceOperator(Nv, n_r, n_s, inp.GAMMA_MAXWELL, inp.B_GAMMA_MAXWELL, inp.L_DOMAIN,
                                 a=0.0, b=inp.R_SUPPORT, threads=1)
        f, Q_exact = inp.bkw(Nv)
        l1, l2, linf = inp.error_norms(op(f), Q_exact, Nv)
        op.close()
        recomputed.append({"Nv": Nv, "N_r": n_r, "N_sigma": n_s, "L1": l1, "L2": l2, "Linf": linf,
                           "source": "reference operator (oracle/_ref), 1 thread, this script"})
        print("bkw", Nv, n_r, n_s, l1, l2, linf)
    with open(os.path.join(HERE, "bkw_known_answers.json"), "w") as fh:
        json.dump({"published": PUBLISHED, "recomputed": recomputed}, fh, indent=1)


if __name__ == "__main__":
    main()
