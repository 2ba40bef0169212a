#!/usr/bin/env python3
"""Golden Q at BASELINE config 4 (64^3, 32 Gauss-Legendre radii x 192-point design) from the C PORT.

The unmodified reference operator cannot run this configuration here (its six batch arrays need
154.6 GB, FFTWBoltzmannOperator.cpp:30-37); the streaming C restatement (oracle/bfsm_oracle.c) can,
in ~1 minute on 8 cores, and it is pinned to the reference at 64^3 to 2e-15
(tests/test_oracle.py::test_port_matches_reference_at_64_cubed).  Stored like reference_q_64cubed.npz:
every second point per axis + per-x-plane sums of Q and Q^2 of the full grid.

    python tests/golden/make_golden_cfg4.py        ->  tests/golden/port_q_cfg4.npz
"""
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import make_input, oracle_args, quadrature  # noqa: E402
from oracle import oracle as O  # noqa: E402

Nv, n_r, n_s, STRIDE = 64, 32, 192, 2
out = {"stride": np.array(STRIDE)}
gl, sd = quadrature(n_r, n_s)
po = O.PortOracle()
for kind in ("maxmix", "noise"):
    t = time.time()
    Q = np.asarray(po.collide((Nv,) * 3, *oracle_args(gl, sd), make_input(kind, Nv))).reshape(Nv, Nv, Nv)
    key = f"Nv{Nv}_r{n_r}_s{n_s}_{kind}"
    out[key + "_Qsub"] = Q[::STRIDE, ::STRIDE, ::STRIDE].copy()
    out[key + "_plane_sum"] = Q.sum(axis=(1, 2))
    out[key + "_plane_sumsq"] = (Q * Q).sum(axis=(1, 2))
    out[key + "_max"] = np.array(np.abs(Q).max())
    print(key, "max|Q| =", float(np.abs(Q).max()), "in", round(time.time() - t, 1), "s")
np.savez_compressed(os.path.join(HERE, "port_q_cfg4.npz"), **out)
