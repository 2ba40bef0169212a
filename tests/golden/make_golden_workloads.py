#!/usr/bin/env python3
"""Golden Q vectors for bench.py's workloads other than BASELINE config 4 (which has its own file,
port_q_cfg4.npz): written by the C PORT of the reference algorithm (oracle/bfsm_oracle.c, pinned to
the unmodified reference operator by tests/test_oracle.py), input maxmix(seed).  bench.py compares
the Q it has just timed with these after the timed loop (`parity_rel_linf` on its JSON line) -- it may
not call the oracle itself.  Stored like the other 64^3 goldens: every second point per axis plus
per-x-plane sums of Q and Q^2 of the full grid.

    python tests/golden/make_golden_workloads.py      ->  tests/golden/port_q_workloads.npz
"""
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import bfsm_b200 as B  # noqa: E402
from helpers import oracle_args, quadrature  # noqa: E402
from oracle import oracle as O  # noqa: E402

inp = B.inputs
STRIDE = 2
#: (Nv, N_r, N_sigma, seeds)
CASES = [(16, 8, 6, [1234]), (32, 16, 32, [1234]), (32, 32, 48, [1234]),
         (32, 16, 94, list(range(1234, 1242)))]   # cfg 5: the 8 distinct cells bench.py tiles

out = {"stride": np.array(STRIDE)}
po = O.PortOracle()
for Nv, n_r, n_s, seeds in CASES:
    gl, sd = quadrature(n_r, n_s)
    for seed in seeds:
        t = time.time()
        Q = np.asarray(po.collide((Nv,) * 3, *oracle_args(gl, sd), inp.maxmix(Nv, seed))).reshape(Nv, Nv, Nv)
        key = f"Nv{Nv}_r{n_r}_s{n_s}_maxmix{seed}"
        out[key + "_Qsub"] = Q[::STRIDE, ::STRIDE, ::STRIDE].copy()
        out[key + "_plane_sum"] = Q.sum(axis=(1, 2))
        out[key + "_plane_sumsq"] = (Q * Q).sum(axis=(1, 2))
        out[key + "_max"] = np.array(np.abs(Q).max())
        print(key, "max|Q| =", float(np.abs(Q).max()), "in", round(time.time() - t, 1), "s", flush=True)
np.savez_compressed(os.path.join(HERE, "port_q_workloads.npz"), **out)
