#!/usr/bin/env python3
"""Is the cost of the reference CPU operator linear in the number of (r, sigma) pairs?

bench.py --impl reference times a SAMPLE of BASELINE config 4 (64^3, all 192 directions, a few radii:
the unmodified reference needs 96*N^3*P bytes = 154.6 GB for all 32 radii) and scales by
n_r / n_r_sample.  This script checks that extrapolation on the host cores of the box it runs on:

  * the unmodified reference operator (oracle/_ref) with 1, 2 and 3 radii,
  * the streaming C port (oracle/bfsm_oracle.c) with 1 radius and with the FULL 32-radius list.

    python tools/reference_scaling.py > profiles/r02_reference_scaling.json      (test-side tool: it runs the oracle)
"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import bfsm_b200 as B  # noqa: E402
from helpers import oracle_args, quadrature  # noqa: E402
from oracle import oracle as O  # noqa: E402

inp = B.inputs
Nv, n_s = 64, 192
cores = os.cpu_count() or 1
f = inp.maxmix(Nv)
out = {"Nv": Nv, "N_sigma": n_s, "cores": cores, "reference": [], "port": []}
if O.reference_available():
    for n_r in (1, 2, 3):
        op = O.ReferenceOperator(Nv, n_r, n_s, inp.GAMMA_MAXWELL, inp.B_GAMMA_MAXWELL, inp.L_DOMAIN,
                                 a=0.0, b=inp.R_SUPPORT, threads=cores)
        op(f, timed=True)
        t = min(op(f, timed=True)[1] for _ in range(2))
        op.close()
        out["reference"].append({"n_r": n_r, "pairs": n_r * n_s, "seconds": t, "seconds_per_pair": t / (n_r * n_s)})
po = O.PortOracle()
po.set_threads(cores)
for n_r in (1, 32):
    gl, sd = quadrature(n_r, n_s)
    t0 = time.perf_counter()
    po.collide((Nv,) * 3, *oracle_args(gl, sd), f)
    t = time.perf_counter() - t0
    out["port"].append({"n_r": n_r, "pairs": n_r * n_s, "seconds": t, "seconds_per_pair": t / (n_r * n_s)})
r = out["reference"]
if len(r) >= 2:
    out["reference_per_pair_spread"] = max(x["seconds_per_pair"] for x in r) / min(x["seconds_per_pair"] for x in r)
out["port_full_over_32x_one_radius"] = out["port"][1]["seconds"] / (32 * out["port"][0]["seconds"])
out["reading"] = ("seconds per pair is flat in the number of radii (fixed cost: one forward FFT + the loss "
                  "term, < 1 % of one radius), so evals/s of the full list = 1 / (t_sample * n_r / n_r_sample)")
print(json.dumps(out, indent=1))
