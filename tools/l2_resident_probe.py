#!/usr/bin/env python3
"""How fast are the two gain kernels when the hybrid scratch between them stays in L2?

    python tools/l2_resident_probe.py [reps]

One radius and designs of 12 ... 192 points give 6 ... 96 folded pairs, i.e. a hybrid scratch of
24 MiB ... 384 MiB at 64^3 written by the plane kernel and read back by the pencil kernel right
away (one launch each).  Below ~100 MiB the scratch is L2 resident (126 MB), above it streams
through HBM.  The per-pair times of the two kernel classes (CUDA events around each launch,
`bfsm_collide_profiled`) as a function of the scratch size say what a design that hands the hybrid
data over through an L2-resident ring could gain at best:

  * plane kernel: LSU bound, expected flat;
  * pencil kernel: HBM bound at 0.73 us/pair when streaming -- the L2-resident figure is its
    LSU/FP64 bound.

Small launches are dominated by ramp-up and work quantisation (6 pairs = 402 plane items on 148
CTAs), so read the trend, and compare the 16/24-pair rows with the 96-pair row.  Tuning aid, not a
benchmark.
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bfsm_b200 as B

inp = B.inputs
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 10
Nv = 64
f = torch.from_numpy(inp.maxmix(Nv)).cuda().reshape(-1)
q = torch.empty_like(f)
for n_s in (12, 32, 48, 70, 94, 192):
    gl = B.GaussLegendreQuadrature(1, 0.0, inp.R_SUPPORT)
    sd = B.SphericalDesign(n_s)
    op = B.BoltzmannOperatorB200(gl, sd, Nv, Nv, Nv, 0.0, inp.B_GAMMA_MAXWELL, inp.L_DOMAIN)
    op.initialize()
    info = op.info()
    pairs = info["pairs_local"]
    for _ in range(3):
        op(q, f)
    acc = {}
    for _ in range(reps):
        for k, (ms, _) in op.profile(q, f).items():
            acc[k] = acc.get(k, 0.0) + ms
    op.close()
    print(json.dumps({
        "pairs": pairs, "hybrid_MiB": pairs * 4, "plane_kernel": info["plane_kernel"],
        "plane_us_per_pair": round(1e3 * acc["plane_gain"] / reps / pairs, 3),
        "pencil_us_per_pair": round(1e3 * acc["pencil_gain"] / reps / pairs, 3),
        "nyquist_us_per_pair": round(1e3 * acc["nyquist"] / reps / pairs, 3),
        "class_ms": {k: round(v / reps, 4) for k, v in acc.items()}}), flush=True)
