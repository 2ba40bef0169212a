#!/usr/bin/env python3
"""Timing of the 32^3 BASELINE configurations (2, 3, 5) incl. batch-of-cells mode. Tuning aid."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bfsm_b200 as B
inp = B.inputs

def run(name, Nv, n_r, n_s, cells=1, reps=5):
    gl = B.GaussLegendreQuadrature(n_r, 0.0, inp.R_SUPPORT); sd = B.SphericalDesign(n_s)
    op = B.BoltzmannOperatorB200(gl, sd, Nv, Nv, Nv, 0.0, inp.B_GAMMA_MAXWELL, inp.L_DOMAIN)
    op.initialize()
    f = torch.from_numpy(inp.maxmix(Nv)).cuda().reshape(-1).repeat(cells); q = torch.empty_like(f)
    op(q, f, n_cells=cells); torch.cuda.synchronize()
    a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): op(q, f, n_cells=cells)
    b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b) / reps
    prof = op.profile(q[:Nv**3], f[:Nv**3]); info = op.info(); op.close()
    print(json.dumps({"cfg": name, "cells": cells, "ms_per_call": round(ms, 3), "cells_per_s": round(1e3 * cells / ms, 1),
                      "pairs": info["pairs_total"], "us_per_pair": round(1e3 * ms / cells / info["pairs_total"], 3),
                      "chunk": info["chunk_pairs"], "prof_ms": {k: round(v[0], 3) for k, v in prof.items()}}), flush=True)

run("cfg1", 16, 8, 6)
run("cfg2", 32, 16, 32)
run("cfg3", 32, 32, 48)
run("cfg5-cell", 32, 16, 94)
run("cfg5-16cells", 32, 16, 94, cells=16, reps=2)
