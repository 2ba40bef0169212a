#!/usr/bin/env python3
"""Env-knob sweep for the 32^3 single-cell configuration (cfg 5 cell: 16 radii x 94 directions)."""
import os, sys, json, itertools
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bfsm_b200 as B
inp = B.inputs
def run(env, Nv=32, n_r=16, n_s=94, reps=20):
    for k in ("BFSM_NYQ_GROUPS", "BFSM_CHUNK_PAIRS", "BFSM_GAIN_CTAS", "BFSM_SIDE_STREAM"):
        os.environ.pop(k, None)
    os.environ.update({k: str(v) for k, v in env.items()})
    gl = B.GaussLegendreQuadrature(n_r, 0.0, inp.R_SUPPORT); sd = B.SphericalDesign(n_s)
    op = B.BoltzmannOperatorB200(gl, sd, Nv, Nv, Nv, 0.0, inp.B_GAMMA_MAXWELL, inp.L_DOMAIN); op.initialize()
    f = torch.from_numpy(inp.maxmix(Nv)).cuda().reshape(-1); q = torch.empty_like(f)
    for _ in range(3): op(q, f)
    torch.cuda.synchronize()
    a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): op(q, f)
    b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b) / reps
    op.close()
    print(json.dumps({"env": env, "ms_per_cell": round(ms, 4), "cells_per_s": round(1e3 / ms, 1)}), flush=True)
for gy in (4, 8, 12, 24):
    run({"BFSM_NYQ_GROUPS": gy})
for chunk in (128, 376, 752):
    run({"BFSM_CHUNK_PAIRS": chunk})
for ctas in (148, 296, 444):
    run({"BFSM_GAIN_CTAS": ctas})
run({"BFSM_SIDE_STREAM": 0})
run({"BFSM_NYQ_GROUPS": 8, "BFSM_CHUNK_PAIRS": 752})
