#!/usr/bin/env python3
"""One small evaluation for ncu captures: python tools/prof_case.py Nv n_r n_s [reps]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bfsm_b200 as B
inp = B.inputs
Nv, n_r, n_s = (int(a) for a in sys.argv[1:4])
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 1
gl = B.GaussLegendreQuadrature(n_r, 0.0, inp.R_SUPPORT); sd = B.SphericalDesign(n_s)
op = B.BoltzmannOperatorB200(gl, sd, Nv, Nv, Nv, 0.0, inp.B_GAMMA_MAXWELL, inp.L_DOMAIN)
op.initialize()
f = torch.from_numpy(inp.maxmix(Nv)).cuda().reshape(-1); q = torch.empty_like(f)
for _ in range(reps):
    op(q, f)
torch.cuda.synchronize()
print("ok", op.info())
