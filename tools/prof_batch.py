#!/usr/bin/env python3
"""One batched call (n_cells cells of Nv^3) for ncu launch lists: python tools/prof_batch.py Nv n_r n_s cells"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bfsm_b200 as B
inp = B.inputs
Nv, n_r, n_s, cells = (int(a) for a in sys.argv[1:5])
gl = B.GaussLegendreQuadrature(n_r, 0.0, inp.R_SUPPORT); sd = B.SphericalDesign(n_s)
op = B.BoltzmannOperatorB200(gl, sd, Nv, Nv, Nv, 0.0, inp.B_GAMMA_MAXWELL, inp.L_DOMAIN)
op.initialize()
f = torch.from_numpy(inp.maxmix(Nv)).cuda().reshape(-1).repeat(cells); q = torch.empty_like(f)
op(q, f, n_cells=cells); torch.cuda.synchronize()
a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
a.record(); op(q, f, n_cells=cells); b.record(); torch.cuda.synchronize()
print("ok", op.info(), "ms_per_cell", a.elapsed_time(b) / cells)
