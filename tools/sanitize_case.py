#!/usr/bin/env python3
"""Small evaluations of every plane-kernel variant: 32^3 and 64^3 with the radix-32 plane kernel (register
and tensor-memory line), a 32^3 batch through the cell-group path, the 64^3 and 16^3 defaults.  Meant for
`compute-sanitizer --tool memcheck|racecheck python tools/sanitize_case.py` where the tool is available (it
is closed on the pool this round ran on); without it, a quick smoke run of all variants."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bfsm_b200 as B
inp = B.inputs
def run(Nv, n_r, n_s, cells=1, **opts):
    gl = B.GaussLegendreQuadrature(n_r, 0.0, inp.R_SUPPORT); sd = B.SphericalDesign(n_s)
    op = B.BoltzmannOperatorB200(gl, sd, Nv, Nv, Nv, 0.0, inp.B_GAMMA_MAXWELL, inp.L_DOMAIN, options=opts or None)
    op.initialize()
    f = torch.from_numpy(inp.maxmix(Nv)).cuda().reshape(-1).repeat(cells); q = torch.empty_like(f)
    op(q, f, n_cells=cells); torch.cuda.synchronize()
    print("ok", Nv, n_r, n_s, cells, opts, float(q.abs().max()), flush=True)
    op.close()
run(32, 2, 6)                       # r32, tensor-memory line (default)
run(32, 2, 6, plane_kernel=3)       # r32, register line
run(32, 2, 6, cells=3)              # cell groups
run(64, 1, 6, plane_kernel=4)       # r32 at 64^3 (cross-lane radix-2), tensor-memory line
run(64, 1, 6)                       # pipelined kernel + TMA x stage
run(16, 2, 6)
