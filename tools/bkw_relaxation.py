#!/usr/bin/env python3
"""BASELINE config 3: BKW relaxation from t0 to t_final on the GPU, error vs the exact solution.
    python tools/bkw_relaxation.py [--Nv 32] [--Nr 32] [--Ns 48] [--t0 5.5] [--tfinal 6.5] [--dt 0.05]"""
import argparse, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bfsm_b200 as B
inp = B.inputs
I = B.submodule("integrate")
ap = argparse.ArgumentParser()
ap.add_argument("--Nv", type=int, default=32); ap.add_argument("--Nr", type=int, default=32)
ap.add_argument("--Ns", type=int, default=48); ap.add_argument("--t0", type=float, default=5.5)
ap.add_argument("--tfinal", type=float, default=6.5); ap.add_argument("--dt", type=float, default=0.05)
a = ap.parse_args()
gl = B.GaussLegendreQuadrature(a.Nr, 0.0, inp.R_SUPPORT); sd = B.SphericalDesign(a.Ns)
op = B.BoltzmannOperatorB200(gl, sd, a.Nv, a.Nv, a.Nv, 0.0, inp.B_GAMMA_MAXWELL, inp.L_DOMAIN); op.initialize()
f = torch.from_numpy(I.bkw_exact(a.Nv, a.t0)).cuda().reshape(-1)
torch.cuda.synchronize(); t = time.perf_counter()
f, steps, evals = I.rk4_torch(op, f, a.t0, a.tfinal, a.dt)
torch.cuda.synchronize(); wall = time.perf_counter() - t
exact = I.bkw_exact(a.Nv, a.tfinal)
l1, l2, linf = inp.error_norms(f.cpu().numpy(), exact, a.Nv)
_, dv = inp.velocity_axis(a.Nv)
print(json.dumps({"Nv": a.Nv, "N_r": a.Nr, "N_sigma": a.Ns, "t0": a.t0, "t_final": a.tfinal, "steps": steps,
                  "operator_evals": evals, "wall_s": round(wall, 4), "evals_per_s": round(evals / wall, 1),
                  "L1": l1, "L2": l2, "Linf": linf, "rel_Linf": linf / float(np.abs(exact).max()),
                  "mass": float(f.sum().item() * dv ** 3)}))
