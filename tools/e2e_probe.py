#!/usr/bin/env python3
"""Device-timed loop vs pipelined host-buffer loop on ONE GPU for a short step (64^3, n_r radii x 192):
does the end-to-end residual of the 8-rank runs (0.79 ms steps) come from the step length alone?
    python tools/e2e_probe.py [n_r] [steps]"""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bfsm_b200 as B
inp = B.inputs
capi = B.submodule("_capi")
n_r = int(sys.argv[1]) if len(sys.argv) > 1 else 4
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 200
Nv, n_s = 64, 192
gl = B.GaussLegendreQuadrature(n_r, 0.0, inp.R_SUPPORT); sd = B.SphericalDesign(n_s)
op = B.BoltzmannOperatorB200(gl, sd, Nv, Nv, Nv, 0.0, inp.B_GAMMA_MAXWELL, inp.L_DOMAIN)
op.initialize()
f_host = torch.from_numpy(inp.maxmix(Nv).reshape(-1).copy()).pin_memory()
depth = capi.BFSM_HOST_PIPE_DEPTH
q_host = [torch.empty(Nv ** 3, dtype=torch.float64).pin_memory() for _ in range(depth)]
f = f_host.cuda(); q = torch.empty_like(f)
for _ in range(5): op(q, f)
torch.cuda.synchronize()
a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(steps): op(q, f)
b.record(); torch.cuda.synchronize()
dev_ms = a.elapsed_time(b) / steps
for k in range(depth): op.submit_host(q_host[k], f_host)
op.flush_host()
t0 = time.perf_counter()
for k in range(steps): op.submit_host(q_host[k % depth], f_host)
op.flush_host()
e2e_ms = 1e3 * (time.perf_counter() - t0) / steps
# blocking host entry point for comparison
t0 = time.perf_counter()
for k in range(20): op(q_host[0].numpy(), f_host.numpy())
blk_ms = 1e3 * (time.perf_counter() - t0) / 20
print(json.dumps({"n_r": n_r, "pairs": op.info()["pairs_total"], "steps": steps, "device_ms_per_step": round(dev_ms, 4),
                  "e2e_pipelined_ms_per_step": round(e2e_ms, 4), "e2e_blocking_ms_per_step": round(blk_ms, 4)}))
