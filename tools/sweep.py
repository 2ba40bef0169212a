#!/usr/bin/env python3
"""Quick device-side timing sweep over BFSM_CHUNK_PAIRS / BFSM_GAIN_CTAS (tuning aid, not a benchmark)."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bfsm_b200 as B
inp = B.inputs

def run(Nv, n_r, n_s, chunk, gy=None, reps=2):
    os.environ["BFSM_CHUNK_PAIRS"] = str(chunk)
    if gy: os.environ["BFSM_GAIN_CTAS"] = str(gy)
    else: os.environ.pop("BFSM_GAIN_CTAS", None)
    gl = B.GaussLegendreQuadrature(n_r, 0.0, inp.R_SUPPORT); sd = B.SphericalDesign(n_s)
    op = B.BoltzmannOperatorB200(gl, sd, Nv, Nv, Nv, 0.0, inp.B_GAMMA_MAXWELL, inp.L_DOMAIN)
    op.initialize()
    f = torch.from_numpy(inp.maxmix(Nv)).cuda().reshape(-1); q = torch.empty_like(f)
    op(q, f); torch.cuda.synchronize()
    a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): op(q, f)
    b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b) / reps
    prof = op.profile(q, f)
    info = op.info(); op.close()
    print(json.dumps({"Nv": Nv, "n_r": n_r, "n_s": n_s, "chunk": chunk, "gy": gy, "ms_per_eval": round(ms, 3),
                      "evals_per_s": round(1e3 / ms, 2), "pairs": info["pairs_total"],
                      "us_per_pair": round(1e3 * ms / info["pairs_total"], 3),
                      "prof_ms": {k: round(v[0], 3) for k, v in prof.items()}}), flush=True)

if __name__ == "__main__":
    which = sys.argv[1] if len(sys.argv) > 1 else "64"
    if which == "64":
        for chunk in (16, 48, 96, 192, 384):
            run(64, 8, 192, chunk)
    elif which == "32":
        for chunk in (16, 32, 64, 128, 256):
            run(32, 16, 32, chunk)
