#!/usr/bin/env python3
"""A/B timing of kernel variants selected by BFSM_* environment knobs (tuning aid, not a benchmark).

    python tools/ab_plane.py one <Nv> <n_r> <n_s> <ref.npy|-> KEY=VAL ...     one configuration
    python tools/ab_plane.py all [64|32] ["KEY=VAL KEY=VAL" ...]              a list of variants

`all` runs every variant in its own subprocess under a timeout, so that a schedule that deadlocks
(the pipelined plane kernel hands buffers over through hand-rolled named barriers) costs one line of
the table, not the GPU call.  The first variant is the reference: every other variant's Q is compared
with it, bitwise and as a relative L-infinity difference (kernel variants that only reschedule are
bitwise equal; variants that change the summation order differ at rounding level).
"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

DEFAULT_VARIANTS = ["", "BFSM_PENCIL_KERNEL=1", "BFSM_PLANE_KERNEL=1"]
#: x-stage work-unit sizes: `all 64 next`
NEXT_VARIANTS = ["", "BFSM_PENCIL_KERNEL=1", "BFSM_SEG_PAIRS=12", "BFSM_SEG_PAIRS=16", "BFSM_SEG_PAIRS=32",
                 "BFSM_SEG_PAIRS=48", "BFSM_SEG_PAIRS=96"]


def one(Nv, n_r, n_s, ref_path, env, reps=5):
    for kv in env:
        k, v = kv.split("=", 1)
        os.environ[k] = v
    import numpy as np
    import torch
    import bfsm_b200 as B
    inp = B.inputs
    gl = B.GaussLegendreQuadrature(n_r, 0.0, inp.R_SUPPORT)
    sd = B.SphericalDesign(n_s)
    op = B.BoltzmannOperatorB200(gl, sd, Nv, Nv, Nv, 0.0, inp.B_GAMMA_MAXWELL, inp.L_DOMAIN)
    op.initialize()
    f = torch.from_numpy(inp.maxmix(Nv)).cuda().reshape(-1)
    q = torch.empty_like(f)
    for _ in range(3):
        op(q, f)
    torch.cuda.synchronize()
    a = torch.cuda.Event(enable_timing=True)
    b = torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        op(q, f)
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / reps
    prof = op.profile(q, f)
    info = op.info()
    qh = q.cpu().numpy()
    same = rel = None
    if ref_path and ref_path != "-":
        if os.path.exists(ref_path):
            ref = np.load(ref_path)
            same = bool(np.array_equal(ref, qh))
            rel = float(np.abs(ref - qh).max() / np.abs(ref).max())
        else:
            np.save(ref_path, qh)
    op.close()
    print(json.dumps({"Nv": Nv, "n_r": n_r, "n_s": n_s, "env": " ".join(env), "chunk": info["chunk_pairs"],
                      "ms_per_eval": round(ms, 4), "evals_per_s": round(1e3 / ms, 2),
                      "us_per_pair": round(1e3 * ms / info["pairs_total"], 4),
                      "bitwise_equal_to_first": same, "rel_linf_vs_first": rel,
                      "prof_ms": {k: round(v[0], 3) for k, v in prof.items()}}), flush=True)


def run_all(which, variants):
    out_dir = os.path.join(ROOT, "gpurun_out")
    os.makedirs(out_dir, exist_ok=True)
    shape = (64, 32, 192) if which == "64" else (32, 16, 94)
    ref = os.path.join(out_dir, "ab_ref_%d.npy" % shape[0])
    if os.path.exists(ref):
        os.remove(ref)
    for var in variants:
        cmd = [sys.executable, os.path.abspath(__file__), "one", *map(str, shape), ref, *var.split()]
        try:
            r = subprocess.run(cmd, capture_output=True, text=True, timeout=120)
            line = r.stdout.strip().splitlines()[-1] if r.stdout.strip() else "no output: " + r.stderr[-400:]
        except subprocess.TimeoutExpired:
            line = json.dumps({"env": var, "error": "timeout (deadlock?)"})
        print(line, flush=True)


if __name__ == "__main__":
    if sys.argv[1] == "one":
        Nv, n_r, n_s = (int(a) for a in sys.argv[2:5])
        one(Nv, n_r, n_s, sys.argv[5], sys.argv[6:])
    else:
        which = sys.argv[2] if len(sys.argv) > 2 else "64"
        variants = sys.argv[3:] or DEFAULT_VARIANTS
        if variants == ["next"]:
            variants = NEXT_VARIANTS
        run_all(which, variants)
