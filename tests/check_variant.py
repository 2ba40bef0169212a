#!/usr/bin/env python3
"""Compare a kernel variant (BFSM_* env knobs given as KEY=VAL arguments) with the default path:
relative L-inf difference of Q on maxmix / noise inputs at 64^3 (32 x 192 and 2 x 12), run-to-run
determinism of the variant, and -- for the small case -- both against the CPU oracle.

    python tests/check_variant.py BFSM_PLANE_WS=1

(lives under tests/: it calls the CPU oracle, which only the test side may do)
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))   # helpers.py
import numpy as np
import torch
import bfsm_b200 as B
from helpers import make_input, make_operator, oracle_args, rel_linf

env = dict(kv.split("=", 1) for kv in sys.argv[1:])


def evaluate(op, f):
    f_dev = torch.from_numpy(f).cuda().reshape(-1)
    q = torch.empty_like(f_dev)
    op(q, f_dev)
    torch.cuda.synchronize()
    return q.cpu().numpy()


def plan(Nv, n_r, n_s, variant):
    for k, v in env.items():
        if variant:
            os.environ[k] = v
        else:
            os.environ.pop(k, None)
    op, gl, sd = make_operator(Nv, n_r, n_s)
    for k in env:
        os.environ.pop(k, None)
    return op, gl, sd


for Nv, n_r, n_s in ((64, 2, 12), (64, 32, 192)):
    op0, gl, sd = plan(Nv, n_r, n_s, False)
    op1, _, _ = plan(Nv, n_r, n_s, True)
    for kind in ("maxmix", "noise", "bkw"):
        f = make_input(kind, Nv)
        q0 = evaluate(op0, f)
        q1 = evaluate(op1, f)
        q1b = evaluate(op1, f)
        rec = {"case": [Nv, n_r, n_s], "input": kind, "variant": env,
               "relLinf_variant_vs_default": rel_linf(q1, q0),
               "variant_deterministic": bool(np.array_equal(q1, q1b))}
        if n_r * n_s <= 64:
            from oracle import oracle as O
            ref = O.PortOracle().collide((Nv,) * 3, *oracle_args(gl, sd), f)
            rec["relLinf_default_vs_oracle"] = rel_linf(q0, ref)
            rec["relLinf_variant_vs_oracle"] = rel_linf(q1, ref)
        print(json.dumps(rec), flush=True)
    op0.close()
    op1.close()
