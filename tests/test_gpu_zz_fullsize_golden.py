"""GPU parity at BASELINE config 4 itself (64^3, 32 radii x 192-point design, 6144 pairs): the CUDA
path against a golden Q written by the C port of the reference algorithm, which needs ~40 s of CPU
time per input there (tests/golden/make_golden_cfg4.py; the port is pinned to the unmodified
reference operator at 64^3 to 2e-15).  Every second point per axis is compared elementwise, the rest
through per-x-plane sums of Q and Q^2.  Named to run last: it is the most expensive GPU test."""
import os

import numpy as np
import pytest
import torch

from helpers import REL_LINF_TOL, make_input, make_operator, oracle_args, rel_linf

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = np.load(os.path.join(ROOT, "tests", "golden", "port_q_cfg4.npz"))


@pytest.mark.parametrize("kind", ["maxmix", "noise"])
def test_flagship_config_matches_port_oracle_golden(kind):
    Nv, n_r, n_s = 64, 32, 192
    op, _, _ = make_operator(Nv, n_r, n_s)
    f_dev = torch.from_numpy(make_input(kind, Nv)).cuda().reshape(-1)
    Q_dev = torch.empty_like(f_dev)
    op(Q_dev, f_dev)
    torch.cuda.synchronize()
    Q = Q_dev.cpu().numpy().reshape(Nv, Nv, Nv)
    key, st = f"Nv{Nv}_r{n_r}_s{n_s}_{kind}", int(GOLD["stride"])
    qmax = float(GOLD[key + "_max"])
    err = np.abs(Q[::st, ::st, ::st] - GOLD[key + "_Qsub"]).max() / qmax
    assert err <= REL_LINF_TOL, f"relLinf {err:.3e}"
    assert np.abs(Q.sum(axis=(1, 2)) - GOLD[key + "_plane_sum"]).max() / (qmax * Nv * Nv) <= REL_LINF_TOL
    assert (np.abs((Q * Q).sum(axis=(1, 2)) - GOLD[key + "_plane_sumsq"]).max()
            / (qmax ** 2 * Nv * Nv) <= REL_LINF_TOL)


@pytest.mark.parametrize("Nv,n_r,n_s,kind,seed", [(32, 32, 48, "maxmix", 0), (32, 32, 48, "noise", 0),
                                                  (32, 16, 94, "maxmix", 3), (32, 16, 94, "noise", 1)],
                         ids=["cfg3-maxmix", "cfg3-noise", "cfg5cell-maxmix", "cfg5cell-noise"])
def test_32_cubed_baseline_configs_match_the_port_oracle(port_oracle, Nv, n_r, n_s, kind, seed):
    """BASELINE configs 3 (32^3, 48-point design; N_r = 32 as the stock driver would choose) and 5 (one
    32^3 cell, 16 x 94) against the C port run live: a few seconds of CPU time each."""
    op, gl, sd = make_operator(Nv, n_r, n_s)
    f = make_input(kind, Nv, seed)
    f_dev = torch.from_numpy(f).cuda().reshape(-1)
    Q_dev = torch.empty_like(f_dev)
    op(Q_dev, f_dev)
    torch.cuda.synchronize()
    Q_ref = port_oracle.collide((Nv,) * 3, *oracle_args(gl, sd), f)
    err = rel_linf(Q_dev.cpu().numpy(), Q_ref)
    assert err <= REL_LINF_TOL, f"relLinf {err:.3e}"
