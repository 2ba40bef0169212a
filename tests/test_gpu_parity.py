"""GPU parity: the CUDA path (through the C ABI) against the CPU oracle on identical inputs.

Bar (BASELINE.json north_star): max|Q_new - Q_ref| / max|Q_ref| <= 1e-12 in fp64.
"""
import numpy as np
import pytest
import torch

from helpers import REL_LINF_TOL, make_input, make_operator, oracle_args, rel_linf

pytestmark = pytest.mark.gpu

# (Nv, N_r, N_sigma): BASELINE configs 1 and 2, and a 64^3 case small enough for the CPU oracle
SMALL_CASES = [(16, 8, 6), (32, 16, 32), (64, 2, 12)]


@pytest.mark.parametrize("pack", [True, False], ids=["packed", "unpacked"])
@pytest.mark.parametrize("Nv,n_r,n_s", SMALL_CASES)
@pytest.mark.parametrize("kind", ["bkw", "maxmix", "noise"])
def test_collide_matches_oracle(port_oracle, Nv, n_r, n_s, kind, pack):
    """`noise` is not band limited: it exercises the exact Nyquist-plane correction of the
    Hermitian-packed path (the naive packing is off by ~30 % there)."""
    op, gl, sd = make_operator(Nv, n_r, n_s, pack=pack)
    assert op.info()["packed"] == int(pack)
    f = make_input(kind, Nv)
    f_dev = torch.from_numpy(f).cuda()
    Q_dev = torch.empty_like(f_dev)
    op(Q_dev, f_dev)
    torch.cuda.synchronize()
    Q_ref = port_oracle.collide((Nv,) * 3, *oracle_args(gl, sd), f)
    err = rel_linf(Q_dev.cpu().numpy(), Q_ref)
    assert err <= REL_LINF_TOL, f"relLinf {err:.3e} > {REL_LINF_TOL}"
