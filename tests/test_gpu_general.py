"""GPU parity of the general-grid path (csrc/bfsm_general.cuh): independent Nvx, Nvy, Nvz as the reference
interface carries them (FFTWBoltzmannOperator.hpp:30-36, per-axis mode tables .cpp:46-57), axis sizes that
are not powers of two, and 128.  The oracle is the C port, which tests/test_oracle.py pins to the
unmodified reference operator on non-cubic grids."""
import numpy as np
import pytest
import torch

import bfsm_b200 as B
from helpers import REL_LINF_TOL, inp, make_input, make_operator, oracle_args, quadrature, rel_linf

pytestmark = pytest.mark.gpu
capi = B.submodule("_capi")


def _eval(op, f):
    f_dev = torch.from_numpy(np.ascontiguousarray(f)).cuda().reshape(-1)
    q = torch.empty_like(f_dev)
    op(q, f_dev)
    torch.cuda.synchronize()
    return q.cpu().numpy()


@pytest.mark.parametrize("shape", [(32, 64, 16), (16, 24, 12), (64, 32, 32), (4, 8, 6), (128, 16, 16), (20, 20, 20)],
                         ids=lambda s: "x".join(map(str, s)))
@pytest.mark.parametrize("kind", ["maxmix", "noise"])
def test_general_grids_match_the_oracle(port_oracle, shape, kind):
    n_r, n_s = 3, 12
    op, gl, sd = make_operator(shape, n_r, n_s)
    info = op.info()
    assert info["general"] == 1 and (info["n"], info["ny"], info["nz"]) == shape
    f = make_input(kind, shape)
    q = _eval(op, f)
    assert np.array_equal(q, _eval(op, f))                      # deterministic
    ref = port_oracle.collide(shape, *oracle_args(gl, sd), f)
    assert rel_linf(q, ref) <= REL_LINF_TOL


def test_one_axis_of_128_points(port_oracle):
    """128^3: the plane no longer fits one SM for the tuned kernels; the general path takes it."""
    shape, n_r, n_s = (128, 128, 128), 1, 6
    op, gl, sd = make_operator(shape, n_r, n_s)
    f = make_input("maxmix", shape)
    ref = port_oracle.collide(shape, *oracle_args(gl, sd), f)
    assert rel_linf(_eval(op, f), ref) <= REL_LINF_TOL


@pytest.mark.parametrize("Nv", [16, 32])
def test_general_path_agrees_with_the_tuned_kernels_on_a_cubic_grid(port_oracle, Nv):
    n_r, n_s = 4, 12
    f = make_input("noise", Nv)
    tuned, gl, sd = make_operator(Nv, n_r, n_s)
    general, _, _ = make_operator(Nv, n_r, n_s, general=True)
    assert tuned.info()["general"] == 0 and general.info()["general"] == 1
    qt, qg = _eval(tuned, f), _eval(general, f)
    ref = port_oracle.collide((Nv,) * 3, *oracle_args(gl, sd), f)
    assert rel_linf(qg, ref) <= REL_LINF_TOL and rel_linf(qt, ref) <= REL_LINF_TOL
    assert rel_linf(qg, qt) <= 1e-13


def test_general_path_entry_points(port_oracle):
    """Same ABI surface as the tuned path: batches, host pointers, Q aliasing f, pair shards whose
    partial Q add up, chunk changes, an unfolded quadrature, moments."""
    shape, n_r, n_s = (16, 32, 8), 3, 12
    n = int(np.prod(shape))
    op, gl, sd = make_operator(shape, n_r, n_s)
    fs = np.stack([make_input("maxmix", shape, seed=c).reshape(-1) for c in range(3)])
    ref = [port_oracle.collide(shape, *oracle_args(gl, sd), fs[c].reshape(shape)).reshape(-1) for c in range(3)]
    f_dev = torch.from_numpy(fs).cuda().reshape(-1)
    q_dev = torch.empty_like(f_dev)
    op(q_dev, f_dev, n_cells=3)
    torch.cuda.synchronize()
    q = q_dev.cpu().numpy().reshape(3, n)
    for c in range(3):
        assert rel_linf(q[c], ref[c]) <= REL_LINF_TOL
    q_host = np.empty(n)
    op(q_host, fs[1].copy())                                    # host pointers
    assert np.array_equal(q_host, q[1])
    alias = torch.from_numpy(fs[2].copy()).cuda()
    op(alias, alias)                                            # Q == f_in
    torch.cuda.synchronize()
    assert np.array_equal(alias.cpu().numpy(), q[2])
    op.set_chunk(1)
    assert rel_linf(_eval(op, fs[0]), q[0]) <= 1e-14
    total = torch.zeros(n, dtype=torch.float64, device="cuda")
    part = torch.empty_like(total)
    f0 = torch.from_numpy(fs[0]).cuda()
    for r in range(3):
        shard = B.BoltzmannOperatorB200(gl, sd, *shape, 0.0, inp.B_GAMMA_MAXWELL, inp.L_DOMAIN,
                                        shard_index=r, shard_count=3)
        shard.initialize()
        shard.collide_partial(part, f0)
        total += part
    torch.cuda.synchronize()
    assert rel_linf(total.cpu().numpy(), ref[0]) <= REL_LINF_TOL
    unfolded = B.BoltzmannOperatorB200(gl, sd, *shape, 0.0, inp.B_GAMMA_MAXWELL, inp.L_DOMAIN, fold=False)
    unfolded.initialize()
    assert unfolded.info()["pairs_total"] == n_r * n_s
    assert rel_linf(_eval(unfolded, fs[0]), ref[0]) <= REL_LINF_TOL
    m = op.moments(f0).cpu().numpy()[0]
    vx, vy, vz = (inp.velocity_axis(k) for k in shape)
    dv3 = vx[1] * vy[1] * vz[1]
    assert abs(m[0] - fs[0].sum() * dv3) <= 1e-12 * abs(m[0])


def test_sizes_outside_the_general_path_are_rejected():
    gl, sd = quadrature(2, 6)
    for shape in ((15, 16, 16), (16, 16, 130), (2, 16, 16), (16, 0, 16)):
        op = B.BoltzmannOperatorB200(gl, sd, *shape, 0.0, inp.B_GAMMA_MAXWELL, inp.L_DOMAIN)
        with pytest.raises(capi.BfsmError) as err:
            op.initialize()
        assert err.value.code in (capi.BFSM_ERR_UNSUPPORTED, capi.BFSM_ERR_INVALID)
