"""GPU robustness: several plans alive at once, plan churn (no leak), user streams, CUDA-graph
capture/replay of one evaluation (the internal side-stream fork/join must be capturable)."""
import numpy as np
import pytest
import torch

import bfsm_b200 as B
from helpers import REL_LINF_TOL, inp, make_input, make_operator, oracle_args, quadrature, rel_linf

pytestmark = pytest.mark.gpu


def _dev(f):
    return torch.from_numpy(np.ascontiguousarray(f)).cuda().reshape(-1)


def test_interleaved_plans_of_different_sizes(port_oracle):
    op16, gl16, sd16 = make_operator(16, 4, 12)
    op32, gl32, sd32 = make_operator(32, 2, 12)
    f16, f32 = _dev(make_input("noise", 16)), _dev(make_input("maxmix", 32))
    q16, q32 = torch.empty_like(f16), torch.empty_like(f32)
    for _ in range(3):
        op16(q16, f16)
        op32(q32, f32)
    torch.cuda.synchronize()
    assert rel_linf(q16.cpu().numpy(), port_oracle.collide((16,) * 3, *oracle_args(gl16, sd16),
                                                          make_input("noise", 16))) <= REL_LINF_TOL
    assert rel_linf(q32.cpu().numpy(), port_oracle.collide((32,) * 3, *oracle_args(gl32, sd32),
                                                          make_input("maxmix", 32))) <= REL_LINF_TOL


def test_plan_churn_does_not_leak_device_memory():
    gl, sd = quadrature(4, 12)
    f = _dev(make_input("maxmix", 32))
    q = torch.empty_like(f)

    def cycle():
        op = B.BoltzmannOperatorB200(gl, sd, 32, 32, 32, 0.0, inp.B_GAMMA_MAXWELL, inp.L_DOMAIN)
        op.initialize()
        op(q, f)
        torch.cuda.synchronize()
        op.close()

    cycle()
    torch.cuda.synchronize()
    free0, _ = torch.cuda.mem_get_info()
    for _ in range(20):
        cycle()
    torch.cuda.synchronize()
    free1, _ = torch.cuda.mem_get_info()
    assert free0 - free1 < 8 * 1024 * 1024, f"device memory shrank by {(free0 - free1) / 2**20:.1f} MiB"


def test_user_stream_and_graph_replay():
    Nv = 32
    op, _, _ = make_operator(Nv, 4, 32)
    f = _dev(make_input("noise", Nv))
    q_ref = torch.empty_like(f)
    op(q_ref, f)
    torch.cuda.synchronize()

    s = torch.cuda.Stream()
    q_s = torch.empty_like(f)
    with torch.cuda.stream(s):
        op(q_s, f)          # the operator picks up torch's current stream
    s.synchronize()
    assert torch.equal(q_s, q_ref)

    # capture one evaluation into a CUDA graph and replay it on new input data
    g = torch.cuda.CUDAGraph()
    f_static, q_static = f.clone(), torch.empty_like(f)
    with torch.cuda.stream(s):
        op(q_static, f_static)      # warm-up on the capture stream
    s.synchronize()
    with torch.cuda.graph(g, stream=s):
        op(q_static, f_static)
    f2 = _dev(make_input("maxmix", Nv))
    q2_ref = torch.empty_like(f2)
    op(q2_ref, f2)
    torch.cuda.synchronize()
    f_static.copy_(f2)
    g.replay()
    torch.cuda.synchronize()
    assert torch.equal(q_static, q2_ref)
