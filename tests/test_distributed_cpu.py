"""world_size-2 gloo test of the multi-GPU orchestration (CPU only).

The collective logic of PairShardedCollision (partial gain spectrum -> ONE all-reduce -> local
finish) runs under gloo with the CPU oracle standing in for the per-rank CUDA operator."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class OracleShard:
    """CPU stand-in with the gain_hat / finish surface of BoltzmannOperatorB200."""

    def __init__(self, port, shape, args, lo, hi):
        self.port, self.shape, self.args, self.lo, self.hi = port, shape, args, lo, hi

    def gain_hat(self, Qhat, f):
        part = self.port.gain_hat(self.shape, *self.args, f.numpy(), self.lo, self.hi)
        Qhat.copy_(torch.from_numpy(np.ascontiguousarray(part).ravel().view(np.float64)))
        return Qhat

    def finish(self, Q, Qhat, f):
        a = self.args
        q = self.port.finish(self.shape, a[0], a[1], a[6], a[7], a[8], f.numpy(),
                             Qhat.numpy().view(np.complex128).reshape(self.shape))
        Q.copy_(torch.from_numpy(q.ravel()))
        return Q


def _worker(rank, world, port_file, out_dir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port_file)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import bfsm_b200 as B
    from helpers import make_input, oracle_args, quadrature
    from oracle import oracle as O
    D = B.submodule("distributed")
    Nv, n_r, n_s = 16, 4, 6
    gl, sd = quadrature(n_r, n_s)
    args = oracle_args(gl, sd)
    port = O.PortOracle()
    port.set_threads(1)
    lo, hi = D.shard_range(n_r * n_s, rank, world)
    op = D.PairShardedCollision(OracleShard(port, (Nv,) * 3, args, lo, hi), Nv ** 3)
    f = torch.from_numpy(make_input("noise", Nv).ravel().copy())
    Q = torch.empty_like(f)
    op(Q, f)
    np.save(os.path.join(out_dir, f"Q_rank{rank}.npy"), Q.numpy())
    # cell sharding: ranks own disjoint contiguous blocks that cover all cells, no collective
    cells = D.shard_cells(5, rank, world)
    np.save(os.path.join(out_dir, f"cells_rank{rank}.npy"), np.array(cells))
    dist.barrier()
    dist.destroy_process_group()


def test_pair_sharded_collision_world2_gloo(tmp_path):
    import socket
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)

    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from helpers import make_input, oracle_args, quadrature, rel_linf
    from oracle import oracle as O
    Nv, n_r, n_s = 16, 4, 6
    gl, sd = quadrature(n_r, n_s)
    Q_full = O.PortOracle().collide((Nv,) * 3, *oracle_args(gl, sd), make_input("noise", Nv))
    Q0 = np.load(tmp_path / "Q_rank0.npy")
    Q1 = np.load(tmp_path / "Q_rank1.npy")
    assert np.array_equal(Q0, Q1), "every rank must hold the same Q after the all-reduce"
    assert rel_linf(Q0, Q_full) < 1e-13
    c0, c1 = np.load(tmp_path / "cells_rank0.npy"), np.load(tmp_path / "cells_rank1.npy")
    assert c0[0] == 0 and c0[1] == c1[0] and c1[1] == 5
