"""Real multi-GPU check (needs >= 2 GPUs; skipped otherwise): pair-sharded evaluation with one NCCL
all-reduce per evaluation must reproduce the single-GPU result."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r'''
import os, sys
sys.path.insert(0, os.environ["BFSM_ROOT"]); sys.path.insert(0, os.path.join(os.environ["BFSM_ROOT"], "tests"))
import numpy as np, torch, torch.distributed as dist
import bfsm_b200 as B
from helpers import make_input, inp, quadrature
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank)
dist.init_process_group("nccl", device_id=torch.device("cuda", rank))
D = B.submodule("distributed")
Nv, n_r, n_s = 32, 6, 32
gl, sd = quadrature(n_r, n_s)
f = torch.from_numpy(make_input("noise", Nv)).cuda().reshape(-1)
full = B.BoltzmannOperatorB200(gl, sd, Nv, Nv, Nv, 0.0, inp.B_GAMMA_MAXWELL, inp.L_DOMAIN, device=rank)
full.initialize()
Q_full = torch.empty_like(f); full(Q_full, f)
shard = B.BoltzmannOperatorB200(gl, sd, Nv, Nv, Nv, 0.0, inp.B_GAMMA_MAXWELL, inp.L_DOMAIN, device=rank,
                                shard_index=rank, shard_count=world)
shard.initialize()
# (a) collective in Python: gain_hat / torch.distributed all_reduce of the spectrum / finish
op = D.PairShardedCollision(shard, Nv ** 3)
Q = torch.empty_like(f); op(Q, f); torch.cuda.synchronize()
err = float((Q - Q_full).abs().max() / Q_full.abs().max())
gathered = [torch.empty_like(Q) for _ in range(world)]
dist.all_gather(gathered, Q)
same = all(torch.equal(g, gathered[0]) for g in gathered)
# (b) collective under the C ABI: bfsm_collide_sharded (one ncclAllReduce of the real partial Q)
comm = D.NcclCommunicator(rank)
opc = D.PairShardedCollision(shard, Nv ** 3, comm=comm)
Qc = torch.empty_like(f); opc(Qc, f); opc(Qc, f); torch.cuda.synchronize()
errc = float((Qc - Q_full).abs().max() / Q_full.abs().max())
gathered = [torch.empty_like(Qc) for _ in range(world)]
dist.all_gather(gathered, Qc)
samec = all(torch.equal(g, gathered[0]) for g in gathered)
# (c) streamed host buffers through the sharded plan (bfsm_collide_host_async with the communicator, four
# steps in flight): each step's Q must equal (b) bit for bit
steps = 7
fh = f.cpu().pin_memory()
qh = [torch.empty(Nv ** 3, dtype=torch.float64).pin_memory() for _ in range(steps)]
for k in range(steps):
    shard.submit_host(qh[k], fh, comm=comm)
shard.flush_host()
sameh = all(torch.equal(q, Qc.cpu()) for q in qh)
if rank == 0:
    print(f"RESULT err={err:.3e} same={same} errc={errc:.3e} samec={samec} sameh={sameh} pairs_local={shard.info()['pairs_local']}")
comm.close()
dist.destroy_process_group()
'''


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs >= 2 GPUs")
def test_pair_sharding_over_nccl(tmp_path):
    world = min(torch.cuda.device_count(), 8)
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    env = dict(os.environ, BFSM_ROOT=ROOT)
    out = subprocess.run(
        [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
         "--master-addr", "127.0.0.1", "--master-port", "29611", str(script)],
        capture_output=True, text=True, env=env, timeout=600)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    line = [l for l in out.stdout.splitlines() if l.startswith("RESULT")][0]
    err = float(line.split("err=")[1].split()[0])
    errc = float(line.split("errc=")[1].split()[0])
    assert err <= 1e-13 and errc <= 1e-13, line
    assert "same=True" in line and "samec=True" in line and "sameh=True" in line, line


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs >= 2 GPUs")
def test_single_process_group_call_over_all_gpus():
    """The C++ hosts' path: ONE process, ncclCommInitAll over every GPU, bfsm_collide_sharded_group
    enqueues all ranks' work and issues the all-reduces as one NCCL group."""
    import ctypes
    import numpy as np
    import bfsm_b200 as B
    from helpers import inp, make_input, quadrature
    capi = B.submodule("_capi")
    lib = capi.load()
    world = min(torch.cuda.device_count(), 8)
    Nv, n_r, n_s = 32, 6, 32
    gl, sd = quadrature(n_r, n_s)
    f = make_input("noise", Nv).reshape(-1)
    full = B.BoltzmannOperatorB200(gl, sd, Nv, Nv, Nv, 0.0, inp.B_GAMMA_MAXWELL, inp.L_DOMAIN, device=0)
    full.initialize()
    f0 = torch.from_numpy(f).to("cuda:0")
    Q_full = torch.empty_like(f0)
    full(Q_full, f0)
    torch.cuda.synchronize()
    devs = (ctypes.c_int * world)(*range(world))
    comms = (ctypes.c_void_p * world)()
    capi.check(lib.bfsm_comm_init_all(comms, world, devs))
    ops, fs, qs = [], [], []
    for k in range(world):
        op = B.BoltzmannOperatorB200(gl, sd, Nv, Nv, Nv, 0.0, inp.B_GAMMA_MAXWELL, inp.L_DOMAIN, device=k,
                                     shard_index=k, shard_count=world)
        op.initialize()
        ops.append(op)
        fs.append(torch.from_numpy(f).to(f"cuda:{k}"))
        qs.append(torch.empty_like(fs[-1]))
    for k in range(world):
        torch.cuda.synchronize(k)
    plans = (ctypes.c_void_p * world)(*[op._plan for op in ops])
    qp = (ctypes.c_void_p * world)(*[q.data_ptr() for q in qs])
    fp = (ctypes.c_void_p * world)(*[t.data_ptr() for t in fs])
    capi.check(lib.bfsm_collide_sharded_group(world, plans, comms, qp, fp, None))
    for k in range(world):
        torch.cuda.synchronize(k)
    for k in range(world):
        err = float((qs[k].to("cuda:0") - Q_full).abs().max() / Q_full.abs().max())
        assert err <= 1e-13, (k, err)
        assert torch.equal(qs[k].cpu(), qs[0].cpu())
    for k in range(world):
        lib.bfsm_comm_destroy(comms[k])
