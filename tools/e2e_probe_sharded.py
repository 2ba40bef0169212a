#!/usr/bin/env python3
"""Device-timed loop vs pipelined host-buffer loop for a PAIR-SHARDED plan (run under torchrun): same
per-rank work as the 8-rank headline run (384 pairs per rank) with fewer ranks.
    torchrun --nproc-per-node 2 tools/e2e_probe_sharded.py 8 200      # n_r = 4 x ranks"""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
import bfsm_b200 as B
inp = B.inputs
capi = B.submodule("_capi"); D = B.submodule("distributed")
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank)
dist.init_process_group("nccl", device_id=torch.device("cuda", rank))
n_r = int(sys.argv[1]) if len(sys.argv) > 1 else 4 * world
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 200
Nv, n_s = 64, 192
gl = B.GaussLegendreQuadrature(n_r, 0.0, inp.R_SUPPORT); sd = B.SphericalDesign(n_s)
op = B.BoltzmannOperatorB200(gl, sd, Nv, Nv, Nv, 0.0, inp.B_GAMMA_MAXWELL, inp.L_DOMAIN, device=rank,
                             shard_index=rank, shard_count=world)
op.initialize()
comm = D.NcclCommunicator(rank)
sh = D.PairShardedCollision(op, Nv ** 3, comm=comm)
f_host = torch.from_numpy(inp.maxmix(Nv).reshape(-1).copy()).pin_memory()
depth = capi.BFSM_HOST_PIPE_DEPTH
q_host = [torch.empty(Nv ** 3, dtype=torch.float64).pin_memory() for _ in range(depth)]
f = f_host.cuda(); q = torch.empty_like(f)
def mx(x):
    t = torch.tensor([x], dtype=torch.float64, device="cuda"); dist.all_reduce(t, op=dist.ReduceOp.MAX); return float(t)
for _ in range(5): sh(q, f)
torch.cuda.synchronize(); dist.barrier()
a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(steps): sh(q, f)
b.record(); torch.cuda.synchronize()
dev_ms = mx(a.elapsed_time(b) / steps)
for k in range(depth): op.submit_host(q_host[k], f_host, comm=comm)
op.flush_host(); dist.barrier()
t0 = time.perf_counter()
for k in range(steps): op.submit_host(q_host[k % depth], f_host, comm=comm)
op.flush_host()
e2e_ms = mx(1e3 * (time.perf_counter() - t0) / steps)
if rank == 0:
    print(json.dumps({"ranks": world, "n_r": n_r, "pairs_per_rank": op.info()["pairs_local"], "steps": steps,
                      "device_ms_per_step": round(dev_ms, 4), "e2e_pipelined_ms_per_step": round(e2e_ms, 4)}))
comm.close(); dist.destroy_process_group()
