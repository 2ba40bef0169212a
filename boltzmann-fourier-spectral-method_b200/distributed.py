"""Multi-GPU execution: one process per GPU, torch.distributed for the plumbing.

Two shardings exist on this path:

* quadrature-pair sharding (BASELINE config 4): the gain spectrum is a plain sum over
  independent (r, sigma) pairs, so rank k evaluates shard k of the pair list
  (`bfsm_gain_hat`), the N^3 complex partial spectra are summed with ONE all-reduce
  (NCCL over NVLink on GPUs, gloo in the CPU tests) and every rank finishes locally
  (`bfsm_finish`: loss term, inverse transform, combine).
* cell sharding (BASELINE config 5): independent spatial cells are split across ranks;
  no data-path collective at all.

The classes here only orchestrate; the arithmetic is behind the `local` operator object
(`BoltzmannOperatorB200` on a GPU).  Tests inject a CPU stand-in to exercise the collective
logic under gloo with world_size 2.
"""
import torch
import torch.distributed as dist


def shard_range(total, index, count):
    """Contiguous, equal (+-1) split -- the same arithmetic as bfsm_plan_create."""
    if count < 1 or not (0 <= index < count):
        raise ValueError("shard index/count out of range")
    return (total * index) // count, (total * (index + 1)) // count


def shard_cells(n_cells, rank, world_size):
    """Cells [lo, hi) owned by `rank` in batch (space-inhomogeneous) mode."""
    return shard_range(n_cells, rank, world_size)


class NcclCommunicator:
    """A bfsm_comm (include/bfsm_b200.h): an NCCL communicator owned by the C library, one rank per
    process.  Rank 0 draws the NCCL unique id; the 128 bytes travel to the other ranks through the
    already initialised torch.distributed group (any backend) -- torch is only the side channel."""

    def __init__(self, device, group=None):
        import ctypes
        from . import _capi
        lib = _capi.load()
        rank, world = dist.get_rank(group), dist.get_world_size(group)
        uid = ctypes.create_string_buffer(_capi.BFSM_UNIQUE_ID_BYTES)
        if rank == 0:
            _capi.check(lib.bfsm_comm_unique_id(uid))
        box = [uid.raw if rank == 0 else None]
        dist.broadcast_object_list(box, src=dist.get_global_rank(group, 0) if group is not None else 0,
                                   group=group)
        self._lib = lib
        self.handle = ctypes.c_void_p()
        self.rank, self.world_size, self.device = rank, world, int(device)
        _capi.check(lib.bfsm_comm_init_rank(ctypes.byref(self.handle), box[0], world, rank, self.device))

    def close(self):
        if self.handle:
            self._lib.bfsm_comm_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class PairShardedCollision:
    """Q(f,f) with the (r, sigma) pairs sharded over the ranks of `group`.

    `local` must provide gain_hat(Qhat, f) and finish(Q, Qhat, f) for THIS rank's shard
    (for the CUDA path: BoltzmannOperatorB200(..., shard_index=rank, shard_count=world)).
    """

    def __init__(self, local, grid_size, group=None, comm=None):
        self.local = local
        self.grid_size = int(grid_size)
        self.group = group
        #: NcclCommunicator: the whole step, collective included, runs inside the C library
        #: (bfsm_collide_sharded); None: gain_hat / all_reduce through torch.distributed / finish
        self.comm = comm
        self._qhat = None

    def _buffer(self, like):
        if self._qhat is None or self._qhat.device != like.device:
            self._qhat = torch.empty(2 * self.grid_size, dtype=torch.float64, device=like.device)
        return self._qhat

    def computeCollision(self, Q, f_in):
        if self.comm is not None:
            return self.local.collide_sharded(Q, f_in, self.comm)
        qhat = self._buffer(f_in)
        self.local.gain_hat(qhat, f_in)
        if dist.is_available() and dist.is_initialized() and dist.get_world_size(self.group) > 1:
            # the path's one exchange step: N^3 complex doubles (4 MiB at 64^3)
            dist.all_reduce(qhat, op=dist.ReduceOp.SUM, group=self.group)
        self.local.finish(Q, qhat, f_in)
        return Q

    def __call__(self, Q, f_in):
        return self.computeCollision(Q, f_in)


class CellShardedCollision:
    """Batch mode: each rank evaluates its own contiguous block of cells, no collective."""

    def __init__(self, local, grid_size, rank, world_size):
        self.local = local
        self.grid_size = int(grid_size)
        self.rank, self.world_size = int(rank), int(world_size)

    def local_cells(self, n_cells):
        return shard_cells(n_cells, self.rank, self.world_size)

    def computeCollision(self, Q_local, f_local):
        """Q_local / f_local hold only this rank's cells."""
        return self.local.computeCollision(Q_local, f_local)

    def __call__(self, Q_local, f_local):
        return self.computeCollision(Q_local, f_local)
