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


class PairShardedCollision:
    """Q(f,f) with the (r, sigma) pairs sharded over the ranks of `group`.

    `local` must provide gain_hat(Qhat, f) and finish(Q, Qhat, f) for THIS rank's shard
    (for the CUDA path: BoltzmannOperatorB200(..., shard_index=rank, shard_count=world)).
    """

    def __init__(self, local, grid_size, group=None):
        self.local = local
        self.grid_size = int(grid_size)
        self.group = group
        self._qhat = None

    def _buffer(self, like):
        if self._qhat is None or self._qhat.device != like.device:
            self._qhat = torch.empty(2 * self.grid_size, dtype=torch.float64, device=like.device)
        return self._qhat

    def computeCollision(self, Q, f_in):
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
