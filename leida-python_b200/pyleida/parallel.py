"""Multi-GPU plumbing for the LEiDA path: one process per GPU, subjects sharded over ranks.

Stages 1, 2 and 4 are independent per subject (reference loop body: _leida.py:279; per-subject
groups: dynamics_metrics/_dynamics_metrics.py:180-189), so they need no communication.  k-means
has one exchange per Lloyd iteration: the sum over ranks of the int64 fixed-point centroid sums,
member counts and 'labels changed' counters.  Integer sums make the result independent
of the number of ranks.  torch.distributed is used for the collective (NCCL on GPUs; the same code
runs over gloo with CPU tensors in the tests).
"""
import numpy as np


def shard_bounds(n_items, world, rank):
    """Contiguous, balanced [lo, hi) block of n_items owned by `rank` (first n%world ranks get one more)."""
    base, rem = divmod(int(n_items), int(world))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def row_offsets(rows_per_rank):
    """Exclusive prefix sum: global row index of each rank's first row, plus the total."""
    off = np.zeros(len(rows_per_rank) + 1, dtype=np.int64)
    np.cumsum(np.asarray(rows_per_rank, dtype=np.int64), out=off[1:])
    return off


class Comm:
    """Thin wrapper over torch.distributed (or a single-process stand-in when world == 1)."""

    def __init__(self, group=None):
        import torch.distributed as dist
        self._dist = dist
        self.group = group
        if dist.is_available() and dist.is_initialized():
            self.world = dist.get_world_size(group)
            self.rank = dist.get_rank(group)
        else:
            self.world, self.rank = 1, 0

    def all_reduce_sum_(self, t):
        if self.world > 1:
            self._dist.all_reduce(t, op=self._dist.ReduceOp.SUM, group=self.group)
        return t

    def all_gather_int(self, value):
        """Gather one python int from every rank -> list."""
        import torch
        if self.world == 1:
            return [int(value)]
        dev = "cuda" if self._dist.get_backend(self.group) == "nccl" else "cpu"
        mine = torch.tensor([int(value)], dtype=torch.int64, device=dev)
        out = [torch.zeros_like(mine) for _ in range(self.world)]
        self._dist.all_gather(out, mine, group=self.group)
        return [int(o.item()) for o in out]

    def barrier(self):
        if self.world > 1:
            self._dist.barrier(group=self.group)


def gather_seed_rows(X_local, n_cols, global_rows, row_lo, comm):
    """Rows `global_rows` (global indices) of the row-sharded matrix, on every rank.

    Each seed row lives on exactly one rank; the others contribute zeros, so a sum all-reduce
    reproduces the rows bit for bit.  X_local: (M_local, pitch) tensor holding global rows
    [row_lo, row_lo + M_local).  Returns a (len(global_rows), pitch) tensor.
    """
    import torch
    idx = torch.as_tensor(np.asarray(global_rows, dtype=np.int64), device=X_local.device)
    m_local = X_local.shape[0]
    mine = (idx >= row_lo) & (idx < row_lo + m_local)
    out = torch.zeros((idx.numel(), X_local.shape[1]), dtype=X_local.dtype, device=X_local.device)
    if bool(mine.any()):
        out[mine] = X_local[(idx[mine] - row_lo)]
    comm.all_reduce_sum_(out)
    return out
