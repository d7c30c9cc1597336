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

    @property
    def backend(self):
        return self._dist.get_backend(self.group) if self.world > 1 else "none"

    def _device(self):
        import torch
        return torch.device("cuda", torch.cuda.current_device()) if self.backend == "nccl" else torch.device("cpu")

    def all_reduce_max_(self, t):
        if self.world > 1:
            self._dist.all_reduce(t, op=self._dist.ReduceOp.MAX, group=self.group)
        return t

    def bcast0_int64(self, arr):
        """numpy int64 array of rank 0 on every rank (the other ranks pass an array of the same shape)."""
        import torch
        if self.world == 1:
            return arr
        t = torch.from_numpy(np.ascontiguousarray(arr, dtype=np.int64)).to(self._device())
        if self.rank != 0:
            t.zero_()
        self._dist.all_reduce(t, op=self._dist.ReduceOp.SUM, group=self.group)
        return t.cpu().numpy()

    def all_gather_object(self, obj):
        """Gather one picklable object from every rank -> list (rank order)."""
        if self.world == 1:
            return [obj]
        out = [None] * self.world
        self._dist.all_gather_object(out, obj, group=self.group)
        return out

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


class PeerExchange:
    """The per-pass exchange of the k-means partial sums over symmetric memory (torch.distributed._symmetric_memory;
    SURVEY 8e), fused into the kernel that consumes it.  Every rank owns a buffer of two halves (+ one flag word per
    rank) that every other rank of the node maps.  One pass:
        leida_kmeans_exchange_publish  copies the rank's int64 words of the still-active runs into this pass's half
                                       and raises the rank's flag on every rank;
        leida_kmeans_update_peers      waits for all flags, then reads every word as the sum over ranks -- one
                                       multimem.ld_reduce through the NVSwitch multicast mapping when the group has
                                       one, else one load per rank over NVLink.
    No NCCL launch, no host involvement, no separate barrier kernel; the payload shrinks with the number of active
    runs.  Two halves alternate, so a rank that is one pass ahead never overwrites what a slower rank is still
    reading: it cannot be two passes ahead because its next update waits for the slower rank's next publish.

    get() is collective.  It returns None on every rank if any rank cannot set the mapping up (no symmetric-memory
    support, ranks on different nodes) or if LEIDA_EXCHANGE=nccl; the caller then uses copy + ncclAllReduce.
    LEIDA_EXCHANGE = auto (default: symmetric memory when available) | peer (same, loads per rank even when multicast
    exists) | nccl.  Buffers are cached per (group, device) and reused by later engines."""
    _cache = {}
    FLAG_PITCH = 4096               # runs per batch that can be exchanged (one flag word per (rank, run))

    def __init__(self, words, dev, group, use_multicast=True):
        import torch
        import torch.distributed._symmetric_memory as symm
        self.words = int(words)
        import torch.distributed as dist
        self.flag_words = max(64, int(dist.get_world_size(group)) * self.FLAG_PITCH)
        self.buf = symm.empty(2 * self.words + self.flag_words, dtype=torch.int64, device=dev)
        self.buf.zero_()
        self.hdl = symm.rendezvous(self.buf, group)
        self.world = int(self.hdl.world_size)
        self.rank = int(self.hdl.rank)
        base = [int(p) for p in self.hdl.buffer_ptrs]
        self._ptrs = [torch.tensor([b + h * self.words * 8 for b in base], dtype=torch.int64, device=dev) for h in (0, 1)]
        self._flag_ptrs = torch.tensor([b + 2 * self.words * 8 for b in base], dtype=torch.int64, device=dev)
        self.flags_local = base[self.rank] + 2 * self.words * 8
        mc = 0
        if use_multicast:
            try:
                mc = int(self.hdl.multicast_ptr or 0)   # 0 / None: the group has no NVSwitch multicast mapping
            except Exception:                           # noqa: BLE001 - no multicast: loads per rank
                mc = 0
        self.multicast = mc
        # device-resident pass counters: [0] passes published, [1] passes consumed (uint64), and the publish ticket
        self.ctr = torch.zeros(4, dtype=torch.int64, device=dev)
        self.ticket = torch.zeros(4, dtype=torch.int32, device=dev)
        self.half = 0
        torch.cuda.synchronize(dev)
        self.hdl.barrier(channel=0)                     # every rank's flags are zero before anyone publishes
        torch.cuda.synchronize(dev)

    @classmethod
    def get(cls, comm, words, dev):
        import os
        import torch
        if comm.world <= 1 or dev.type != "cuda":
            return None
        dist = comm._dist
        group = comm.group if comm.group is not None else dist.group.WORLD
        key = (group.group_name, dev.index)
        ex = cls._cache.get(key)
        ok = 1
        mode = os.environ.get("LEIDA_EXCHANGE", "auto")
        if mode == "nccl":
            ok = 0
        elif ex is None or ex.words < words:
            try:
                ex = cls(max(int(words), 1 << 19), dev, group, use_multicast=(mode != "peer"))
            except Exception as e:                      # noqa: BLE001 - any failure means "use NCCL"
                import warnings
                warnings.warn(f"symmetric-memory exchange unavailable ({type(e).__name__}: {e}); using the NCCL all-reduce")
                ex, ok = None, 0
        flag = torch.tensor([ok, 1 if (ex is not None and ex.multicast) else 0], dtype=torch.int32, device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=comm.group)
        if int(flag[0].item()) == 0:
            return None
        if int(flag[1].item()) == 0:
            ex.multicast = 0                            # all ranks read the same way
        cls._cache[key] = ex
        return ex

    def next_half(self):
        """(pointer to this rank's half of the pass, device array of every rank's pointer to that half, multicast
        address of the half or 0); alternates."""
        h = self.half
        self.half ^= 1
        off = h * self.words * 8
        return self.buf.data_ptr() + off, self._ptrs[h], (self.multicast + off) if self.multicast else 0
