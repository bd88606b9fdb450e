"""Multi-GPU plumbing: user sharding + the per-epoch item-replica exchange.

The reference is a single process: nothing here has a reference
counterpart.  Users are cut into contiguous ranges balanced by interaction count; every
rank trains its range against a full replica of V / B; at the epoch boundary the ranks
exchange what they changed:   V <- V_start + sum_r (V_r - V_start)   (one NCCL all-reduce
of the deltas over NVLink; U is never communicated).  One process per GPU, torch.distributed
for the collective, our own kernels (b200_delta_make / b200_delta_apply) around it.
"""
import numpy as np


def shard_users_by_nnz(indptr, world_size):
    """Contiguous user ranges with (almost) equal numbers of interactions.

    Returns an int64 array `bounds` of length world_size+1; rank r owns users
    [bounds[r], bounds[r+1]).  Preserves the sampling law of BPR: a positive is uniform over
    interactions, so equal nnz per rank = equal work per rank."""
    indptr = np.asarray(indptr)
    n_users = len(indptr) - 1
    nnz = int(indptr[-1])
    targets = (np.arange(1, world_size) * nnz) // world_size
    cuts = np.searchsorted(indptr, targets, side="left")
    bounds = np.concatenate([[0], cuts, [n_users]]).astype(np.int64)
    return np.maximum.accumulate(bounds)


def shard_csr(indptr, indices, bounds, rank):
    """The CSR rows of `rank` with row offsets rebased to 0 (item ids stay global)."""
    lo, hi = int(bounds[rank]), int(bounds[rank + 1])
    a, b = int(indptr[lo]), int(indptr[hi])
    return (np.asarray(indptr[lo:hi + 1]) - a).astype(np.int32), np.asarray(indices[a:b])


class _CudaDeltaOps:
    @staticmethod
    def make(x, snapshot, delta):
        from . import engine
        engine.delta_make(x, snapshot, delta)

    @staticmethod
    def apply(x, snapshot, delta):
        from . import engine
        engine.delta_apply(x, snapshot, delta)


class ItemReplicaSync:
    """Keeps the replicas of one or more item-side tensors (V, B) consistent across ranks.

        sync = ItemReplicaSync([V, B])          # after the replicas were initialised identically
        for epoch ...:
            run_local_epoch()
            sync.exchange()                      # V, B now hold start + sum of all ranks' changes

    `ops` is the pair of element-wise kernels (defaults to the CUDA ones); tests inject a
    torch-CPU stand-in to exercise the collective logic over gloo."""

    def __init__(self, tensors, group=None, ops=None):
        import torch
        self.tensors = list(tensors)
        self.group = group
        self.ops = ops or _CudaDeltaOps
        self.snapshots = [t.clone() for t in self.tensors]
        self.deltas = [torch.empty_like(t) for t in self.tensors]

    def exchange(self):
        import torch.distributed as dist
        single = not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(self.group) == 1
        for t, s, d in zip(self.tensors, self.snapshots, self.deltas):
            self.ops.make(t, s, d)
        if not single:
            for d in self.deltas:
                dist.all_reduce(d, op=dist.ReduceOp.SUM, group=self.group)
        for t, s, d in zip(self.tensors, self.snapshots, self.deltas):
            self.ops.apply(t, s, d)


def bpr_fit_sharded(indptr, indices, n_items, U, V, B, lr, reg, use_bias, max_iter, key=0, atomic=True):
    """Multi-GPU BPR training of ONE model: call from every rank (one process per GPU, process
    group initialised, torch.cuda.set_device done) with the FULL host CSR matrix and factor arrays.

    Rank r trains users [bounds[r], bounds[r+1]) (equal interaction counts) against its replica of
    V / B; item-side changes are all-reduced once per epoch.  On return every rank's V and B hold
    the shared result and U[bounds[r]:bounds[r+1]] holds this rank's trained rows (other rows of U
    are untouched on this rank; gather them with torch.distributed if one process needs all)."""
    import torch.distributed as dist
    from . import engine
    world = dist.get_world_size() if dist.is_initialized() else 1
    rank = dist.get_rank() if dist.is_initialized() else 0
    bounds = shard_users_by_nnz(indptr, world)
    ip, ix = shard_csr(indptr, indices, bounds, rank)
    lo, hi = int(bounds[rank]), int(bounds[rank + 1])
    U_shard = U[lo:hi]                      # a view: trained in place
    hist, _ = engine.bpr_train_host(ip, ix, n_items, U_shard, V, B, lr, reg, use_bias, max_iter,
                                    key=(int(key) << 8) + rank, atomic=atomic, replica_sync=world > 1,
                                    on_epoch=lambda *a: None)
    return bounds, hist
