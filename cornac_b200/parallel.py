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


def shard_ratings_by_user(rid, n_users, world_size):
    """User ranges with (almost) equal rating counts for MF: `bounds` as in shard_users_by_nnz (rank r owns the users
    [bounds[r], bounds[r+1]) and every rating of theirs)."""
    counts = np.bincount(np.asarray(rid), minlength=int(n_users))
    indptr = np.concatenate([[0], np.cumsum(counts)])
    return shard_users_by_nnz(indptr, world_size)


def shard_ratings(rid, cid, val, bounds, rank):
    """The ratings of `rank`'s users in their STORED order (the order backend_cpu.fit_sgd applies them in), with
    user ids rebased to the shard (item ids stay global)."""
    rid = np.asarray(rid)
    lo, hi = int(bounds[rank]), int(bounds[rank + 1])
    keep = (rid >= lo) & (rid < hi)
    return (rid[keep] - lo), np.asarray(cid)[keep], np.asarray(val)[keep]


class _DeviceMf:
    """The device operations mf_fit_sharded needs (the test seam: tests/test_parallel_cpu.py drives the same
    function over gloo with a CPU stand-in)."""

    def __init__(self):
        import torch
        from . import engine
        engine.require_cuda()
        self.engine, self.torch = engine, torch
        self.ops = None                                  # ItemReplicaSync default: the CUDA delta kernels

    def ids(self, a):
        return self.engine.to_device(np.asarray(a), self.torch.int32)

    def f32(self, a):
        return self.engine.to_device(np.ascontiguousarray(a, dtype=np.float32), self.torch.float32)

    def zeros1(self):
        return self.torch.zeros(1, dtype=self.torch.float32, device="cuda")

    def epoch(self, rid, cid, val, U, V, Bu, Bi, lr, reg, mu, use_bias, loss, atomic):
        self.engine.mf_epoch(rid, cid, val, U, V, Bu, Bi, lr, reg, mu, use_bias, loss, ordered=False, atomic=atomic)

    def to_host(self, host, dev):
        host[...] = dev.cpu().numpy()


def mf_fit_sharded(rid, cid, val, U, V, Bu, Bi, lr, reg, mu, use_bias, max_iter, early_stop=False, atomic=True,
                   _device=None):
    """Multi-GPU MF training of ONE model: call from every rank (one process per GPU, process
    group initialised, torch.cuda.set_device done) with the FULL host rating list and parameter arrays.

    Rank r owns the users [bounds[r], bounds[r+1]) -- equal rating counts -- and all their ratings: U and Bu rows of
    those users are trained locally and never communicated; V and Bi are replicated and, once per epoch, every
    replica becomes  start + sum over ranks of the local changes  (one all-reduce of the deltas).  The epoch loss
    is all-reduced too, so `early_stop` takes the same decision on every rank (backend_cpu.pyx:85-93).
    On return V, Bi hold the shared result on every rank; U[lo:hi], Bu[lo:hi] hold this rank's rows.
    Returns (bounds, losses)."""
    import torch
    import torch.distributed as dist
    dev = _device or _DeviceMf()
    active = dist.is_available() and dist.is_initialized()
    world = dist.get_world_size() if active else 1
    rank = dist.get_rank() if active else 0
    bounds = shard_ratings_by_user(rid, U.shape[0], world)
    lo, hi = int(bounds[rank]), int(bounds[rank + 1])
    r_s, c_s, v_s = shard_ratings(rid, cid, val, bounds, rank)
    d_rid, d_cid, d_val = dev.ids(r_s), dev.ids(c_s), dev.f32(v_s)
    dU, dBu = dev.f32(U[lo:hi]), dev.f32(Bu[lo:hi])
    dV, dBi = dev.f32(V), dev.f32(Bi)
    sync = ItemReplicaSync([dV, dBi], ops=dev.ops) if world > 1 else None
    loss_dev = dev.zeros1()
    lr32, reg32 = float(np.float32(lr)), float(np.float32(reg))
    losses, loss = [], np.float32(0)
    for _ in range(int(max_iter)):
        last = loss
        if len(v_s):
            dev.epoch(d_rid, d_cid, d_val, dU, dV, dBu, dBi, lr32, reg32, float(mu), use_bias, loss_dev, atomic)
        else:
            loss_dev.zero_()
        if sync is not None:
            sync.exchange()
            dist.all_reduce(loss_dev, op=dist.ReduceOp.SUM)
        loss = np.float32(0.5) * np.float32(loss_dev.item())
        losses.append(float(loss))
        if early_stop and abs(np.float32(loss - last)) < 1e-5:
            break
    dev.to_host(U[lo:hi], dU)
    dev.to_host(V, dV)
    if use_bias:
        dev.to_host(Bu[lo:hi], dBu)
        dev.to_host(Bi, dBi)
    return bounds, losses
