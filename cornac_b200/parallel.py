"""Multi-GPU plumbing: user sharding + the per-epoch item-replica exchange.

The reference is a single process: nothing here has a reference
counterpart.  Users are cut into contiguous ranges balanced by interaction count; every
rank trains its range against a full replica of V / B; at the epoch boundary the ranks
exchange what they changed:   V <- V_start + mean over the ranks that changed the row of
(V_r - V_start)   (U is never communicated).  One process per GPU; the exchange is one fused
kernel over NVLink peer memory (PeerItemExchange, csrc/p2p.cu) or our delta kernels around a
torch.distributed all-reduce (ItemReplicaSync).

Why a mean and not the sum (SURVEY 8(e) proposed the sum "to match the single-GPU step count"): the sum is right
only while the ranks change DIFFERENT rows.  A popular item is trained by every rank; each rank alone moves it
most of the way to its local optimum within the epoch, so the sum overshoots `world`-fold and the epochs oscillate
with growing amplitude from 4 ranks on -- tools/sim_localsgd.py reproduces it on the CPU with the reference loop
(pairwise accuracy 0.81 -> 0.30 at 4 ranks, |B| 3 -> 1e5 at 8), and on the GPU box the 8-rank MF check went to NaN
and the 8-rank bench model broke the rank leg.  The mean over the ranks that changed an element is a convex
combination of their results: stable at any world size, identical to the sum where one rank alone touched the
row (reduce="sum" keeps the old rule for A/B runs).
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
            sync.exchange()                      # V, B now hold start + the ranks' changes (mean over the ranks that changed a row)

    `ops` is the pair of element-wise kernels (defaults to the CUDA ones); tests inject a
    torch-CPU stand-in to exercise the collective logic over gloo."""

    def __init__(self, tensors, group=None, ops=None, reduce="mean_touched"):
        import torch
        if reduce not in ("mean_touched", "sum"):
            raise ValueError("reduce must be 'mean_touched' or 'sum'")
        self.tensors = list(tensors)
        self.group = group
        self.ops = ops or _CudaDeltaOps
        self.reduce = reduce
        self.snapshots = [t.clone() for t in self.tensors]
        self.deltas = [torch.empty_like(t) for t in self.tensors]

    def exchange(self):
        import torch.distributed as dist
        single = not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(self.group) == 1
        for t, s, d in zip(self.tensors, self.snapshots, self.deltas):
            self.ops.make(t, s, d)
        if not single:
            for d in self.deltas:
                if self.reduce == "mean_touched":
                    touched = (d != 0).to(d.dtype)              # which elements this rank changed
                    dist.all_reduce(touched, op=dist.ReduceOp.SUM, group=self.group)
                dist.all_reduce(d, op=dist.ReduceOp.SUM, group=self.group)
                if self.reduce == "mean_touched":
                    d.div_(touched.clamp_(min=1))
        for t, s, d in zip(self.tensors, self.snapshots, self.deltas):
            self.ops.apply(t, s, d)


class PeerItemExchange:
    """ItemReplicaSync's exchange as ONE kernel per tensor over NVLink peer memory (b200_item_exchange, csrc/p2p.cu):
    the ranks map each other's replicas with CUDA IPC at construction (handles travel through torch.distributed); every
    exchange() then launches, per tensor, a kernel in which this rank reduces and rewrites its slice of every replica.
    No NCCL collective and no delta buffers on the data path; the result is V <- V_start + mean over the ranks that changed the
    element of (V_r - V_start) (reduce="sum": the plain sum), formed in rank order, bit-equal on all ranks.  One process per GPU, all GPUs of one NVLink domain."""

    def __init__(self, tensors, group=None, reduce="mean_touched"):
        import ctypes
        import torch
        import torch.distributed as dist
        from . import _lib
        if reduce not in ("mean_touched", "sum"):
            raise ValueError("reduce must be 'mean_touched' or 'sum'")
        self.reduce = reduce
        self._L = _lib.load()
        self._check = _lib.check
        self.group = group
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)
        if self.world > 8:
            raise _lib.B200Error("PeerItemExchange supports up to 8 ranks (one NVLink domain)")
        self.tensors = list(tensors)
        self.seq = 0
        self._state = []
        self._opened = {}                               # (peer rank, IPC handle) -> mapped base: an allocation is opened once
        for t in self.tensors:
            if not (t.is_cuda and t.dtype == torch.float32 and t.is_contiguous()):
                raise _lib.B200Error("PeerItemExchange needs contiguous float32 CUDA tensors")
            flags = torch.zeros(32, dtype=torch.int32, device=t.device)
            lo, hi = ctypes.c_int64(), ctypes.c_int64()
            self._check(self._L.b200_item_exchange_slice(self.rank, self.world, t.numel(), ctypes.byref(lo), ctypes.byref(hi)),
                        "b200_item_exchange_slice")
            snap = t.view(-1)[lo.value:hi.value].clone() if hi.value > lo.value else torch.zeros(4, device=t.device)
            mine = []
            for buf in (t, flags):
                h = (ctypes.c_ubyte * 64)()
                off = ctypes.c_int64()
                self._check(self._L.b200_ipc_export(buf.data_ptr(), h, ctypes.byref(off)), "b200_ipc_export")
                mine.append((bytes(h), off.value))
            everyone = [None] * self.world
            dist.all_gather_object(everyone, mine, group=group)
            x_ptrs, f_ptrs = (ctypes.c_void_p * self.world)(), (ctypes.c_void_p * self.world)()
            for r in range(self.world):
                if r == self.rank:
                    x_ptrs[r], f_ptrs[r] = t.data_ptr(), flags.data_ptr()
                    continue
                for slot, (h, off) in zip((x_ptrs, f_ptrs), everyone[r]):
                    if (r, h) not in self._opened:
                        out = ctypes.c_void_p()
                        self._check(self._L.b200_ipc_open(h, 0, ctypes.byref(out)), "b200_ipc_open")
                        self._opened[(r, h)] = out.value
                    slot[r] = self._opened[(r, h)] + off
            self._state.append(dict(t=t, flags=flags, snap=snap, x_ptrs=x_ptrs, f_ptrs=f_ptrs))
        torch.cuda.synchronize()
        dist.barrier(group=group)                       # every flag buffer is zeroed and mapped before the first exchange

    def exchange(self):
        from ._lib import current_stream
        self.seq += 1
        for st in self._state:
            self._check(self._L.b200_item_exchange(self.rank, self.world, st["x_ptrs"], st["f_ptrs"], st["snap"].data_ptr(),
                                                   st["t"].numel(), self.seq, int(self.reduce == "mean_touched"), current_stream()),
                        "b200_item_exchange")

    def failed(self):
        """True when a peer did not show up in some exchange (bounded wait expired); synchronises."""
        return any(int(st["flags"][17].item()) != 0 for st in self._state)

    def close(self):
        """Unmap the peers' allocations (after the last exchange has completed on every rank)."""
        import torch.distributed as dist
        if self._opened:
            dist.barrier(group=self.group)
            for base in self._opened.values():
                self._L.b200_ipc_close(base, 0)
            self._opened = {}


def make_item_sync(tensors, group=None, kind="auto", reduce="mean_touched"):
    """The per-epoch item exchange for `tensors` ([V, B]): 'p2p' = PeerItemExchange (one fused NVLink kernel per tensor),
    'nccl' = ItemReplicaSync (delta kernels around a torch.distributed all-reduce; also the gloo / CPU test seam),
    'auto' = p2p on CUDA tensors with the NCCL backend, else nccl."""
    import torch.distributed as dist
    if kind == "auto":
        on_gpu = all(getattr(t, "is_cuda", False) for t in tensors)
        kind = "p2p" if (on_gpu and dist.is_initialized() and dist.get_backend(group) == "nccl") else "nccl"
    if kind == "p2p":
        return PeerItemExchange(tensors, group=group, reduce=reduce)
    return ItemReplicaSync(tensors, group=group, reduce=reduce)


def bpr_fit_sharded(indptr, indices, n_items, U, V, B, lr, reg, use_bias, max_iter, key=0, atomic=True):
    """Multi-GPU BPR training of ONE model: call from every rank (one process per GPU, process
    group initialised, torch.cuda.set_device done) with the FULL host CSR matrix and factor arrays.

    Rank r trains users [bounds[r], bounds[r+1]) (equal interaction counts) against its replica of
    V / B; item-side changes are exchanged once per epoch (make_item_sync: mean over the ranks that changed a row).  On return every rank's V and B hold
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
    replica becomes  start + the ranks' local changes averaged over the ranks that changed the element  (make_item_sync).  The epoch loss
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
    sync = (ItemReplicaSync([dV, dBi], ops=dev.ops) if dev.ops is not None else make_item_sync([dV, dBi])) if world > 1 else None
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
