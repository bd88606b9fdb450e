"""Device-side engine: thin, typed Python wrappers over the C ABI.

Everything here operates on torch CUDA tensors used as plain device buffers.  The
functions mirror the reference kernels one to one (see include/b200cornac.h):
    bpr_epoch / bpr_epoch_replay   <-> BPR._fit_sgd          (cornac/models/bpr/recom_bpr.pyx:208-269)
    MTSampler                      <-> RNGVector             (cornac/models/bpr/recom_bpr.pyx:54-62)
    mf_epoch                       <-> backend_cpu.fit_sgd   (cornac/models/mf/backend_cpu.pyx:58-83)
    score_batch                    <-> fast_dot              (cornac/utils/fast_dot.pyx:40-43)
    topk_rows / rank_topk          <-> Recommender.rank      (cornac/models/recommender.py:476-530)
"""
import numpy as np
import torch

from . import _lib
from ._lib import B200Error, check, current_stream, ptr


def require_cuda():
    if not torch.cuda.is_available():
        raise B200Error("cornac_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
    return _lib.load()


def warmup():
    """Create the CUDA context and load libb200cornac.so now instead of inside the first fit()/rank() (a fresh process
    pays seconds for the context; an experiment that times its models should not charge that to whichever runs first)."""
    L = require_cuda()
    torch.zeros(1, device="cuda")
    torch.cuda.synchronize()
    return L


def _dev(t, dtype, name):
    if not (isinstance(t, torch.Tensor) and t.is_cuda and t.dtype == dtype and t.is_contiguous()):
        raise B200Error("%s must be a contiguous CUDA tensor of dtype %s" % (name, dtype))
    return t


def to_device(a, dtype=None, pinned=True):
    """H2D copy of a numpy array (through pinned memory) -> CUDA tensor."""
    t = torch.from_numpy(np.ascontiguousarray(a))
    if dtype is not None and t.dtype != dtype:
        t = t.to(dtype)
    if pinned:
        t = t.pin_memory()
    return t.cuda(non_blocking=True)


class BprData:
    """Device copy of train_set.matrix: the CSR arrays (indptr int32 [n_users+1], sorted indices
    int32 [nnz]) + the COO row array (BPR._prepare_data, recom_bpr.pyx:154-161) used by the
    parity kernel, and -- built on first use by b200_bpr_prepare -- the (pairs, membership
    table) store the throughput kernel gathers from."""

    def __init__(self, indptr, indices, coo_row=None):
        self.indptr = _dev(indptr, torch.int32, "indptr")
        self.indices = _dev(indices, torch.int32, "indices")
        self.n_users = self.indptr.numel() - 1
        self.nnz = self.indices.numel()
        self._coo_row = None if coo_row is None else _dev(coo_row, torch.int32, "coo_row")
        if self._coo_row is not None and self._coo_row.numel() != self.nnz:
            raise B200Error("coo_row length %d != nnz %d" % (self._coo_row.numel(), self.nnz))
        self.pairs = self.table = None

    @property
    def coo_row(self):
        if self._coo_row is None:
            counts = (self.indptr[1:] - self.indptr[:-1]).to(torch.int64)
            self._coo_row = torch.repeat_interleave(
                torch.arange(self.n_users, device=self.indptr.device, dtype=torch.int32), counts)
        return self._coo_row

    def prepare(self):
        """Build pairs + membership table (idempotent)."""
        if self.pairs is None:
            L = require_cuda()
            slots = int(L.b200_bpr_table_slots(self.nnz))
            pairs = torch.empty((max(self.nnz, 1), 2), dtype=torch.int32, device=self.indptr.device)
            table = torch.empty(slots, dtype=torch.int64, device=self.indptr.device)
            check(L.b200_bpr_prepare(ptr(self.indptr), ptr(self.indices), self.n_users, self.nnz, ptr(pairs),
                                     ptr(table), slots, current_stream()), "b200_bpr_prepare")
            self.pairs, self.table = pairs, table
        return self

    @classmethod
    def from_host(cls, indptr, indices):
        if len(indices) >= 2 ** 31:
            raise B200Error("nnz >= 2^31 per shard is not supported (int32 CSR offsets)")
        return cls(to_device(np.asarray(indptr), torch.int32), to_device(np.asarray(indices), torch.int32))


def bpr_epoch(data, n_neg, U, V, B, lr, reg, use_bias, seed, epoch, stats, n_samples=None,
              sample_base=0, atomic=True, exact_exp=False, unbounded=False, neg_weighted=False, hinge=False, blocked=False):
    """One Hogwild BPR epoch on the current stream; `stats` (int64[2] CUDA) accumulates
    (correct, skipped)."""
    L = require_cuda()
    k = U.shape[1]
    _dev(U, torch.float32, "U"), _dev(V, torch.float32, "V"), _dev(B, torch.float32, "B")
    _dev(stats, torch.int64, "stats")
    flags = ((_lib.SGD_ATOMIC if atomic else 0) | (_lib.SGD_EXACT_EXP if exact_exp else 0)
             | (_lib.SGD_UNBOUNDED if unbounded else 0) | (_lib.BPR_NEG_WEIGHTED if neg_weighted else 0)
             | (_lib.BPR_LOSS_HINGE if hinge else 0) | (_lib.BPR_BLOCKED if blocked else 0))
    n = data.nnz if n_samples is None else int(n_samples)
    data.prepare()
    check(L.b200_bpr_epoch(ptr(data.pairs), ptr(data.table), data.table.numel(), data.nnz, data.n_users, int(n_neg), n,
                           ptr(U), ptr(V), ptr(B), int(k), float(lr), float(reg), int(bool(use_bias)),
                           int(seed) & (2 ** 64 - 1), int(epoch), int(sample_base), flags, ptr(stats),
                           current_stream()), "b200_bpr_epoch")


def bpr_block_plan(n_users, n_neg, k):
    """(windows of the interaction list, item blocks) of the cache-blocked sample order for this model size."""
    import ctypes
    L = _lib.load()
    a, b = ctypes.c_uint32(), ctypes.c_uint32()
    check(L.b200_bpr_block_plan(int(n_users), int(n_neg), int(k), ctypes.byref(a), ctypes.byref(b)), "b200_bpr_block_plan")
    return a.value, b.value


def bpr_draw_host(seed, epoch, n, nnz, n_neg, sample_base=0, plan=(1, 1)):
    """The (i_index, j_id) stream that bpr_epoch(seed, epoch) consumes, computed on the host (plan = bpr_block_plan(...)
    for an epoch run with blocked=True)."""
    L = _lib.load()
    ii = np.empty(n, dtype=np.int64)
    jj = np.empty(n, dtype=np.int32)
    check(L.b200_bpr_draw_host2(int(seed) & (2 ** 64 - 1), int(epoch), int(sample_base), int(n), int(nnz), int(n_neg),
                                int(plan[0]), int(plan[1]), ii.ctypes.data, jj.ctypes.data), "b200_bpr_draw_host")
    return ii, jj


def bpr_epoch_replay(data, i_index, j_id, U, V, B, lr, reg, use_bias, stats, hinge=False):
    """Serial-equivalent application of an explicit sample stream (parity mode)."""
    L = require_cuda()
    _dev(i_index, torch.int64, "i_index"), _dev(j_id, torch.int32, "j_id")
    _dev(U, torch.float32, "U"), _dev(V, torch.float32, "V"), _dev(B, torch.float32, "B")
    check(L.b200_bpr_epoch_replay2(ptr(i_index), ptr(j_id), i_index.numel(), ptr(data.indptr), ptr(data.indices),
                                   ptr(data.coo_row), int(U.shape[0]), int(V.shape[0]), ptr(U), ptr(V), ptr(B), int(U.shape[1]),
                                   float(lr), float(reg), int(bool(use_bias)), _lib.BPR_LOSS_HINGE if hinge else 0,
                                   ptr(_dev(stats, torch.int64, "stats")), current_stream()),
          "b200_bpr_epoch_replay")


def bpr_train_host(indptr, indices, n_neg, U, V, B, lr, reg, use_bias, max_iter, key=0, replay_seeds=None,
                   atomic=True, on_epoch=None, keep_device=False, replica_sync=False, weighted_seed=None,
                   neg_weighted=False, hinge=False, blocked=True):
    """Host-buffer entry of BPR training (what BPR.fit calls): uploads the CSR matrix and the
    factors, runs `max_iter` epochs, writes the trained factors back INTO the given numpy
    arrays U, V, B (pinned staging both ways).

    replay_seeds = (seed_pos, seed_neg): deterministic mode -- per epoch the two mt19937 streams
    of the reference's RNGVector are drawn on the host and applied by the serial-equivalent
    replay kernel.  Otherwise Hogwild epochs with the on-device Philox sampler keyed by `key`.
    replica_sync=True (multi-GPU, one process per GPU, torch.distributed initialised): `indptr/indices/U`
    are this rank's USER SHARD, V/B are replicas; after every epoch the ranks exchange their item-side
    changes (parallel.ItemReplicaSync: make-delta -> NCCL all-reduce -> apply).
    Returns (per-epoch (correct, skipped) list or [], device tensors (U, V, B) if keep_device)."""
    require_cuda()
    data = BprData.from_host(indptr, indices)
    nnz = data.nnz
    if replay_seeds is None and weighted_seed is None:
        data.prepare()                  # pair store + membership table: runs while the factors are still copying
    # the factor matrices go up on a side stream so that the copy engine overlaps b200_bpr_prepare
    main = torch.cuda.current_stream()
    side = _copy_stream()               # no dependency on `main`: the uploads may start right away
    with torch.cuda.stream(side):
        dU, dV, dB = to_device(U, torch.float32), to_device(V, torch.float32), to_device(B, torch.float32)
    main.wait_stream(side)
    for t in (dU, dV, dB):
        t.record_stream(main)
    stats = torch.zeros(2, dtype=torch.int64, device="cuda")
    lr, reg = float(np.float32(lr)), float(np.float32(reg))
    history = []
    sync = None
    if replica_sync:
        from .parallel import make_item_sync
        sync = make_item_sync([dV, dB])
    if weighted_seed is not None or replay_seeds is not None:
        # Deterministic mode.  BPR: the two mt19937 streams of the reference's RNGVector (recom_bpr.pyx:54-62).  WBPR:
        # ONE stream, each sample takes (pos draw, neg draw) from it and the negative is the item of the drawn
        # interaction (recom_wbpr.pyx:125-136).  The epochs are PIPELINED: the host draws epoch e + 1 while the GPU
        # applies epoch e (three pinned staging sets, fenced by events; per-epoch stats stay on the device and come
        # back once at the end), unless a per-epoch callback needs the numbers right away.
        n_sets = 3 if nnz <= (1 << 24) else 2
        if weighted_seed is not None:
            g = MTSampler(weighted_seed)
            h_ij = np.empty(2 * nnz, dtype=np.int64)
            host_indices = np.asarray(indices)
        else:
            g_pos, g_neg = MTSampler(replay_seeds[0]), MTSampler(replay_seeds[1])
        h_i = [torch.empty(nnz, dtype=torch.int64).pin_memory() for _ in range(n_sets)]
        h_j = [torch.empty(nnz, dtype=torch.int32).pin_memory() for _ in range(n_sets)]
        d_i = [torch.empty(nnz, dtype=torch.int64, device="cuda") for _ in range(n_sets)]
        d_j = [torch.empty(nnz, dtype=torch.int32, device="cuda") for _ in range(n_sets)]
        fence = [None] * n_sets
        stats_all = torch.zeros((max_iter, 2), dtype=torch.int64, device="cuda")
        for epoch in range(max_iter):
            b = epoch % n_sets
            if fence[b] is not None:
                fence[b].synchronize()          # the upload of the epoch that last used this staging set has finished
            if weighted_seed is not None:
                g.fill(nnz - 1, 2 * nnz, out=h_ij)
                h_i[b].numpy()[:] = h_ij[0::2]
                h_j[b].numpy()[:] = host_indices[h_ij[1::2]]
            else:
                g_pos.fill(nnz - 1, nnz, out=h_i[b].numpy())
                g_neg.fill(int(n_neg) - 1, nnz, out=h_j[b].numpy())
            d_i[b].copy_(h_i[b], non_blocking=True)
            d_j[b].copy_(h_j[b], non_blocking=True)
            fence[b] = torch.cuda.Event()
            fence[b].record()
            bpr_epoch_replay(data, d_i[b], d_j[b], dU, dV, dB, lr, reg, use_bias, stats_all[epoch], hinge=hinge)
            if on_epoch:
                on_epoch(epoch, *stats_all[epoch].cpu().tolist())
        history = [tuple(r) for r in stats_all.cpu().tolist()]
    else:
        for epoch in range(max_iter):
            stats.zero_()
            bpr_epoch(data, n_neg, dU, dV, dB, lr, reg, use_bias, key, epoch, stats, atomic=atomic,
                      neg_weighted=neg_weighted, hinge=hinge, blocked=blocked)
            if sync is not None:
                sync.exchange()
            if on_epoch:
                history.append(tuple(stats.cpu().tolist()))
                on_epoch(epoch, *history[-1])
    if sync is not None and hasattr(sync, "close"):
        torch.cuda.synchronize()
        sync.close()
    for host, dev in ((U, dU), (V, dV), (B, dB)):
        _to_host_into(host, dev)
    return history, ((dU, dV, dB) if keep_device else None)


def tri_train_host(kind, indptr, indices, aux, n_items, U, V, B, hyper, max_iter, key=0, replay_seeds=None, on_epoch=None,
                   keep_device=False):
    """Host-buffer entry of the BPR siblings with a third item per sample (csrc/bprx.cu), what VEBPR.fit / SBPR.fit call.

    kind "vebpr": aux = (view_indptr, view_indices) of the viewed-not-purchased CSR, hyper = dict(lr, reg, alpha), B = None,
                  replay_seeds = (pos, view, neg) mt19937 seeds of the three RNGVectors (recom_vebpr.pyx:198-200);
    kind "sbpr":  aux = (social_indptr, social_item_ids, social_item_counts), hyper = dict(lr, lambda_u, lambda_v, lambda_b,
                  use_bias), replay_seeds = (pos, neg) (recom_sbpr.pyx:173-174).
    replay_seeds given -> the seeded streams are drawn on the host in the reference's order (b200_*_draw_host) and applied
    by the serial-equivalent replay kernel, the host drawing epoch e + 1 while the GPU applies epoch e; else Hogwild epochs
    with on-device Philox sampling keyed by `key`.  Trained factors are written back INTO U, V (, B).
    Returns (per-epoch (correct, skipped) list, device tensors (U, V, B or None) if keep_device)."""
    L = require_cuda()
    assert kind in ("vebpr", "sbpr")
    data = BprData.from_host(indptr, indices)
    nnz = data.nnz
    coo = data.coo_row
    aux_dev = [to_device(np.ascontiguousarray(a, dtype=np.int32), torch.int32) if len(a) else
               torch.zeros(1, dtype=torch.int32, device="cuda") for a in aux]
    dU, dV = to_device(U, torch.float32), to_device(V, torch.float32)
    dB = to_device(B, torch.float32) if B is not None else None
    n_users, k = int(dU.shape[0]), int(dU.shape[1])
    max_iter = int(max_iter)
    stats_all = torch.zeros((max(max_iter, 1), 2), dtype=torch.int64, device="cuda")
    f32 = lambda x: float(np.float32(x))          # noqa: E731
    st = current_stream

    def hogwild(epoch):
        if kind == "vebpr":
            check(L.b200_vebpr_epoch(ptr(data.indptr), ptr(data.indices), ptr(coo), n_users, int(n_items), nnz,
                                     ptr(aux_dev[0]), ptr(aux_dev[1]), ptr(dU), ptr(dV), k, f32(hyper["lr"]), f32(hyper["reg"]),
                                     f32(hyper["alpha"]), int(key) & ((1 << 64) - 1), epoch, nnz, ptr(stats_all[epoch]), st()),
                  "b200_vebpr_epoch")
        else:
            check(L.b200_sbpr_epoch(ptr(data.indptr), ptr(data.indices), ptr(coo), n_users, int(n_items), nnz,
                                    ptr(aux_dev[0]), ptr(aux_dev[1]), ptr(aux_dev[2]), len(aux[1]), ptr(dU), ptr(dV), ptr(dB), k,
                                    f32(hyper["lr"]), f32(hyper["lambda_u"]), f32(hyper["lambda_v"]), f32(hyper["lambda_b"]),
                                    int(bool(hyper["use_bias"])), int(key) & ((1 << 64) - 1), epoch, nnz, ptr(stats_all[epoch]), st()),
                  "b200_sbpr_epoch")

    if replay_seeds is None:
        for epoch in range(max_iter):
            hogwild(epoch)
            if on_epoch:
                on_epoch(epoch, *stats_all[epoch].cpu().tolist())
    else:
        gens = [MTSampler(s_) for s_ in replay_seeds]
        h_coo = np.repeat(np.arange(len(indptr) - 1, dtype=np.int32), np.diff(np.asarray(indptr)).astype(np.int64))
        h_aux = [np.ascontiguousarray(a, dtype=np.int32) for a in aux]
        third = torch.int32 if kind == "vebpr" else torch.int64
        n_sets = 2
        h_i = [torch.empty(nnz, dtype=torch.int64).pin_memory() for _ in range(n_sets)]
        h_j = [torch.empty(nnz, dtype=torch.int32).pin_memory() for _ in range(n_sets)]
        h_t = [torch.empty(nnz, dtype=third).pin_memory() for _ in range(n_sets)]
        d_i = [torch.empty(nnz, dtype=torch.int64, device="cuda") for _ in range(n_sets)]
        d_j = [torch.empty(nnz, dtype=torch.int32, device="cuda") for _ in range(n_sets)]
        d_t = [torch.empty(nnz, dtype=third, device="cuda") for _ in range(n_sets)]
        fence = [None] * n_sets
        for epoch in range(max_iter):
            b = epoch % n_sets
            if fence[b] is not None:
                fence[b].synchronize()
            if kind == "vebpr":
                check(L.b200_vebpr_draw_host(gens[0]._h, gens[1]._h, gens[2]._h, nnz, int(n_items), h_coo.ctypes.data,
                                             h_aux[0].ctypes.data, h_aux[1].ctypes.data, nnz, h_i[b].data_ptr(), h_t[b].data_ptr(),
                                             h_j[b].data_ptr()), "b200_vebpr_draw_host")
            else:
                check(L.b200_sbpr_draw_host(gens[0]._h, gens[1]._h, nnz, int(n_items), h_coo.ctypes.data, h_aux[0].ctypes.data, nnz,
                                            h_i[b].data_ptr(), h_j[b].data_ptr(), h_t[b].data_ptr()), "b200_sbpr_draw_host")
            d_i[b].copy_(h_i[b], non_blocking=True), d_j[b].copy_(h_j[b], non_blocking=True), d_t[b].copy_(h_t[b], non_blocking=True)
            fence[b] = torch.cuda.Event()
            fence[b].record()
            if kind == "vebpr":
                check(L.b200_vebpr_epoch_replay(ptr(d_i[b]), ptr(d_t[b]), ptr(d_j[b]), nnz, ptr(data.indptr), ptr(data.indices), ptr(coo),
                                                ptr(aux_dev[0]), ptr(aux_dev[1]), ptr(dU), ptr(dV), k, f32(hyper["lr"]), f32(hyper["reg"]),
                                                f32(hyper["alpha"]), ptr(stats_all[epoch]), st()), "b200_vebpr_epoch_replay")
            else:
                check(L.b200_sbpr_epoch_replay(ptr(d_i[b]), ptr(d_j[b]), ptr(d_t[b]), nnz, ptr(data.indptr), ptr(data.indices), ptr(coo),
                                               ptr(aux_dev[0]), ptr(aux_dev[1]), ptr(aux_dev[2]), len(aux[1]), ptr(dU), ptr(dV), ptr(dB), k,
                                               f32(hyper["lr"]), f32(hyper["lambda_u"]), f32(hyper["lambda_v"]), f32(hyper["lambda_b"]),
                                               int(bool(hyper["use_bias"])), ptr(stats_all[epoch]), st()), "b200_sbpr_epoch_replay")
            if on_epoch:
                on_epoch(epoch, *stats_all[epoch].cpu().tolist())
    history = [tuple(r) for r in stats_all[:max_iter].cpu().tolist()]
    _to_host_into(U, dU)
    _to_host_into(V, dV)
    if B is not None:
        _to_host_into(B, dB)
    return history, ((dU, dV, dB) if keep_device else None)


_COPY_STREAMS = {}


def _copy_stream():
    dev = torch.cuda.current_device()
    if dev not in _COPY_STREAMS:
        _COPY_STREAMS[dev] = torch.cuda.Stream(device=dev)
    return _COPY_STREAMS[dev]


def _to_host_into(host, dev):
    """D2H into an existing numpy array (through pinned staging when it is not pinned itself)."""
    out = torch.from_numpy(host) if (host.flags["C_CONTIGUOUS"] and host.flags.writeable) else None
    if out is not None and out.dtype == dev.dtype and tuple(out.shape) == tuple(dev.shape):
        out.copy_(dev)
        return host
    raise B200Error("destination array must be a writable C-contiguous %s array of shape %s" % (dev.dtype, tuple(dev.shape)))


class MTSampler:
    """boost::random::mt19937 + uniform_int_distribution<long>(0, hi) on the host
    (RNGVector of the reference, one thread)."""

    def __init__(self, seed):
        self._L = _lib.load()
        self._h = self._L.b200_mt_sampler_create(int(seed) & 0xFFFFFFFF)
        if not self._h:
            raise B200Error("b200_mt_sampler_create failed")

    def fill(self, hi, n, dtype=np.int64, out=None):
        out = np.empty(n, dtype=dtype) if out is None else out
        fn = self._L.b200_mt_sampler_fill_i64 if out.dtype == np.int64 else self._L.b200_mt_sampler_fill_i32
        check(fn(self._h, int(hi), int(n), out.ctypes.data), "b200_mt_sampler_fill")
        return out

    def __del__(self):
        if getattr(self, "_h", None):
            self._L.b200_mt_sampler_destroy(self._h)
            self._h = None


def mf_epoch(rid, cid, val, U, V, Bu, Bi, lr, reg, mu, use_bias, loss, ordered=False, atomic=True, unbounded=False):
    """One MF epoch; `loss` (float32[1] CUDA) receives sum(err^2).  U = V = None: the bias-only model
    (BaselineOnly), sized by Bu / Bi."""
    L = require_cuda()
    if rid.dtype not in (torch.int32, torch.int64) or cid.dtype != rid.dtype:
        raise B200Error("rid/cid must both be int32 or int64")
    _dev(rid, rid.dtype, "rid"), _dev(cid, rid.dtype, "cid"), _dev(val, torch.float32, "val")
    for n_, t_ in (("U", U), ("V", V), ("Bu", Bu), ("Bi", Bi), ("loss", loss)):
        if t_ is not None or n_ not in ("U", "V"):
            _dev(t_, torch.float32, n_)
    if (U is None) != (V is None):
        raise B200Error("U and V must both be given or both be None")
    k = 0 if U is None else int(U.shape[1])
    check(L.b200_mf_epoch(ptr(rid), ptr(cid), ptr(val), val.numel(), int(rid.dtype == torch.int32),
                          int(Bu.shape[0]), int(Bi.shape[0]), ptr(U), ptr(V), ptr(Bu), ptr(Bi), k, float(lr), float(reg), float(mu),
                          int(bool(use_bias)), int(bool(ordered)),
                          (_lib.SGD_ATOMIC if atomic else 0) | (_lib.SGD_UNBOUNDED if unbounded else 0), ptr(loss),
                          current_stream()), "b200_mf_epoch")


class WmfTrainer:
    """Device state of one WMF fit (the variables and Adam slots of the reference's TF-1 graph, wmf/wmf.py:34-55) and
    the step `sess.run([model.opt, model.loss], feed_dict)` of recom_wmf.py:197-199, one mini-batch of item ids at a
    time.  `csc` is train_set.csc_matrix (scipy); U, V numpy float32 arrays (copied up; read back with .download())."""

    BETA1, BETA2, EPS = 0.9, 0.999, 1e-8              # tf.train.AdamOptimizer defaults

    def __init__(self, csc, U, V, a, b, lambda_u, lambda_v, lr):
        require_cuda()
        csc = csc.tocsc()
        csc.sort_indices()
        if csc.nnz >= 2 ** 31:
            raise B200Error("nnz >= 2^31 is not supported (int32 CSC offsets)")
        self.n_users, self.n_items = (int(x) for x in csc.shape)
        if U.shape[0] != self.n_users or V.shape[0] != self.n_items or U.shape[1] != V.shape[1]:
            raise B200Error("U / V shapes %s / %s do not match the %d x %d rating matrix" % (U.shape, V.shape, self.n_users, self.n_items))
        self.k = int(U.shape[1])
        self.indptr = to_device(csc.indptr, torch.int32)
        self.rows = to_device(csc.indices if csc.nnz else np.zeros(1, np.int32), torch.int32)
        self.vals = to_device(np.asarray(csc.data if csc.nnz else np.zeros(1), dtype=np.float32), torch.float32)
        self.U = to_device(np.ascontiguousarray(U, dtype=np.float32), torch.float32)
        self.V = to_device(np.ascontiguousarray(V, dtype=np.float32), torch.float32)
        self.mU, self.vU = torch.zeros_like(self.U), torch.zeros_like(self.U)
        self.mV, self.vV = torch.zeros_like(self.V), torch.zeros_like(self.V)
        self.slot_of = torch.full((self.n_items,), -1, dtype=torch.int32, device="cuda")
        self.loss = torch.zeros(1, dtype=torch.float64, device="cuda")
        self.gV = None
        f32 = np.float32
        self.a, self.b, self.lambda_u, self.lambda_v = f32(a), f32(b), f32(lambda_u), f32(lambda_v)
        self.lr = f32(lr)
        self.b1_pow, self.b2_pow = f32(self.BETA1), f32(self.BETA2)     # beta^t for the step about to be taken (t = 1)

    def step(self, ids, want_loss=True):
        """One optimisation step on the item mini-batch `ids` (distinct item indices); returns the batch loss (float) or
        None.  Reading the loss synchronises with the device, like sess.run does."""
        L = _lib.load()
        ids = np.ascontiguousarray(ids, dtype=np.int32)
        b = int(len(ids))
        if b == 0:
            return 0.0
        if ids.min() < 0 or ids.max() >= self.n_items:
            raise B200Error("item id out of range in the mini-batch")
        d_ids = to_device(ids, torch.int32, pinned=False)
        if self.gV is None or self.gV.numel() < b * self.k:
            self.gV = torch.empty(b * self.k, dtype=torch.float32, device="cuda")
        f32 = np.float32
        lr_t = self.lr * np.sqrt(f32(1) - self.b2_pow) / (f32(1) - self.b1_pow)          # AdamOptimizer._prepare / _apply_*
        check(L.b200_wmf_step(ptr(self.indptr), ptr(self.rows), ptr(self.vals), ptr(d_ids), b, self.n_users, self.n_items,
                              self.k, ptr(self.U), ptr(self.V), ptr(self.mU), ptr(self.vU), ptr(self.mV), ptr(self.vV),
                              float(self.a), float(self.b), float(self.lambda_u), float(self.lambda_v), float(f32(lr_t)),
                              self.BETA1, self.BETA2, self.EPS, ptr(self.slot_of), ptr(self.gV), ptr(self.loss),
                              current_stream()), "b200_wmf_step")
        self.b1_pow = f32(self.b1_pow * f32(self.BETA1))              # _finish: the beta powers advance once per step
        self.b2_pow = f32(self.b2_pow * f32(self.BETA2))
        return float(self.loss.item()) if want_loss else None

    def download(self):
        return self.U.cpu().numpy(), self.V.cpu().numpy()


def score_batch(U, V, user_idx=None, item_base=None, user_off=None, n_items=None, out=None):
    """out[q, i] = (item_base[i] + user_off[q]) + dot(U[user_idx[q]], V[i]) for i < n_items."""
    L = require_cuda()
    _dev(U, torch.float32, "U"), _dev(V, torch.float32, "V")
    n_items = V.shape[0] if n_items is None else int(n_items)
    n_q = U.shape[0] if user_idx is None else user_idx.numel()
    if user_idx is not None:
        _dev(user_idx, torch.int64, "user_idx")
    if out is None:
        out = torch.empty((n_q, n_items), dtype=torch.float32, device=U.device)
    check(L.b200_score_batch(ptr(U), ptr(user_idx), n_q, ptr(V), n_items, int(V.shape[1]), ptr(item_base),
                             ptr(user_off), ptr(_dev(out, torch.float32, "out")), current_stream()),
          "b200_score_batch")
    return out


def topk_rows(scores, topk, excl_indptr=None, excl_indices=None):
    """Exact top-k (score desc, id asc) of each row of `scores`, with per-row exclusions."""
    L = require_cuda()
    _dev(scores, torch.float32, "scores")
    n_q, n_items = scores.shape
    ids = torch.empty((n_q, topk), dtype=torch.int32, device=scores.device)
    sc = torch.empty((n_q, topk), dtype=torch.float32, device=scores.device)
    if excl_indptr is not None:
        _dev(excl_indptr, torch.int64, "excl_indptr"), _dev(excl_indices, torch.int32, "excl_indices")
    check(L.b200_topk_rows(ptr(scores), n_q, n_items, ptr(excl_indptr), ptr(excl_indices), int(topk), ptr(ids),
                           ptr(sc), current_stream()), "b200_topk_rows")
    return ids, sc


def rank_pack_items(V, item_base=None, n_items=None):
    """The item side of the fused rank packed once (b200_rank_pack_items): fp16 tile images of V[:n_items] with the item
    base folded in + the scaling scalars, as a uint8 CUDA tensor to pass to rank_topk(packed_items=...).  Valid as long as
    V / item_base do not change.  Returns None for shapes the tensor-core pass does not take."""
    L = require_cuda()
    _dev(V, torch.float32, "V")
    n_items = V.shape[0] if n_items is None else int(n_items)
    k = int(V.shape[1])
    nbytes = int(L.b200_rank_items_bytes(n_items, k))
    if nbytes <= 0:
        return None
    packed = torch.empty(nbytes, dtype=torch.uint8, device=V.device)
    check(L.b200_rank_pack_items(ptr(V), n_items, k, ptr(item_base), ptr(packed), nbytes, current_stream()),
          "b200_rank_pack_items")
    return packed


def rank_topk(U, V, topk, user_idx=None, item_base=None, user_off=None, excl_indptr=None, excl_indices=None,
              n_items=None, workspace=None, packed_items=None):
    """Fused score + exclusion + top-k on device tensors (b200_rank_topk).  Returns (ids int32 [n_q, topk],
    scores f32 [n_q, topk]) CUDA tensors ordered by (score desc, item id asc); ids are -1 padded.
    packed_items: result of rank_pack_items(V, item_base, n_items) for the SAME V / item_base / n_items (skips the two
    passes over V that every call otherwise makes)."""
    L = require_cuda()
    _dev(U, torch.float32, "U"), _dev(V, torch.float32, "V")
    n_items = V.shape[0] if n_items is None else int(n_items)
    n_q = U.shape[0] if user_idx is None else user_idx.numel()
    k = int(V.shape[1])
    ids = torch.empty((n_q, topk), dtype=torch.int32, device=U.device)
    sc = torch.empty((n_q, topk), dtype=torch.float32, device=U.device)
    nbytes = int(L.b200_rank_topk_workspace_bytes(n_q, n_items, k, int(topk)))
    if workspace is None or workspace.numel() < nbytes:
        workspace = torch.empty(max(nbytes, 16), dtype=torch.uint8, device=U.device)
    if packed_items is not None:
        _dev(packed_items, torch.uint8, "packed_items")
        if packed_items.numel() != int(L.b200_rank_items_bytes(n_items, k)):
            raise B200Error("packed_items was built for another item count / factor width")
    check(L.b200_rank_topk_packed(ptr(U), ptr(user_idx), n_q, ptr(V), n_items, k, ptr(item_base), ptr(user_off),
                                  ptr(excl_indptr), ptr(excl_indices), int(topk), ptr(ids), ptr(sc), ptr(packed_items),
                                  ptr(workspace), workspace.numel(), current_stream()), "b200_rank_topk")
    return ids, sc


def rank_topk_host(U, V, topk, user_idx, item_base=None, excl_indptr=None, excl_indices=None, out_ids=None,
                   out_scores=None, workspace=None, packed_items=None):
    """Host-buffer entry of the rank path (what the plug-ins' rank_batch calls): the factor matrices are
    device resident (model state), the REQUEST -- user indices and the per-user sorted exclusion lists in CSR
    form, numpy / pinned -- is copied in, ids + scores are copied back into numpy arrays."""
    uidx = to_device(np.asarray(user_idx, dtype=np.int64), torch.int64)
    ep = ei = None
    if excl_indptr is not None:
        ep = to_device(np.asarray(excl_indptr, dtype=np.int64), torch.int64)
        ei = to_device(np.asarray(excl_indices, dtype=np.int32), torch.int32) if len(excl_indices) else \
            torch.zeros(1, dtype=torch.int32, device="cuda")
    ids, sc = rank_topk(U, V, topk, user_idx=uidx, item_base=item_base, excl_indptr=ep, excl_indices=ei,
                        workspace=workspace, packed_items=packed_items)
    n_q = len(user_idx)
    out_ids = np.empty((n_q, topk), dtype=np.int32) if out_ids is None else out_ids
    out_scores = np.empty((n_q, topk), dtype=np.float32) if out_scores is None else out_scores
    torch.from_numpy(out_ids).copy_(ids)
    torch.from_numpy(out_scores).copy_(sc)
    return out_ids, out_scores


def topk_metrics(ids, pos_indptr, pos_indices, kinds, ks, user_idx=None, topk=None):
    """Per-user values of the @k ranking metrics for device ranked lists (b200_topk_metrics).
    ids int32 [n_q, >=topk] CUDA; pos_* the test-positives CSR (int64 / int32 CUDA); kinds / ks python lists.
    Returns a float64 CUDA tensor [n_metrics, n_q]."""
    L = require_cuda()
    _dev(ids, torch.int32, "ids"), _dev(pos_indptr, torch.int64, "pos_indptr"), _dev(pos_indices, torch.int32, "pos_indices")
    n_q, stride = ids.shape
    topk = stride if topk is None else int(topk)
    if user_idx is not None:
        _dev(user_idx, torch.int64, "user_idx")
    mk = torch.tensor(list(kinds), dtype=torch.int32).to(ids.device)
    kk = torch.tensor(list(ks), dtype=torch.int32).to(ids.device)
    out = torch.empty((len(kinds), n_q), dtype=torch.float64, device=ids.device)
    check(L.b200_topk_metrics(ptr(ids), n_q, topk, stride, ptr(user_idx), ptr(pos_indptr), ptr(pos_indices), ptr(mk),
                              ptr(kk), len(kinds), ptr(out), current_stream()), "b200_topk_metrics")
    return out


def rank_counts(scores, pos_indptr, pos_indices, user_idx=None, excl_indptr=None, excl_indices=None, less=None, pos_score=None):
    """The counts behind AUC / MAP / MRR for a batch of score rows (b200_rank_counts).  `scores` f32 [n_q, n_items] CUDA is
    MODIFIED (excluded entries become NaN).  pos_* = CSR of the test positives (int64 / int32 CUDA), row user_idx[q] for
    score row q.  Returns (less int64 [len(pos_indices)], pos_score f32 [len(pos_indices)], n_cand int64 [n_q],
    before_first int64 [n_q]); `less` / `pos_score` are filled only at the positions of the listed users' positives."""
    L = require_cuda()
    _dev(scores, torch.float32, "scores"), _dev(pos_indptr, torch.int64, "pos_indptr"), _dev(pos_indices, torch.int32, "pos_indices")
    n_q, n_items = scores.shape
    dev = scores.device
    if less is None:
        less = torch.zeros(max(pos_indices.numel(), 1), dtype=torch.int64, device=dev)
    if pos_score is None:
        pos_score = torch.zeros(max(pos_indices.numel(), 1), dtype=torch.float32, device=dev)
    n_cand = torch.zeros(max(n_q, 1), dtype=torch.int64, device=dev)
    before = torch.zeros(max(n_q, 1), dtype=torch.int64, device=dev)
    if user_idx is not None:
        _dev(user_idx, torch.int64, "user_idx")
    if excl_indptr is not None:
        _dev(excl_indptr, torch.int64, "excl_indptr"), _dev(excl_indices, torch.int32, "excl_indices")
    check(L.b200_rank_counts(ptr(scores), n_q, n_items, ptr(excl_indptr), ptr(excl_indices), ptr(user_idx), ptr(pos_indptr),
                             ptr(pos_indices), ptr(less), ptr(pos_score), ptr(n_cand), ptr(before), current_stream()),
          "b200_rank_counts")
    return less, pos_score, n_cand[:n_q], before[:n_q]


def delta_make(x, snapshot, delta):
    L = require_cuda()
    check(L.b200_delta_make(ptr(x), ptr(snapshot), ptr(delta), x.numel(), current_stream()), "b200_delta_make")


def delta_apply(x, snapshot, delta):
    L = require_cuda()
    check(L.b200_delta_apply(ptr(x), ptr(snapshot), ptr(delta), x.numel(), current_stream()), "b200_delta_apply")


def device_info():
    import ctypes
    L = require_cuda()
    a, b, c = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
    check(L.b200_device_info(ctypes.byref(a), ctypes.byref(b), ctypes.byref(c)), "b200_device_info")
    return dict(sm_count=a.value, cc=(b.value, c.value))
