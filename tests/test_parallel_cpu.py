"""N>1 host logic on CPU: sharding + the item-replica exchange over gloo (world_size 2)."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import synth_csr
from cornac_b200.parallel import (ItemReplicaSync, mf_fit_sharded, shard_csr, shard_ratings, shard_ratings_by_user,
                                  shard_users_by_nnz)


def test_shard_users_balanced_and_complete():
    indptr, indices = synth_csr(5000, 300, 60000, seed=1)
    for world in (1, 2, 3, 8):
        b = shard_users_by_nnz(indptr, world)
        assert b[0] == 0 and b[-1] == 5000 and np.all(np.diff(b) >= 0)
        sizes = [int(indptr[b[r + 1]] - indptr[b[r]]) for r in range(world)]
        assert sum(sizes) == len(indices)
        assert max(sizes) - min(sizes) <= 2 * int(np.diff(indptr).max())
        got = np.concatenate([shard_csr(indptr, indices, b, r)[1] for r in range(world)])
        assert np.array_equal(got, indices)
        ip, _ = shard_csr(indptr, indices, b, world - 1)
        assert ip[0] == 0 and ip[-1] == sizes[-1]
    # degenerate: more ranks than users with interactions
    b = shard_users_by_nnz(np.array([0, 0, 5, 5]), 4)
    assert b[0] == 0 and b[-1] == 3 and np.all(np.diff(b) >= 0)


class _CpuOps:          # test double for the two element-wise CUDA kernels
    @staticmethod
    def make(x, s, d):
        torch.sub(x, s, out=d)

    @staticmethod
    def apply(x, s, d):
        x.copy_(s + d)
        s.copy_(x)


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    g = torch.Generator().manual_seed(0)
    V = torch.randn(50, 8, generator=g)
    B = torch.randn(50, generator=g)
    V0, B0 = V.clone(), B.clone()
    # default rule: every element moves by the MEAN of the changes of the ranks that changed it (rows r, r + 3, ... are left
    # alone by rank r here, so the count differs from row to row); reduce="sum" is the plain sum
    sync = ItemReplicaSync([V, B], ops=_CpuOps)
    Vs = V.clone()
    sync_sum = ItemReplicaSync([Vs], ops=_CpuOps, reduce="sum")
    want, want_sum = V0.clone(), V0.clone()

    def change(epoch, r):
        gr = torch.Generator().manual_seed(100 * epoch + r)
        dv, db = torch.randn(50, 8, generator=gr) * 0.01, torch.randn(50, generator=gr) * 0.01
        dv[r::3] = 0
        return dv, db

    for epoch in range(3):
        dv, db = change(epoch, rank)
        V += dv
        Vs += dv
        B += db
        sync.exchange()
        sync_sum.exchange()
        tot, cnt = torch.zeros_like(V), torch.zeros_like(V)
        for r in range(world):
            d, _ = change(epoch, r)
            tot += d
            cnt += (d != 0).float()
        want += tot / cnt.clamp(min=1)
        want_sum += tot
    ok = torch.allclose(V, want, atol=1e-6) and torch.allclose(Vs, want_sum, atol=1e-6) and not torch.allclose(want, want_sum, atol=1e-4)
    gathered = [torch.empty_like(V) for _ in range(world)]
    dist.all_gather(gathered, V)
    same = all(torch.equal(gathered[0], t) for t in gathered)
    out[rank] = bool(ok and same and not torch.equal(B, B0))
    dist.destroy_process_group()


def test_item_replica_exchange_gloo_world2():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    assert out[0] and out[1]


# ---------------------------------------------------------------------------------------------
# MF sharded by user: the host logic of parallel.mf_fit_sharded over gloo, world size 2, with the
# oracle's sequential epoch as the stand-in for b200_mf_epoch and torch-CPU stand-ins for the delta kernels.
class _OracleMf:
    ops = _CpuOps

    def ids(self, a):
        return np.ascontiguousarray(a, dtype=np.int64)

    def f32(self, a):
        return torch.from_numpy(np.array(a, dtype=np.float32, copy=True))

    def zeros1(self):
        return torch.zeros(1)

    def epoch(self, rid, cid, val, U, V, Bu, Bi, lr, reg, mu, use_bias, loss, atomic):
        from oracle import oracle as O
        half = O.mf_epoch(rid, cid, val.numpy(), U.numpy(), V.numpy(), Bu.numpy(), Bi.numpy(), lr, reg, mu, use_bias)
        loss[0] = 2.0 * half                           # the device kernel reports sum(err^2)

    def to_host(self, host, dev):
        host[...] = dev.numpy()


def _mf_problem():
    rng = np.random.RandomState(3)
    n_users, n_items, n, k = 120, 40, 3000, 8
    rid = rng.randint(n_users, size=n).astype(np.int64)
    cid = rng.randint(n_items, size=n).astype(np.int64)
    val = rng.randint(1, 6, size=n).astype(np.float32)
    U = rng.normal(0, 0.01, (n_users, k)).astype(np.float32)
    V = rng.normal(0, 0.01, (n_items, k)).astype(np.float32)
    return rid, cid, val, U, V, np.zeros(n_users, np.float32), np.zeros(n_items, np.float32)


def test_shard_ratings_by_user_partitions_the_rating_list():
    rid, cid, val, U, *_ = _mf_problem()
    for world in (1, 2, 3):
        b = shard_ratings_by_user(rid, U.shape[0], world)
        parts = [shard_ratings(rid, cid, val, b, r) for r in range(world)]
        assert sum(len(p[2]) for p in parts) == len(val)
        sizes = [len(p[2]) for p in parts]
        assert max(sizes) - min(sizes) <= 2 * np.bincount(rid).max()
        for r, (rr, cc, vv) in enumerate(parts):
            keep = (rid >= b[r]) & (rid < b[r + 1])
            assert np.array_equal(rr + b[r], rid[keep]) and np.array_equal(cc, cid[keep]) and np.array_equal(vv, val[keep])


def _mf_worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import oracle as O
    rid, cid, val, U, V, Bu, Bi = _mf_problem()
    U0, V0, Bu0, Bi0 = U.copy(), V.copy(), Bu.copy(), Bi.copy()
    bounds, losses = mf_fit_sharded(rid, cid, val, U, V, Bu, Bi, 0.01, 0.02, 3.0, True, max_iter=1, _device=_OracleMf())
    lo, hi = int(bounds[rank]), int(bounds[rank + 1])
    # expectation for one epoch: every shard trained from the SAME start; V, Bi = start + the shards' changes averaged over the
    # shards that changed the element (parallel.py: why not the sum)
    dV, dBi, tot = np.zeros_like(V0), np.zeros_like(Bi0), 0.0
    cV, cBi = np.zeros_like(V0), np.zeros_like(Bi0)
    for r in range(world):
        rr, cc, vv = shard_ratings(rid, cid, val, bounds, r)
        a, b = int(bounds[r]), int(bounds[r + 1])
        Ur, Vr, Bur, Bir = U0[a:b].copy(), V0.copy(), Bu0[a:b].copy(), Bi0.copy()
        tot += O.mf_epoch(rr, cc, vv, Ur, Vr, Bur, Bir, 0.01, 0.02, 3.0, True)
        dV += Vr - V0
        dBi += Bir - Bi0
        cV += (Vr - V0 != 0)
        cBi += (Bir - Bi0 != 0)
        if r == rank:
            mine = (Ur, Bur)
    ok = (np.allclose(V, V0 + dV / np.maximum(cV, 1), atol=1e-6) and np.allclose(Bi, Bi0 + dBi / np.maximum(cBi, 1), atol=1e-6)
          and np.array_equal(U[lo:hi], mine[0]) and np.array_equal(Bu[lo:hi], mine[1])
          and np.array_equal(np.delete(U, np.s_[lo:hi], axis=0), np.delete(U0, np.s_[lo:hi], axis=0))
          and abs(losses[0] - tot) <= 1e-3 * tot)
    # several epochs + early stop: same decision and same replicas everywhere
    _, l2 = mf_fit_sharded(rid, cid, val, U, V, Bu, Bi, 0.01, 0.02, 3.0, True, max_iter=4, early_stop=True, _device=_OracleMf())
    g = [torch.empty(V.shape) for _ in range(world)]
    dist.all_gather(g, torch.from_numpy(V.copy()))
    n_ep = torch.tensor([len(l2)])
    dist.all_reduce(n_ep, op=dist.ReduceOp.MAX)
    out[rank] = bool(ok and all(torch.equal(g[0], t) for t in g) and int(n_ep.item()) == len(l2) and l2[-1] < losses[0])
    dist.destroy_process_group()


def test_mf_fit_sharded_gloo_world2():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_mf_worker, args=(2, port, out), nprocs=2, join=True)
    assert out[0] and out[1]
