"""N>1 host logic on CPU: sharding + the item-replica exchange over gloo (world_size 2)."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import synth_csr
from cornac_b200.parallel import ItemReplicaSync, shard_csr, shard_users_by_nnz


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
    sync = ItemReplicaSync([V, B], ops=_CpuOps)
    total_dv = torch.zeros_like(V)
    for epoch in range(3):
        gl = torch.Generator().manual_seed(100 * epoch + rank)
        dv, db = torch.randn(50, 8, generator=gl) * 0.01, torch.randn(50, generator=gl) * 0.01
        V += dv
        B += db
        sync.exchange()
        for r in range(world):
            gr = torch.Generator().manual_seed(100 * epoch + r)
            total_dv += torch.randn(50, 8, generator=gr) * 0.01
            torch.randn(50, generator=gr)
    ok = torch.allclose(V, V0 + total_dv, atol=1e-6)
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
