"""N-GPU functional check (run under torchrun on the GPU box):
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/mgpu_check.py
Trains one BPR model with parallel.bpr_fit_sharded, verifies that the item replicas agree bit-for-bit
across ranks, that every rank trained only its own users, and that the model learnt the planted structure."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cornac_b200 import parallel  # noqa: E402


def mf_check(rank, world):
    """parallel.mf_fit_sharded: replicas of V / Bi identical on every rank, only the rank's own users trained, loss falls."""
    rng = np.random.RandomState(9)                      # same data on every rank
    n_users, n_items, k, n = 20000, 3000, 32, 600000
    P, Q = rng.normal(0, 0.5, (n_users, 4)), rng.normal(0, 0.5, (n_items, 4))
    rid = rng.randint(n_users, size=n).astype(np.int64)
    cid = rng.randint(n_items, size=n).astype(np.int64)
    val = (3.0 + np.einsum("nk,nk->n", P[rid], Q[cid]) + rng.normal(0, 0.1, n)).astype(np.float32)
    U = rng.normal(0, 0.01, (n_users, k)).astype(np.float32)
    V = rng.normal(0, 0.01, (n_items, k)).astype(np.float32)
    Bu, Bi = np.zeros(n_users, np.float32), np.zeros(n_items, np.float32)
    U0 = U.copy()
    bounds, losses = parallel.mf_fit_sharded(rid, cid, val, U, V, Bu, Bi, 0.02, 0.01, float(val.mean()), True, max_iter=10)
    lo, hi = int(bounds[rank]), int(bounds[rank + 1])
    Vd = torch.from_numpy(np.concatenate([V.ravel(), Bi])).cuda()
    gathered = [torch.empty_like(Vd) for _ in range(world)]
    dist.all_gather(gathered, Vd)
    same = all(torch.equal(gathered[0], g) for g in gathered)
    untouched = np.array_equal(np.delete(U, np.s_[lo:hi], axis=0), np.delete(U0, np.s_[lo:hi], axis=0))
    # from a 0.01-scale start this problem leaves the saddle slowly: the sequential reference loop (oracle, one process) goes
    # 79.4 K -> 77.1 K in 10 epochs; the sharded Hogwild run must track that (measured: 79402 -> 77089 on 2 GPUs)
    falling = all(b < a for a, b in zip(losses, losses[1:]))
    ok = same and untouched and falling and 0.96 * losses[0] < losses[-1] < 0.98 * losses[0]
    print("rank %d/%d MF users [%d,%d): replicas_equal=%s untouched=%s loss %.1f -> %.1f -> %s"
          % (rank, world, lo, hi, same, untouched, losses[0], losses[-1], "OK" if ok else "FAIL"), flush=True)
    return ok


def exchange_check(rank, world):
    """PeerItemExchange (one fused NVLink kernel) against ItemReplicaSync (delta kernels + NCCL all-reduce) on the same
    local changes: every replica ends at start + sum of all ranks' changes; the p2p result is bit-equal on all ranks."""
    ok = True
    for n, reps in ((1_000_003, 3), (128 * 50_000, 2), (7, 2)):
        g = torch.Generator(device="cuda").manual_seed(100 + n)          # same start on every rank
        start = torch.randn(n, generator=g, device="cuda")
        xa, xb = start.clone(), start.clone()
        pa = parallel.PeerItemExchange([xa])
        pb = parallel.ItemReplicaSync([xb])
        want = start.double().clone()
        for it in range(reps):
            gl = torch.Generator(device="cuda").manual_seed(7 * n + 31 * it)
            deltas = [torch.randn(n, generator=gl, device="cuda") * 0.01 * (r + 1) for r in range(world)]     # known on every rank
            xa += deltas[rank]
            xb += deltas[rank]
            pa.exchange()
            pb.exchange()
            tot, cnt = torch.zeros_like(want), torch.zeros_like(want)
            for d in deltas:
                tot += d.double()
                cnt += (d != 0).double()
            want += tot / cnt.clamp(min=1)          # the default rule: mean over the ranks that changed the element
        torch.cuda.synchronize()
        err_a = float((xa.double() - want).abs().max())
        err_b = float((xb.double() - want).abs().max())
        gathered = [torch.empty_like(xa) for _ in range(world)]
        dist.all_gather(gathered, xa)
        same = all(torch.equal(gathered[0], t) for t in gathered)
        good = err_a < 1e-4 and err_b < 1e-4 and same and not pa.failed()
        # timing of one exchange of this size (both ways), after the correctness rounds
        ts = []
        for sync in (pa, pb):
            torch.cuda.synchronize(); dist.barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(5):
                sync.exchange()
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1) / 5)
        pa.close()
        print("rank %d/%d exchange n=%d: p2p err %.2e, nccl err %.2e, replicas bit-equal=%s, p2p %.3f ms vs nccl path %.3f ms -> %s"
              % (rank, world, n, err_a, err_b, same, ts[0], ts[1], "OK" if good else "FAIL"), flush=True)
        ok = ok and good
    return ok


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ok_x = exchange_check(rank, world)
    rng = np.random.RandomState(4)                      # same data on every rank
    n_users, n_items, k = 20000, 3000, 32
    cu, ci = rng.randint(8, size=n_users), rng.randint(8, size=n_items)
    rows = []
    for u in range(n_users):
        own = np.flatnonzero(ci == cu[u])
        rows.append(np.sort(rng.choice(own, size=20, replace=False)))
    indptr = np.concatenate([[0], np.cumsum([len(r) for r in rows])]).astype(np.int32)
    indices = np.concatenate(rows).astype(np.int32)
    U = ((rng.uniform(0, 1, (n_users, k)).astype(np.float32) - 0.5) / k)
    V = ((rng.uniform(0, 1, (n_items, k)).astype(np.float32) - 0.5) / k)
    B = np.zeros(n_items, np.float32)
    U0 = U.copy()
    bounds, hist = parallel.bpr_fit_sharded(indptr, indices, n_items, U, V, B, 0.05, 0.001, True, max_iter=15, key=7)
    lo, hi = int(bounds[rank]), int(bounds[rank + 1])
    # replicas identical
    Vd = torch.from_numpy(V).cuda()
    gathered = [torch.empty_like(Vd) for _ in range(world)]
    dist.all_gather(gathered, Vd)
    same = all(torch.equal(gathered[0], g) for g in gathered)
    # only own rows trained
    untouched = np.array_equal(np.delete(U, np.s_[lo:hi], axis=0), np.delete(U0, np.s_[lo:hi], axis=0))
    moved = np.abs(U[lo:hi] - U0[lo:hi]).max() > 1e-3
    # learnt: own users score their cluster's items above the others
    s = U[lo:hi] @ V.T + B
    mine = (ci[None, :] == cu[lo:hi, None])
    auc_like = float((s[mine].mean() - s[~mine].mean()) / (s.std() + 1e-9))
    c, sk = hist[-1]
    acc = c / max(1, (indptr[hi] - indptr[lo]) - sk)
    ok = same and untouched and moved and acc > 0.9 and auc_like > 1.0
    print("rank %d/%d users [%d,%d): replicas_equal=%s untouched=%s moved=%s acc=%.3f sep=%.2f -> %s"
          % (rank, world, lo, hi, same, untouched, moved, acc, auc_like, "OK" if ok else "FAIL"), flush=True)
    ok = mf_check(rank, world) and ok and ok_x
    flag = torch.tensor([0 if ok else 1], device="cuda")
    dist.all_reduce(flag)
    dist.destroy_process_group()
    sys.exit(int(flag.item() != 0))


if __name__ == "__main__":
    main()
