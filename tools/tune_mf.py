"""MF Hogwild kernel throughput on a C4-shaped synthetic rating list (dev tool, GPU box)."""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cornac_b200 import engine  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=100_000_000)
    ap.add_argument("--users", type=int, default=1_000_000)
    ap.add_argument("--items", type=int, default=100_000)
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    import bench
    g = torch.Generator(device=dev)
    g.manual_seed(1)
    indptr, indices = bench.synth_interactions(args.users, args.items, args.n, 1234, dev)
    counts = (indptr[1:] - indptr[:-1]).to(torch.int64)
    rid_sorted = torch.repeat_interleave(torch.arange(args.users, device=dev, dtype=torch.int32), counts)
    perm = torch.randperm(args.n, generator=g, device=dev)
    val_all = torch.randint(1, 6, (args.n,), generator=g, device=dev).float()
    loss = torch.zeros(1, device=dev)
    for order, k in (("shuffled", 64), ("by-user", 64), ("shuffled", 128)):
        if order == "shuffled":
            rid, cid, val = rid_sorted[perm].contiguous(), indices[perm].contiguous(), val_all
        else:
            rid, cid, val = rid_sorted, indices, val_all
        U = torch.randn((args.users, k), generator=g, device=dev) * 0.01
        V = torch.randn((args.items, k), generator=g, device=dev) * 0.01
        Bu, Bi = torch.zeros(args.users, device=dev), torch.zeros(args.items, device=dev)
        for at in (1,):
            for _ in range(2):
                engine.mf_epoch(rid, cid, val, U, V, Bu, Bi, 0.01, 0.02, 3.0, True, loss, atomic=bool(at))
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(3):
                engine.mf_epoch(rid, cid, val, U, V, Bu, Bi, 0.01, 0.02, 3.0, True, loss, atomic=bool(at))
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 3
            gbs = args.n * (16 * k + 28) / ms / 1e6
            print("MF %s k=%d atomic=%d: %8.2f ms  %6.3f G ratings/s  %7.1f GB/s algorithmic" % (order, k, at, ms, args.n / ms / 1e6, gbs), flush=True)


if __name__ == "__main__":
    main()
