"""Sweep launch shapes of the BPR throughput kernel on the bench workload (dev tool, GPU box).

    python tools/tune_bpr.py [--k 64] [--scale 1.0]
Prints one line per configuration: ms per epoch, G updates/s."""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from cornac_b200 import engine  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--k", type=int, default=64)
    ap.add_argument("--scale", type=float, default=1.0)
    ap.add_argument("--uniform", type=int, default=0)
    ap.add_argument("--items", type=int, default=0)
    ap.add_argument("--users", type=int, default=0)
    ap.add_argument("--nnz", type=int, default=0)
    ap.add_argument("--c3", type=int, default=0, help="1: block 0 of bench.py's configs[2] model (1.25M x 1M x 125M, k=128)")
    ap.add_argument("--epochs", type=int, default=3)
    ap.add_argument("--blocked", type=int, default=0, help="1: cache-blocked sample order (B200_BPR_BLOCKED)")
    args = ap.parse_args()
    W = dict(bench.WORKLOADS["c2"])
    if args.c3:
        W = dict(bench.WORKLOADS["c3"])
        args.k = W["k"]
    W["k"] = args.k
    W["n_users"] = int(W["n_users"] * args.scale)
    W["nnz"] = int(W["nnz"] * args.scale)
    if args.items:
        W["n_items"] = args.items
    if args.users:
        W["n_users"] = args.users
    if args.nnz:
        W["nnz"] = args.nnz
    dev = torch.device("cuda", 0)
    if args.c3:
        indptr, indices = bench.synth_shard(W, [0], dev)
        W["n_users"] //= bench.N_BLOCKS
    else:
        indptr, indices = bench.synth_interactions(W["n_users"], W["n_items"], W["nnz"], 1234, dev)
    if args.uniform:
        indices = torch.randint(0, W["n_items"], indices.shape, device=dev, dtype=torch.int32)   # law check only
    data = engine.BprData(indptr, indices).prepare()
    g = torch.Generator(device=dev).manual_seed(99)
    U = (torch.rand((W["n_users"], args.k), generator=g, device=dev) - 0.5) / args.k
    V = (torch.rand((W["n_items"], args.k), generator=g, device=dev) - 0.5) / args.k
    B = torch.zeros(W["n_items"], device=dev)
    stats = torch.zeros(2, dtype=torch.int64, device=dev)
    configs = [("chunk minb4 thr=%d blk=%d" % (thr, blk), "0,%d,%d" % (thr, blk), 1) for thr in (256, 128) for blk in (0, 3, 2)]
    configs += [("chunk minb3 thr=256", "16,256,0", 1), ("old S=1 thr=256 blk=4", "1,256,4", 1), ("old S=1 thr=128", "1,128,0", 1),
                ("chunk minb4 plain-stores", "0,256,0", 0)]
    if os.environ.get("B200_TUNE_EXPERIMENT"):
        configs = [("default", os.environ.get("B200_BPR_TUNE", "0,256,0"), 1)]
    if os.environ.get("B200_TUNE_DEPTH"):
        configs = [("depth1 (default)", "0,256,0", 1), ("depth2 minb3", "32,256,0", 1), ("depth1 blk=2", "0,256,2", 1),
                   ("depth2 blk=2", "32,256,2", 1)]
    for name, tune, at in configs:
        os.environ["B200_BPR_TUNE"] = tune
        for e in range(2):
            engine.bpr_epoch(data, W["n_items"], U, V, B, W["lr"], W["reg"], True, 1, e, stats, atomic=bool(at), blocked=bool(args.blocked))
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for e in range(args.epochs):
            engine.bpr_epoch(data, W["n_items"], U, V, B, W["lr"], W["reg"], True, 1, 10 + e, stats, atomic=bool(at), blocked=bool(args.blocked))
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / args.epochs
        c, sk = stats.cpu().tolist()
        print("k=%d %-32s %8.2f ms  %6.3f G samples/s  blocked=%d plan=%s acc=%.4f" % (args.k, name, ms, data.nnz / ms / 1e6, args.blocked,
              engine.bpr_block_plan(W["n_users"], W["n_items"], args.k), c / max(1, (2 + args.epochs) * data.nnz - sk)), flush=True)
        stats.zero_()


if __name__ == "__main__":
    main()
