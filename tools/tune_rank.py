"""Throughput of b200_rank_topk (tensor-core fused rank) on C2- and C5-shaped problems (dev tool, GPU box)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cornac_b200._lib import check, current_stream, load, ptr  # noqa: E402


def run(n_users, n_items, k, n_q, topk, n_excl, reps=3):
    dev = torch.device("cuda", 0)
    g = torch.Generator(device=dev)
    g.manual_seed(1)
    U = torch.randn((n_users, k), generator=g, device=dev) * 0.1
    V = torch.randn((n_items, k), generator=g, device=dev) * 0.1
    B = torch.randn(n_items, generator=g, device=dev) * 0.1
    uidx = torch.randperm(n_users, generator=g, device=dev)[:n_q].contiguous()
    ex_ptr = ex_idx = None
    if n_excl:
        ex = torch.randint(0, n_items, (n_q, n_excl), generator=g, device=dev, dtype=torch.int32)
        ex_idx = torch.sort(ex, dim=1)[0].contiguous().view(-1)
        ex_ptr = (torch.arange(n_q + 1, device=dev, dtype=torch.int64) * n_excl).contiguous()
    L = load()
    if os.environ.get("B200_ALT_LIB"):          # A/B against another build of the library (only the two rank entry points are bound)
        import ctypes
        L = ctypes.CDLL(os.environ["B200_ALT_LIB"])
        vp, i64, ci = ctypes.c_void_p, ctypes.c_int64, ctypes.c_int
        L.b200_rank_topk_workspace_bytes.restype = i64
        L.b200_rank_topk_workspace_bytes.argtypes = [i64, i64, ci, ci]
        L.b200_rank_topk.restype = ci
        L.b200_rank_topk.argtypes = [vp, vp, i64, vp, i64, ci, vp, vp, vp, vp, ci, vp, vp, vp, i64, vp]
        L.b200_last_error.restype = ctypes.c_char_p
    ids = torch.empty((n_q, topk), dtype=torch.int32, device=dev)
    sc = torch.empty((n_q, topk), dtype=torch.float32, device=dev)
    nb = int(L.b200_rank_topk_workspace_bytes(n_q, n_items, k, topk))
    ws = torch.empty(max(nb, 16), dtype=torch.uint8, device=dev)

    def go():
        rc = L.b200_rank_topk(ptr(U), ptr(uidx), n_q, ptr(V), n_items, k, ptr(B), None, ptr(ex_ptr), ptr(ex_idx), topk,
                              ptr(ids), ptr(sc), ptr(ws), nb, current_stream())
        assert rc == 0, L.b200_last_error()
    go()
    torch.cuda.synchronize()
    if os.environ.get("TUNE_DEBUG_AFTER_WARMUP"):      # e.g. 32: the timed calls start from the warm-up call's final thresholds
        os.environ["B200_RANK_DEBUG"] = os.environ["TUNE_DEBUG_AFTER_WARMUP"]
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        go()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    print("rank n_items=%d k=%d n_q=%d top%d excl=%d: %9.3f ms  %10.0f users/s  %7.1f TFLOP/s  ws=%.0f MB"
          % (n_items, k, n_q, topk, n_excl, ms, n_q / ms * 1e3, 2.0 * k * n_items * n_q / ms / 1e9, nb / 1e6), flush=True)


if __name__ == "__main__":
    tc = os.environ.get("B200_RANK_TC", "1")
    print("B200_RANK_TC =", tc, " B200_RANK_DEBUG =", os.environ.get("B200_RANK_DEBUG"))
    if os.environ.get("TUNE_ONLY") == "c5":
        run(1_000_000, 1_000_000, 128, 18944, int(os.environ.get("TUNE_TOPK", "100")), int(os.environ.get("TUNE_EXCL", "100")))
        sys.exit(0)
    run(1_000_000, 100_000, 64, 4096, 100, 100)
    run(1_000_000, 100_000, 64, 18944, 100, 100)
    run(1_000_000, 100_000, 64, 75776, 100, 100)
    run(1_000_000, 100_000, 64, 75776, 100, 0)
    run(1_000_000, 100_000, 64, 75776, 10, 100)
    if tc != "0":
        run(1_000_000, 1_000_000, 128, 18944, 100, 100)
        run(1_000_000, 1_000_000, 128, 75776, 100, 100)
