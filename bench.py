#!/usr/bin/env python
"""bench.py -- BPR triplet-updates/s (and ranked users/s) on B200 vs the reference CPU path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload (BASELINE.json configs[1]): synthetic 1M users x 100K items x 100M interactions,
BPR k=64, one B200.  A "step" is one BPR epoch = nnz sampled triplets (what one call of the
reference's BPR._fit_sgd does, cornac/models/bpr/recom_bpr.pyx:208-269).  For N > 1 every
rank holds its own 1M-user / 100M-interaction shard (weak scaling), the 100K-item matrix is
replicated and the item deltas are all-reduced once per epoch inside the timed region.

value  = non-skipped triplet updates per second, inputs resident in HBM (CUDA events, max
         over ranks);
e2e    = the same through the host-buffer entry the plug-in's fit() uses
         (engine.bpr_train_host): pinned-host CSR + factors -> H2D -> one epoch -> D2H;
roofline / cpu_baseline / rank (ranked users/s) are reported alongside.
--impl reference times the reference's own Cython/OpenMP kernel (oracle/_ref) on the host
cores, on a bounded sample of the same workload.
"""
import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOAD = dict(n_users=1_000_000, n_items=100_000, nnz=100_000_000, k=64, lr=0.05, reg=0.01, use_bias=True)
# one GPU's share of BASELINE.json configs[2] (10M users x 1M items x 1B interactions, k=128, over 8 GPUs); `--workload c3shard`
WORKLOAD_C3_SHARD = dict(n_users=1_250_000, n_items=1_000_000, nnz=125_000_000, k=128, lr=0.05, reg=0.01, use_bias=True)
RANK_WORKLOAD = dict(n_q=75776, topk=100)          # secondary metric: ranked users/s on the same model
CPU_SAMPLE = dict(n_users=100_000, nnz_target=10_000_000)   # bounded sample for the CPU legs (same k, same items)


def log(*a):
    print(*a, file=sys.stderr, flush=True)


# --------------------------------------------------------------------------------------
# synthetic data (torch on the GPU is used as a fast array library here; not product code)
def synth_interactions(n_users, n_items, nnz, seed, device):
    """Unique (u, i) pairs: user activity ~ log-normal (mean degree nnz/n_users), item
    popularity ~ Zipf(1.0); returns CSR (indptr int32, indices int32) on `device`."""
    import torch
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    w_u = torch.exp(torch.randn(n_users, generator=g, device=device, dtype=torch.float64))
    cdf_u = torch.cumsum(w_u / w_u.sum(), 0)
    w_i = 1.0 / torch.arange(1, n_items + 1, device=device, dtype=torch.float64)
    perm = torch.randperm(n_items, generator=g, device=device)          # popular ids spread over the id range
    cdf_i = torch.cumsum(w_i / w_i.sum(), 0)
    keys = torch.empty(0, dtype=torch.int64, device=device)
    need = nnz
    while need > 0:
        m = int(need * 1.25) + 1024
        u = torch.searchsorted(cdf_u, torch.rand(m, generator=g, device=device, dtype=torch.float64)).clamp_(max=n_users - 1)
        i = perm[torch.searchsorted(cdf_i, torch.rand(m, generator=g, device=device, dtype=torch.float64)).clamp_(max=n_items - 1)]
        keys = torch.unique(torch.cat([keys, u * n_items + i]))
        del u, i
        need = nnz - keys.numel()
    if keys.numel() > nnz:
        drop = torch.randperm(keys.numel(), generator=g, device=device)[: keys.numel() - nnz]
        mask = torch.ones(keys.numel(), dtype=torch.bool, device=device)
        mask[drop] = False
        keys = keys[mask]
    u = keys // n_items
    indices = (keys % n_items).to(torch.int32)
    counts = torch.bincount(u, minlength=n_users)
    indptr = torch.zeros(n_users + 1, dtype=torch.int64, device=device)
    indptr[1:] = torch.cumsum(counts, 0)
    return indptr.to(torch.int32), indices.contiguous()


def init_factors(n_users, n_items, k, seed, device):
    """(U[0,1) - 0.5) / k like BPR._init (recom_bpr.pyx:145-152)."""
    import torch
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    U = (torch.rand((n_users, k), generator=g, device=device) - 0.5) / k
    V = (torch.rand((n_items, k), generator=g, device=device) - 0.5) / k
    B = torch.zeros(n_items, device=device)
    return U, V, B


# --------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 20 ms; samples are stamped on arrival and
    only those inside the timed window are reported (fallback: all samples of the run)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx, self.proc, self.lines = gpu_index, None, []
        self.t0 = self.t1 = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append((time.time(), line.strip()))

    def window_begin(self):
        self.t0 = time.time()

    def window_end(self):
        self.t1 = time.time()

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.1)
        self.proc.terminate()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]

        def digest(rows):
            sm, smax, pw, reasons = [], [], [], set()
            for _, l in rows:
                f = [x.strip() for x in l.split(",")]
                if len(f) < 9:
                    continue
                try:
                    sm.append(float(f[1])); smax.append(float(f[2])); pw.append(float(f[3]))
                except ValueError:
                    continue
                for nm, v in zip(names, f[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(nm)
            return sm, smax, pw, reasons
        inside = [r for r in self.lines if self.t0 is not None and self.t0 - 0.01 <= r[0] <= (self.t1 or 1e30) + 0.03]
        window = "timed region"
        if len(inside) < 2:
            inside, window = self.lines, "whole run (timed region shorter than the sampling period)"
        sm, smax, pw, reasons = digest(inside)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "window": window, "reasons": sorted(reasons)}


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def tensor_peak():
    """dense 16-bit tensor TFLOP/s: the sustained figure of MEASURED_PEAKS.json (the kernel runs for many ms)."""
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        for key in ("bf16_tflops_sustained", "bf16_tflops"):
            if key in d:
                return float(d[key])
    return 2250.0


def algorithmic_bytes(k, updates, skipped, mean_deg):
    """3 rows read + 3 written (24k B) + 4 biases R/W (16 B) + indices/coo/indptr
    (16 B) + 4*ceil(log2(deg+1)) B binary search per update; a skipped sample costs only the
    sampling + search bytes."""
    search = 4 * math.ceil(math.log2(mean_deg + 1))
    return updates * (24 * k + 16 + 16 + search) + skipped * (16 + search)


# --------------------------------------------------------------------------------------
def reference_fit_sgd_runner(indptr, indices, n_items, k, lr, reg, n_threads):
    """Returns (run_epoch() -> (correct, skipped), kind, cores).  Uses the UNMODIFIED compiled
    reference (oracle/_ref: cornac.models.bpr.recom_bpr.BPR._fit_sgd + RNGVector) when it is
    importable, else the oracle's OpenMP port."""
    ref = os.path.join(ROOT, "oracle", "_ref")
    n_users = len(indptr) - 1
    rng = np.random.RandomState(1)
    U = ((rng.uniform(0, 1, (n_users, k)).astype(np.float32) - 0.5) / k)
    V = ((rng.uniform(0, 1, (n_items, k)).astype(np.float32) - 0.5) / k)
    B = np.zeros(n_items, np.float32)
    user_ids = np.repeat(np.arange(n_users), np.diff(indptr)).astype(np.int32)
    try:
        if ref not in sys.path:
            sys.path.insert(0, ref)
        import multiprocessing
        from cornac.models.bpr.recom_bpr import BPR as RefBPR, RNGVector
        n_threads = n_threads or multiprocessing.cpu_count()
        m = RefBPR(k=k, learning_rate=lr, lambda_reg=reg, use_bias=True)
        rp = RNGVector(n_threads, len(user_ids) - 1, 11)
        rn = RNGVector(n_threads, n_items - 1, 12)
        neg = np.arange(n_items, dtype=np.int32)

        def run():
            return m._fit_sgd(rp, rn, n_threads, user_ids, indices, neg, indptr, U, V, B)
        return run, "reference", n_threads
    except Exception as e:   # reference install absent: the oracle port, all threads
        log("[bench] reference install not importable (%s); using the oracle OpenMP port" % (e,))
        from oracle import oracle as O
        n_threads = n_threads or O.n_threads()
        state = {"e": 0}

        def run():
            state["e"] += 1
            return O.bpr_epoch_omp(indptr, indices, n_items, U, V, B, lr, reg, True, n_threads, seed=state["e"])
        return run, "port", n_threads


def cpu_sample_csr(indptr_host, indices_host):
    """Bounded sample of the workload for the CPU legs: the first users of the same synthetic
    matrix (same k, same 100K items, ~1/10 of the interactions)."""
    n_u = CPU_SAMPLE["n_users"]
    n_u = min(n_u, len(indptr_host) - 1)
    end = int(indptr_host[n_u])
    return np.ascontiguousarray(indptr_host[: n_u + 1]), np.ascontiguousarray(indices_host[:end])


def time_cpu_epochs(run, nnz, min_seconds=8.0, max_epochs=6):
    run()                                             # warm-up epoch (page-in, thread pool)
    t_tot, upd, eps = 0.0, 0, 0
    while t_tot < min_seconds and eps < max_epochs:
        t0 = time.perf_counter()
        c, s = run()
        t_tot += time.perf_counter() - t0
        upd += nnz - s
        eps += 1
    return upd / t_tot, t_tot, eps


# --------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--scale", type=float, default=1.0, help="shrink the workload (debug only; result then INVALID)")
    ap.add_argument("--workload", default="c2", choices=["c2", "c3shard"],
                    help="c2 = BASELINE configs[1] per GPU (default, the quoted metric); c3shard = one GPU's share of configs[2]")
    ap.add_argument("--atomic", type=int, default=1, help="1: red.global.add scatter (default), 0: plain racy stores")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-rank", action="store_true")
    args = ap.parse_args()

    # keep stdout to the one JSON line: NCCL prints its version banner to stdout at NCCL_DEBUG=VERSION and above (WARN
    # included).  An explicit value that names no level silences it whatever the box's environment or nccl.conf say.
    os.environ["NCCL_DEBUG"] = os.environ.get("B200_NCCL_DEBUG", "NONE")
    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    W = dict(WORKLOAD if args.workload == "c2" else WORKLOAD_C3_SHARD)
    if args.scale != 1.0:
        W["n_users"] = max(1000, int(W["n_users"] * args.scale))
        W["nnz"] = max(10000, int(W["nnz"] * args.scale))
    k = W["k"]

    if args.impl == "reference":
        if rank != 0:
            return 0
        return run_reference_arm(args, W)

    assert torch.cuda.is_available(), "bench.py needs a GPU (no CPU fallback)"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    from cornac_b200 import engine
    from cornac_b200.parallel import ItemReplicaSync

    # ---- data: every rank builds its own user shard (weak scaling), items are shared
    t0 = time.time()
    indptr, indices = synth_interactions(W["n_users"], W["n_items"], W["nnz"], seed=1234 + rank, device=dev)
    U, V, B = init_factors(W["n_users"], W["n_items"], k, seed=99, device=dev)     # V/B identical on all ranks
    if world > 1:
        dist.broadcast(V, 0); dist.broadcast(B, 0)
    data = engine.BprData(indptr, indices)
    nnz = data.nnz
    mean_deg = nnz / W["n_users"]
    torch.cuda.synchronize()
    log("[bench] rank %d data ready in %.1fs: %d users x %d items x %d nnz" % (rank, time.time() - t0, W["n_users"], W["n_items"], nnz))
    sync = ItemReplicaSync([V, B]) if world > 1 else None
    stats = torch.zeros(2, dtype=torch.int64, device=dev)
    key = 0xB200

    def step(epoch):
        engine.bpr_epoch(data, W["n_items"], U, V, B, W["lr"], W["reg"], W["use_bias"], key + rank, epoch, stats,
                         atomic=bool(args.atomic))
        if sync is not None:
            sync.exchange()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for e in range(args.warmup):
        step(e)
    barrier()
    stats.zero_()
    clocks = ClockSampler(local_rank)
    if rank == 0:
        clocks.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    kern_ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    barrier()
    clocks.window_begin()
    ev0.record()
    for e in range(args.steps):
        kern_ev[e][0].record()
        engine.bpr_epoch(data, W["n_items"], U, V, B, W["lr"], W["reg"], W["use_bias"], key + rank, args.warmup + e,
                         stats, atomic=bool(args.atomic))
        kern_ev[e][1].record()
        if sync is not None:
            sync.exchange()
    ev1.record()
    barrier()
    clocks.window_end()
    clk = clocks.stop() if rank == 0 else None
    ms_total = ev0.elapsed_time(ev1)
    kern_ms = float(np.mean([a.elapsed_time(b) for a, b in kern_ev]))
    correct, skipped = stats.cpu().tolist()
    t = torch.tensor([ms_total, float(nnz * args.steps - skipped), float(skipped), kern_ms], dtype=torch.float64, device=dev)
    if world > 1:
        tmax = t.clone(); dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        tsum = t.clone(); dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
        ms_total, updates, skipped_all, kern_ms = tmax[0].item(), tsum[1].item(), tsum[2].item(), tmax[3].item()
    else:
        updates, skipped_all = t[1].item(), t[2].item()
    value = updates / (ms_total * 1e-3)

    # ---- roofline of the dominant kernel (bpr_hogwild_kernel), this rank's launches
    peak, peak_src = measured_peaks()
    upd_per_launch = (nnz * args.steps - skipped) / args.steps
    alg_bytes = algorithmic_bytes(k, upd_per_launch, skipped / args.steps, mean_deg)
    achieved = alg_bytes / (kern_ms * 1e-3) / 1e9
    traffic = None
    prof = os.path.join(ROOT, "profiles", "bpr_hogwild_dram_bytes.json")
    if os.path.exists(prof) and args.workload == "c2" and args.scale == 1.0:      # the capture is of this workload
        try:
            traffic = json.load(open(prof)).get("dram_bytes_per_launch")
        except Exception:
            traffic = None
    roofline = {"kernel": "bpr_hogwild_chunk_kernel<G=%d,NPL=1,VEC,%s>" % (min(32, max(4, k // 4)), "ATOMIC" if args.atomic else "PLAIN"), "bound": "hbm", "achieved": round(achieved, 1),
                "peak": peak, "unit": "GB/s", "frac": round(achieved / peak, 4), "traffic": traffic,
                "peak_source": peak_src, "algorithmic_bytes_per_update": 24 * k + 32 + 4 * math.ceil(math.log2(mean_deg + 1)),
                "kernel_ms": round(kern_ms, 3)}

    # ---- e2e: host buffers through the plug-in's training entry, copies inside the timed region
    e2e = None
    if not args.no_e2e:
        e2e = run_e2e(args, W, engine, indptr, indices, dev, world, rank)

    # ---- secondary metric: ranked users/s (score + exclusion + top-k) on the trained model
    # (users shard across the GPUs with the item side replicated and no collective: every rank ranks its own users,
    # the time is the max over ranks)
    def over_ranks(ms):
        if world == 1:
            return ms
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    rank_metric = None
    if not args.no_rank:
        rank_metric = run_rank(W, engine, data, U, V, B, dev, world, over_ranks)

    mf_metric = None
    rank_c5 = None
    if not args.no_rank:
        if rank == 0:
            mf_metric = run_mf(W, engine, data, dev)
        if args.workload == "c2":
            del data
            torch.cuda.empty_cache()
            rank_c5 = run_rank_c5(engine, dev, world, over_ranks)

    # ---- CPU baseline on rank 0, N = 1 only
    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        ip, ix = cpu_sample_csr(indptr.cpu().numpy(), indices.cpu().numpy())
        run, kind, cores = reference_fit_sgd_runner(ip, ix, W["n_items"], k, W["lr"], W["reg"], 0)
        v, secs, eps = time_cpu_epochs(run, len(ix))
        cpu_baseline = {"value": round(v, 1), "unit": "triplet-updates/s", "cores": cores, "kind": kind,
                        "sample": "first %d users of the same matrix: %d interactions, %d items, k=%d, %d epoch(s) in %.1fs"
                                  % (len(ip) - 1, len(ix), W["n_items"], k, eps, secs)}

    if rank == 0:
        out = {
            "metric": "BPR triplet-updates/sec", "value": round(value, 1), "unit": "updates/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(ms_total / args.steps, 3),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": ("BASELINE.json configs[1]" if args.workload == "c2" else "1/8 of BASELINE.json configs[2]")
                                   + ": %d users x %d items x %d interactions per GPU, BPR k=%d, "
                                   "lr=%g reg=%g use_bias; user activity log-normal, item popularity Zipf(1.0), unique pairs"
                                   % (W["n_users"], W["n_items"], nnz, k, W["lr"], W["reg"]),
                       "step": "one epoch = nnz sampled triplets (Hogwild, on-device Philox sampling)",
                       "l2": "working set (U %d MB + pair store %d MB + membership table per GPU) exceeds the 126 MB L2; no flush needed"
                             % (W["n_users"] * k * 4 // 1000000, nnz * 8 // 1000000),
                       "parallelism": "users sharded x%d, items replicated, 1 all-reduce of item deltas per epoch" % world,
                       "scatter": "red.global.add.v4.f32" if args.atomic else "st.global.cg.v4.f32 (Hogwild)"},
            "samples_per_s": round((nnz * args.steps * world) / (ms_total * 1e-3), 1),
            "skipped_frac": round(skipped_all / (nnz * args.steps * world), 5),
            "gpu_launches": args.steps * (1 + (2 * 2 if world > 1 else 0)),
            "clocks": clk, "roofline": roofline, "cpu_baseline": cpu_baseline, "e2e": e2e, "rank": rank_metric, "rank_c5": rank_c5, "mf": mf_metric,
        }
        if args.scale != 1.0:
            out["INVALID"] = "scaled-down debug run (--scale %g)" % args.scale
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def run_e2e(args, W, engine, indptr, indices, dev, world, rank):
    import torch
    import torch.distributed as dist
    k = W["k"]
    pin = lambda t: t.cpu().pin_memory()
    h_indptr, h_indices = pin(indptr), pin(indices)
    U, V, B = init_factors(W["n_users"], W["n_items"], k, seed=7, device=dev)
    hU, hV, hB = pin(U), pin(V), pin(B)
    del U, V, B
    torch.cuda.empty_cache()
    nnz = h_indices.numel()
    h2d = sum(t.numel() * t.element_size() for t in (h_indptr, h_indices, hU, hV, hB))
    d2h = sum(t.numel() * t.element_size() for t in (hU, hV, hB)) + 16
    steps = max(2, min(args.steps, 3))

    def one(e):
        hist, _ = engine.bpr_train_host(h_indptr.numpy(), h_indices.numpy(), W["n_items"], hU.numpy(), hV.numpy(),
                                        hB.numpy(), W["lr"], W["reg"], W["use_bias"], 1, key=77 + e + rank,
                                        atomic=bool(args.atomic), on_epoch=lambda *a: None, replica_sync=world > 1)
        return hist[0]

    one(0)                                             # warm-up (allocator, pinned staging)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    upd = 0
    for e in range(steps):
        c, s = one(1 + e)
        upd += nnz - s
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    t = torch.tensor([dt, float(upd)], dtype=torch.float64, device=dev)
    if world > 1:
        tm = t.clone(); dist.all_reduce(tm, op=dist.ReduceOp.MAX)
        ts = t.clone(); dist.all_reduce(ts, op=dist.ReduceOp.SUM)
        dt, upd = tm[0].item(), ts[1].item()
    return {"value": round(upd / dt, 1), "unit": "updates/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
            "steps": steps, "ms_per_step": round(dt / steps * 1e3, 2),
            "path": "engine.bpr_train_host (the call BPR.fit makes): pinned host CSR + U/V/B -> H2D -> prepare -> 1 epoch"
                    + (" -> item-delta all-reduce" if world > 1 else "") + " -> D2H U/V/B + stats"}


def reference_rank_users_per_s(U_host, V_host, B_host, topk, budget_s=8.0):
    """ranked users/s of the UNMODIFIED reference on the host cores: cornac.models.BPR.rank(user, k=topk) -- score() =
    copy(B) + fast_dot(U[u], V) (recom_bpr.pyx:272-297, OpenMP over the items) + argpartition top-k
    (recommender.py:476-530) -- called once per user like ranking_eval does, for a bounded number of users.
    Returns (users/s, users timed, seconds) or None when the compiled reference is not importable."""
    ref = os.path.join(ROOT, "oracle", "_ref")
    if not os.path.isdir(os.path.join(ref, "cornac")):
        return None
    if ref not in sys.path:
        sys.path.insert(0, ref)
    from cornac.models import BPR as RefBPR
    m = RefBPR(k=U_host.shape[1])
    m.num_users, m.num_items = U_host.shape[0], V_host.shape[0]
    m.u_factors = np.ascontiguousarray(U_host, dtype=np.float32)
    m.i_factors = np.ascontiguousarray(V_host, dtype=np.float32)
    m.i_biases = np.ascontiguousarray(B_host, dtype=np.float32)
    m.rank(0, k=topk)                                     # warm-up (thread pool, page faults)
    n, t0 = 0, time.perf_counter()
    while n < U_host.shape[0]:
        m.rank(n, k=topk)
        n += 1
        if (n & 15) == 0 and time.perf_counter() - t0 > budget_s:
            break
    secs = time.perf_counter() - t0
    return n / secs, n, secs


def run_rank(W, engine, data, U, V, B, dev, world=1, over_ranks=lambda ms: ms):
    """ranked users/s: score + exclusion of train positives + top-100 for a batch of users (tensor-core fused
    kernel), device-resident request (`value`) and through the host-buffer entry (`e2e`)."""
    import torch
    from cornac_b200._lib import load
    L = load()
    n_q, topk, k = RANK_WORKLOAD["n_q"], RANK_WORKLOAD["topk"], W["k"]
    n_q = min(n_q, W["n_users"])
    uidx = torch.arange(n_q, device=dev, dtype=torch.int64)
    ex_ptr = data.indptr[: n_q + 1].to(torch.int64).contiguous()
    ex_idx = data.indices[: int(ex_ptr[-1].item())].contiguous()
    nb = int(L.b200_rank_topk_workspace_bytes(n_q, W["n_items"], k, topk))
    ws = torch.empty(max(nb, 16), dtype=torch.uint8, device=dev)

    def go():
        return engine.rank_topk(U, V, topk, user_idx=uidx, item_base=B, excl_indptr=ex_ptr, excl_indices=ex_idx, workspace=ws)
    go()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        go()
    e1.record()
    torch.cuda.synchronize()
    if world > 1:
        dist_barrier()
    ms = over_ranks(e0.elapsed_time(e1) / 3)
    # end to end: pinned host request (user ids + exclusion CSR) -> H2D -> kernels -> D2H ids + scores
    h_u = uidx.cpu().pin_memory()
    h_p, h_i = ex_ptr.cpu().pin_memory(), ex_idx.cpu().pin_memory()
    o_i = torch.empty((n_q, topk), dtype=torch.int32).pin_memory()
    o_s = torch.empty((n_q, topk), dtype=torch.float32).pin_memory()

    def go_host():
        engine.rank_topk_host(U, V, topk, h_u.numpy(), item_base=B, excl_indptr=h_p.numpy(), excl_indices=h_i.numpy(),
                              out_ids=o_i.numpy(), out_scores=o_s.numpy(), workspace=ws)
    go_host()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(3):
        go_host()
    torch.cuda.synchronize()
    ms_h = over_ranks((time.perf_counter() - t0) / 3 * 1e3)
    h2d = h_u.numel() * 8 + h_p.numel() * 8 + h_i.numel() * 4
    cpu = None
    if world == 1:
        try:                                                # a reported baseline, never a reason for the bench line to fail
            r = reference_rank_users_per_s(U[:4096].cpu().numpy(), V.cpu().numpy(), B.cpu().numpy(), topk)
            if r is not None:
                cpu = {"value": round(r[0], 1), "unit": "users/s", "cores": os.cpu_count(), "kind": "reference",
                       "sample": "cornac.models.BPR.rank(u, k=%d) of the compiled reference for %d users of the same model "
                                 "(%d items, k=%d) in %.1fs, no exclusion list" % (topk, r[1], W["n_items"], k, r[2])}
        except Exception as exc:                            # noqa: BLE001
            cpu = {"unavailable": "%s: %s" % (type(exc).__name__, exc)}
    return {"metric": "ranked users/sec", "cpu_baseline": cpu, "value": round(world * n_q / (ms * 1e-3), 1), "unit": "users/s", "n_gpus": world,
            "config": "%d users per GPU x %d items k=%d top-%d, train positives excluded (the bench's BPR model)" % (n_q, W["n_items"], k, topk),
            "ms": round(ms, 3), "tflops": round(world * 2.0 * k * W["n_items"] * n_q / (ms * 1e-3) / 1e12, 2),
            "e2e": {"value": round(world * n_q / (ms_h * 1e-3), 1), "unit": "users/s", "h2d_bytes_per_step": int(h2d),
                    "d2h_bytes_per_step": int(n_q * topk * 8), "ms": round(ms_h, 3),
                    "path": "engine.rank_topk_host: pinned user ids + exclusion CSR -> H2D -> b200_rank_topk -> D2H ids + scores"}}


def dist_barrier():
    import torch
    import torch.distributed as dist
    dist.barrier()
    torch.cuda.synchronize()


def run_rank_c5(engine, dev, world=1, over_ranks=lambda ms: ms):
    """ranked users/s on the item side of BASELINE.json configs[4] (1 M items, k = 128, top-100, 100 seen items
    excluded per user): one call of b200_rank_topk for 75 776 users, random N(0, 0.1) factors and biases."""
    import torch
    from cornac_b200._lib import load
    L = load()
    n_items, k, n_q, topk, n_excl = 1_000_000, 128, 75776, 100, 100
    g = torch.Generator(device=dev)
    g.manual_seed(11)
    U = torch.randn((n_q, k), generator=g, device=dev) * 0.1
    V = torch.randn((n_items, k), generator=g, device=dev) * 0.1
    B = torch.randn(n_items, generator=g, device=dev) * 0.1
    ex = torch.randint(0, n_items, (n_q, n_excl), generator=g, device=dev, dtype=torch.int32)
    ex_idx = torch.sort(ex, dim=1)[0].contiguous().view(-1)
    ex_ptr = (torch.arange(n_q + 1, device=dev, dtype=torch.int64) * n_excl).contiguous()
    nb = int(L.b200_rank_topk_workspace_bytes(n_q, n_items, k, topk))
    ws = torch.empty(max(nb, 16), dtype=torch.uint8, device=dev)

    def go():
        return engine.rank_topk(U, V, topk, item_base=B, excl_indptr=ex_ptr, excl_indices=ex_idx, workspace=ws)
    go()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        go()
    e1.record()
    torch.cuda.synchronize()
    if world > 1:
        dist_barrier()
    ms = over_ranks(e0.elapsed_time(e1) / 3)
    tf_peak = tensor_peak() * world
    tfl = world * 2.0 * k * n_items * n_q / (ms * 1e-3) / 1e12
    return {"metric": "ranked users/sec", "value": round(world * n_q / (ms * 1e-3), 1), "unit": "users/s", "n_gpus": world,
            "config": "%d users per GPU x %d items k=%d top-%d, %d excluded items per user (BASELINE.json configs[4] item side)"
                      % (n_q, n_items, k, topk, n_excl), "ms": round(ms, 3),
            "roofline": {"kernel": "rank_tc_kernel (tcgen05 fp16 -> f32) + finish", "bound": "tensor", "achieved": round(tfl, 1),
                         "peak": tf_peak, "unit": "TFLOP/s", "frac": round(tfl / tf_peak, 4) if tf_peak else None,
                         "note": "2*k*n_items flop per user over the whole call (pack + tensor pass + exact finish)"}}


def reference_mf_ratings_per_s(rid, cid, val, n_users, n_items, k, budget_s=8.0):
    """ratings/s of the UNMODIFIED reference kernel backend_cpu.fit_sgd (cornac/models/mf/backend_cpu.pyx:35-97) with all
    host threads on a bounded rating sample.  Returns (ratings/s, threads, epochs, seconds) or None."""
    ref = os.path.join(ROOT, "oracle", "_ref")
    if not os.path.isdir(os.path.join(ref, "cornac")):
        return None
    if ref not in sys.path:
        sys.path.insert(0, ref)
    import multiprocessing
    from cornac.models.mf import backend_cpu
    rng = np.random.RandomState(5)
    n_u = int(rid.max()) + 1 if len(rid) else 1
    U = rng.normal(0, 0.01, (n_u, k)).astype(np.float32)
    V = rng.normal(0, 0.01, (n_items, k)).astype(np.float32)
    Bu, Bi = np.zeros(n_u, np.float32), np.zeros(n_items, np.float32)
    rid64, cid64 = np.ascontiguousarray(rid, dtype=np.int64), np.ascontiguousarray(cid, dtype=np.int64)
    val32 = np.ascontiguousarray(val, dtype=np.float32)
    threads = multiprocessing.cpu_count()
    backend_cpu.fit_sgd(rid64[:100000], cid64[:100000], val32[:100000], U, V, Bu, Bi, 0.01, 0.02, 3.0, 1, threads, True, False, False)
    epochs, t0 = 0, time.perf_counter()
    while epochs < 1 or time.perf_counter() - t0 < budget_s / 2:
        backend_cpu.fit_sgd(rid64, cid64, val32, U, V, Bu, Bi, 0.01, 0.02, 3.0, 1, threads, True, False, False)
        epochs += 1
    secs = time.perf_counter() - t0
    return epochs * len(val32) / secs, threads, epochs, secs


def run_mf(W, engine, data, dev):
    """secondary metric: MF ratings/s (b200_mf_epoch, Hogwild + atomic scatter) on the same interaction matrix
    with synthetic ratings in {1..5}, stored by user (CSR order), k = 128 as in BASELINE.json configs[3]."""
    import torch
    k = 128
    g = torch.Generator(device=dev)
    g.manual_seed(5)
    n = data.nnz
    rid = data.coo_row
    cid = data.indices
    val = torch.randint(1, 6, (n,), generator=g, device=dev).float()
    U = torch.randn((W["n_users"], k), generator=g, device=dev) * 0.01
    V = torch.randn((W["n_items"], k), generator=g, device=dev) * 0.01
    Bu, Bi = torch.zeros(W["n_users"], device=dev), torch.zeros(W["n_items"], device=dev)
    loss = torch.zeros(1, device=dev)
    for _ in range(2):
        engine.mf_epoch(rid, cid, val, U, V, Bu, Bi, 0.01, 0.02, 3.0, True, loss)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        engine.mf_epoch(rid, cid, val, U, V, Bu, Bi, 0.01, 0.02, 3.0, True, loss)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    peak, _ = measured_peaks()
    gbs = n * (16 * k + 28) / (ms * 1e-3) / 1e9
    cpu = None
    try:                                                    # a reported baseline, never a reason for the bench line to fail
        n_s = min(n, 10_000_000)                            # bounded sample: the first 10 M ratings (users in CSR order)
        r = reference_mf_ratings_per_s(rid[:n_s].cpu().numpy(), cid[:n_s].cpu().numpy(), val[:n_s].cpu().numpy(),
                                       W["n_users"], W["n_items"], k)
        if r is not None:
            cpu = {"value": round(r[0], 1), "unit": "ratings/s", "cores": r[1], "kind": "reference",
                   "sample": "backend_cpu.fit_sgd of the compiled reference, %d epoch(s) over the first %d ratings of the same "
                             "list (%d items, k=%d) in %.1fs" % (r[2], n_s, W["n_items"], k, r[3])}
    except Exception as exc:                                # noqa: BLE001
        cpu = {"unavailable": "%s: %s" % (type(exc).__name__, exc)}
    return {"metric": "MF ratings/sec", "cpu_baseline": cpu, "value": round(n / (ms * 1e-3), 1), "unit": "ratings/s",
            "config": "%d users x %d items x %d ratings, k=%d, use_bias, Hogwild + red.global.add" % (W["n_users"], W["n_items"], n, k),
            "ms_per_epoch": round(ms, 3),
            "roofline": {"bound": "hbm", "achieved": round(gbs, 1), "peak": peak, "unit": "GB/s", "frac": round(gbs / peak, 4),
                         "algorithmic_bytes_per_rating": 16 * k + 28}}


def run_reference_arm(args, W):
    """--impl reference: the reference's own CPU kernel on a bounded sample of the workload."""
    import torch
    k = W["k"]
    if torch.cuda.is_available():
        indptr, indices = synth_interactions(W["n_users"], W["n_items"], W["nnz"], seed=1234, device=torch.device("cuda", 0))
        ip, ix = cpu_sample_csr(indptr.cpu().numpy(), indices.cpu().numpy())
        del indptr, indices
    else:   # no GPU: generate the sample directly at sample size
        frac = CPU_SAMPLE["n_users"] / W["n_users"]
        ipt, ixt = synth_interactions(CPU_SAMPLE["n_users"], W["n_items"], int(W["nnz"] * frac), seed=1234, device=torch.device("cpu"))
        ip, ix = ipt.numpy(), ixt.numpy()
    run, kind, cores = reference_fit_sgd_runner(ip, ix, W["n_items"], k, W["lr"], W["reg"], 0)
    nnz = len(ix)
    for _ in range(max(1, min(args.warmup, 2))):
        run()
    t0 = time.perf_counter()
    upd = 0
    steps = args.steps
    for _ in range(steps):
        c, s = run()
        upd += nnz - s
    dt = time.perf_counter() - t0
    v = upd / dt
    sample = "first %d users of the configs[1] matrix: %d interactions, %d items, k=%d" % (len(ip) - 1, nnz, W["n_items"], k)
    out = {"impl": "reference", "metric": "BPR triplet-updates/sec", "value": round(v, 1), "unit": "updates/s",
           "n_gpus": args.gpus, "steps": steps, "warmup": args.warmup, "ms_per_step": round(dt / steps * 1e3, 2),
           "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
           "config": {"workload": "BASELINE.json configs[1] (1M x 100K x 100M, BPR k=64), each step = one _fit_sgd epoch "
                                  "over a bounded sample: " + sample},
           "cpu_baseline": {"value": round(v, 1), "unit": "updates/s", "cores": cores, "kind": kind, "sample": sample},
           "e2e": {"value": round(v, 1), "unit": "updates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
           "gpu_launches": 0}
    print(json.dumps(out), flush=True)
    return 0


if __name__ == "__main__":
    sys.exit(main())
