"""Host-side pieces of bench.py that do not need a GPU: the synthetic interaction generator (run on the CPU device),
the algorithmic-byte accounting of SURVEY.md 8(d), the bounded CPU sample and the reference legs."""
import json
import os
import subprocess
import sys

import numpy as np
import torch

from conftest import ROOT, needs_cornac

sys.path.insert(0, ROOT)
import bench  # noqa: E402


def test_synthetic_interactions_are_a_valid_sorted_csr_of_unique_pairs():
    n_users, n_items, nnz = 3000, 500, 60000
    indptr, indices = bench.synth_interactions(n_users, n_items, nnz, seed=7, device=torch.device("cpu"))
    ip, ix = indptr.numpy(), indices.numpy()
    assert ip.dtype == np.int32 and ix.dtype == np.int32 and ip[0] == 0 and ip[-1] == nnz == len(ix)
    assert np.all(np.diff(ip) >= 0) and ix.min() >= 0 and ix.max() < n_items
    rows = np.repeat(np.arange(n_users), np.diff(ip))
    keys = rows.astype(np.int64) * n_items + ix
    assert np.all(np.diff(keys) > 0)                       # sorted rows, no duplicate (u, i)
    # skew: the busiest users / most popular items carry far more than the average
    deg = np.diff(ip)
    pop = np.bincount(ix, minlength=n_items)
    assert deg.max() > 5 * deg.mean() and pop.max() > 10 * pop.mean()
    # deterministic in the seed
    ip2, ix2 = bench.synth_interactions(n_users, n_items, nnz, seed=7, device=torch.device("cpu"))
    assert torch.equal(indptr, ip2) and torch.equal(indices, ix2)


def test_algorithmic_bytes_follow_survey_8d():
    # k=64, mean degree 100: 24k + 32 + 4 ceil(log2 101) = 1596 B per update; a skipped sample costs 16 + 28 B
    assert bench.algorithmic_bytes(64, 1, 0, 100) == 1596
    assert bench.algorithmic_bytes(128, 1, 0, 100) == 3132
    assert bench.algorithmic_bytes(64, 0, 1, 100) == 16 + 28
    assert bench.algorithmic_bytes(64, 10, 3, 100) == 10 * 1596 + 3 * 44


def test_cpu_sample_is_a_prefix_of_the_matrix():
    indptr, indices = bench.synth_interactions(2000, 300, 30000, seed=3, device=torch.device("cpu"))
    ip, ix = bench.cpu_sample_csr(indptr.numpy(), indices.numpy())
    n = len(ip) - 1
    assert ip[0] == 0 and ip[-1] == len(ix) and np.array_equal(ip, indptr.numpy()[: n + 1])
    assert np.array_equal(ix, indices.numpy()[: len(ix)])


@needs_cornac
def test_reference_legs_run_on_the_host():
    rng = np.random.RandomState(0)
    U = rng.normal(0, 0.1, (64, 16)).astype(np.float32)
    V = rng.normal(0, 0.1, (5000, 16)).astype(np.float32)
    B = rng.normal(0, 0.1, 5000).astype(np.float32)
    r = bench.reference_rank_users_per_s(U, V, B, 10, budget_s=0.2)
    assert r is not None and r[0] > 0 and 16 <= r[1] <= 64
    n = 50000
    rid = np.sort(rng.randint(2000, size=n)).astype(np.int32)
    cid = rng.randint(5000, size=n).astype(np.int32)
    val = rng.randint(1, 6, size=n).astype(np.float32)
    m = bench.reference_mf_ratings_per_s(rid, cid, val, 2000, 5000, 16, budget_s=0.2)
    assert m is not None and m[0] > 0 and m[2] >= 1


@needs_cornac
def test_reference_arm_prints_one_json_line_without_a_gpu():
    """`bench.py --impl reference` is pure host code: on a box without a GPU it still times the compiled reference on a
    (smaller) bounded sample and prints the contract's JSON line."""
    if torch.cuda.is_available():
        import pytest
        pytest.skip("covered by the GPU run of the reference arm")
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1",
                        "--workload", "c2", "--scale", "0.02"], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=600)
    assert p.returncode == 0, p.stderr[-2000:]
    lines = p.stdout.splitlines()
    assert len(lines) == 1, lines                          # stdout carries the JSON record and nothing else
    d = json.loads(lines[0])
    assert d["scaling"] == "strong" and d["cpu_baseline"]["host"]["threads_used"] == d["cpu_baseline"]["cores"] >= 1
    assert d["impl"] == "reference" and d["value"] > 0 and d["unit"] == "updates/s" and d["cpu_baseline"]["kind"] in ("reference", "port")
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0


def test_model_blocks_are_the_same_for_every_world_size():
    """Strong scaling shards ONE model: the union of the ranks' blocks is the same matrix for N = 1, 2, 4, 8, every rank owns
    the same number of interactions, and all blocks share one item-popularity law."""
    W = dict(bench.WORKLOADS["c2"])
    W["n_users"], W["nnz"] = 8000, 160000
    cpu = torch.device("cpu")
    whole_ptr, whole_idx = bench.synth_shard(W, bench.rank_blocks(0, 1), cpu)
    assert whole_ptr.numel() == W["n_users"] + 1 and whole_idx.numel() == W["nnz"] and int(whole_ptr[-1]) == W["nnz"]
    for world in (2, 4, 8):
        parts = [bench.synth_shard(W, bench.rank_blocks(r, world), cpu) for r in range(world)]
        assert all(p[1].numel() == W["nnz"] // world for p in parts)
        assert torch.equal(torch.cat([p[1] for p in parts]), whole_idx)
        deg = torch.cat([p[0][1:] - p[0][:-1] for p in parts])
        assert torch.equal(deg, whole_ptr[1:] - whole_ptr[:-1])
    # one popularity law: the most popular items of two different blocks largely coincide
    a = torch.bincount(bench.synth_shard(W, [0], cpu)[1].long(), minlength=W["n_items"]).topk(50).indices
    b = torch.bincount(bench.synth_shard(W, [5], cpu)[1].long(), minlength=W["n_items"]).topk(50).indices
    assert len(set(a.tolist()) & set(b.tolist())) >= 35


def test_host_cores_reports_what_the_process_owns():
    h = bench.host_cores()
    assert 1 <= h["threads_used"] <= h["sched_affinity"] <= (h["os_cpu_count"] or 10 ** 6)
    if h["cgroup_cpu_quota"] is not None:
        assert h["threads_used"] <= max(1, int(h["cgroup_cpu_quota"] + 0.5))
