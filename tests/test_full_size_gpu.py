"""Size-independent properties at BASELINE.json's full sizes -- configs[1] (1M users x 100K items x 100M interactions,
k=64) and one GPU's share of configs[2] / [3] / [4] (1.25M users x 1M items x 125M interactions, k=128: the item side,
factor width and per-GPU interaction count of the 8-GPU target).  Where the oracle cannot replay whole epochs in
seconds, the kernels are checked through identities that hold at any size, and against the oracle on slices.
GPU only (needs ~20 GB of HBM, ~2 minutes)."""
import os
import sys

import numpy as np
import pytest

from conftest import ROOT
from oracle import oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def c2():
    import torch
    sys.path.insert(0, ROOT)
    import bench
    from cornac_b200 import engine
    dev = torch.device("cuda", 0)
    W = bench.WORKLOADS["c2"]
    indptr, indices = bench.synth_interactions(W["n_users"], W["n_items"], W["nnz"], 1234, dev)
    data = engine.BprData(indptr, indices).prepare()
    U, V, B = _init_factors(W["n_users"], W["n_items"], W["k"], 99, dev)
    return dict(W=W, indptr=indptr, indices=indices, data=data, U=U, V=V, B=B, dev=dev)


def _init_factors(n_users, n_items, k, seed, dev):
    import torch
    g = torch.Generator(device=dev).manual_seed(seed)
    U = (torch.rand((n_users, k), generator=g, device=dev) - 0.5) / k
    V = (torch.rand((n_items, k), generator=g, device=dev) - 0.5) / k
    return U, V, torch.zeros(n_items, device=dev)


@pytest.fixture(scope="module")
def c3():
    """One GPU's share of configs[2] at 8 GPUs: block 0 of bench.py's model (1.25M users x 1M items x 125M, k=128)."""
    import torch
    sys.path.insert(0, ROOT)
    import bench
    from cornac_b200 import engine
    dev = torch.device("cuda", 0)
    W = dict(bench.WORKLOADS["c3"])
    indptr, indices = bench.synth_shard(W, [0], dev)
    data = engine.BprData(indptr, indices).prepare()
    U = bench.init_user_factors(W, [0], dev)
    V, B = bench.init_item_factors(W, dev)
    return dict(W=W, indptr=indptr, indices=indices, data=data, U=U, V=V, B=B, dev=dev)


def test_prepare_is_a_permutation_free_copy_and_a_complete_set(c2):
    """pairs == (COO row, column) of every interaction; the table holds exactly nnz distinct keys."""
    import torch
    d = c2["data"]
    assert torch.equal(d.pairs[:, 1], c2["indices"])
    assert torch.equal(d.pairs[:, 0], d.coo_row)
    t = d.table
    n_keys = int((t != -1).sum().item())
    assert n_keys == d.nnz
    # checksum of checksums: xor / sum of stored keys == xor / sum of (u << 32 | i) over the matrix
    keys = (d.pairs[:, 0].to(torch.int64) << 32) | d.pairs[:, 1].to(torch.int64)
    stored = t[t != -1]
    assert int(stored.sum().item()) == int(keys.sum().item())
    assert d.nnz == c2["W"]["nnz"] and int(c2["indptr"][-1].item()) == d.nnz


def test_epoch_with_lr0_is_identity_and_counts_match_oracle_on_a_stream_slice(c2):
    """lr = 0 over a full 100M-sample epoch leaves every factor bit-identical; (correct, skipped) of a 4M-sample
    slice of the same epoch equal the oracle's counts on that slice of the Philox stream."""
    import torch
    from cornac_b200 import engine
    W, data = c2["W"], c2["data"]
    U, V, B = c2["U"].clone(), c2["V"].clone(), c2["B"].clone()
    B += torch.randn(B.shape, device=B.device, generator=torch.Generator(device=B.device).manual_seed(3)) * 0.01
    U0, V0, B0 = U.clone(), V.clone(), B.clone()
    stats = torch.zeros(2, dtype=torch.int64, device=c2["dev"])
    engine.bpr_epoch(data, W["n_items"], U, V, B, 0.0, 0.01, True, 4242, 1, stats, exact_exp=True)
    c_all, s_all = stats.cpu().tolist()
    assert torch.equal(U, U0) and torch.equal(V, V0) and torch.equal(B, B0)
    assert 0 < s_all < 0.02 * data.nnz and 0 < c_all < data.nnz
    n, base = 4_000_000, 37_000_000
    stats.zero_()
    engine.bpr_epoch(data, W["n_items"], U, V, B, 0.0, 0.01, True, 4242, 1, stats, n_samples=n, sample_base=base, exact_exp=True)
    c, s = stats.cpu().tolist()
    ii, jj = engine.bpr_draw_host(4242, 1, n, data.nnz, W["n_items"], sample_base=base)
    indptr, indices = c2["indptr"].cpu().numpy(), c2["indices"].cpu().numpy()
    Uh, Vh, Bh = U0.cpu().numpy(), V0.cpu().numpy(), B0.cpu().numpy()
    c_ref, s_ref = O.bpr_replay(ii, jj, indptr, indices, Uh, Vh, Bh, 0.0, 0.01, True)
    assert s == s_ref and abs(c - c_ref) <= max(3, int(2e-5 * n))


def test_full_epoch_conserves_item_mass_with_atomic_scatter(c2):
    """reg = 0: every update adds +d to V[i], B[i] and -d to V[j], B[j] => column sums of V and the sum of B are
    invariants of a whole 100M-sample epoch when no update is lost (red.global.add)."""
    import torch
    from cornac_b200 import engine
    W, data = c2["W"], c2["data"]
    U, V, B = c2["U"].clone(), c2["V"].clone(), c2["B"].clone()
    col0 = V.double().sum(0)
    stats = torch.zeros(2, dtype=torch.int64, device=c2["dev"])
    engine.bpr_epoch(data, W["n_items"], U, V, B, 0.05, 0.0, True, 7, 0, stats, atomic=True)
    moved = (V.double() - c2["V"].double()).abs().sum().item()
    assert moved > 1e3
    assert (V.double().sum(0) - col0).abs().max().item() < 1e-5 * moved / W["k"] + 1e-3
    assert abs(B.double().sum().item()) < 1e-5 * B.double().abs().sum().item() + 1e-3
    c, s = stats.cpu().tolist()
    assert c + s <= data.nnz and abs(s / data.nnz - 0.0022) < 0.002      # skip rate of this matrix


def test_fused_rank_equals_exact_path_at_full_catalogue(c2, monkeypatch):
    """tensor-core rank == exact rank (ids and scores) on 2 048 users x the full 100K-item catalogue with each
    user's train positives excluded; and ranking is idempotent."""
    import torch
    from cornac_b200 import engine
    W, data = c2["W"], c2["data"]
    g = torch.Generator(device=c2["dev"]).manual_seed(11)
    U = torch.randn(c2["U"].shape, device=c2["dev"], generator=g) * 0.1
    V = torch.randn(c2["V"].shape, device=c2["dev"], generator=g) * 0.1
    B = torch.randn(c2["B"].shape, device=c2["dev"], generator=g) * 0.1
    users = torch.arange(5000, 5000 + 2048, device=c2["dev"], dtype=torch.int64)
    lo, hi = int(c2["indptr"][5000].item()), int(c2["indptr"][5000 + 2048].item())
    ex_ptr = (c2["indptr"][5000:5000 + 2049].to(torch.int64) - lo).contiguous()
    ex_idx = c2["indices"][lo:hi].contiguous()
    a = engine.rank_topk(U, V, 100, user_idx=users, item_base=B, excl_indptr=ex_ptr, excl_indices=ex_idx)
    a2 = engine.rank_topk(U, V, 100, user_idx=users, item_base=B, excl_indptr=ex_ptr, excl_indices=ex_idx)
    monkeypatch.setenv("B200_RANK_TC", "0")
    b = engine.rank_topk(U, V, 100, user_idx=users, item_base=B, excl_indptr=ex_ptr, excl_indices=ex_idx)
    assert torch.equal(a[0], b[0]) and torch.equal(a[1], b[1])
    assert torch.equal(a[0], a2[0]) and torch.equal(a[1], a2[1])
    # sortedness + exclusion honoured
    sc = a[1].cpu().numpy()
    assert np.all(sc[:, :-1] >= sc[:, 1:])
    ids = a[0].cpu().numpy()
    exp, exi = ex_ptr.cpu().numpy(), ex_idx.cpu().numpy()
    for q in range(0, 2048, 97):
        assert not np.intersect1d(ids[q], exi[exp[q]:exp[q + 1]]).size


# ------------------------------------------------------------------------------------------------------------------
# the target shape: 1 M items, k = 128 (configs[2] / [3] / [4], one GPU's share at 8 GPUs)
def test_c3_prepare_is_complete(c3):
    import torch
    d = c3["data"]
    assert d.nnz == c3["W"]["nnz"] // 8 and d.n_users == c3["W"]["n_users"] // 8
    assert torch.equal(d.pairs[:, 1], c3["indices"]) and torch.equal(d.pairs[:, 0], d.coo_row)
    t = d.table
    assert int((t != -1).sum().item()) == d.nnz
    keys = (d.pairs[:, 0].to(torch.int64) << 32) | d.pairs[:, 1].to(torch.int64)
    assert int(t[t != -1].sum().item()) == int(keys.sum().item())


def test_c3_epoch_with_lr0_is_identity_and_slice_counts_match_the_oracle(c3):
    """k = 128 / 1 M items: lr = 0 over the full 125M-sample epoch leaves every factor bit-identical (32-lane groups, V not
    L2-resident); (correct, skipped) of a 2M-sample slice equal the oracle's counts on that slice of the Philox stream."""
    import torch
    from cornac_b200 import engine
    W, data = c3["W"], c3["data"]
    U, V, B = c3["U"].clone(), c3["V"].clone(), c3["B"].clone()
    B += torch.randn(B.shape, device=B.device, generator=torch.Generator(device=B.device).manual_seed(3)) * 0.01
    U0, V0, B0 = U.clone(), V.clone(), B.clone()
    stats = torch.zeros(2, dtype=torch.int64, device=c3["dev"])
    engine.bpr_epoch(data, W["n_items"], U, V, B, 0.0, 0.01, True, 777, 2, stats, exact_exp=True)
    c_all, s_all = stats.cpu().tolist()
    assert torch.equal(U, U0) and torch.equal(V, V0) and torch.equal(B, B0)
    assert 0 < s_all < 0.02 * data.nnz and 0 < c_all < data.nnz
    n, base = 2_000_000, 91_000_000
    stats.zero_()
    engine.bpr_epoch(data, W["n_items"], U, V, B, 0.0, 0.01, True, 777, 2, stats, n_samples=n, sample_base=base, exact_exp=True)
    c, s = stats.cpu().tolist()
    ii, jj = engine.bpr_draw_host(777, 2, n, data.nnz, W["n_items"], sample_base=base)
    indptr, indices = c3["indptr"].cpu().numpy(), c3["indices"].cpu().numpy()
    Uh, Vh, Bh = U0.cpu().numpy(), V0.cpu().numpy(), B0.cpu().numpy()
    c_ref, s_ref = O.bpr_replay(ii, jj, indptr, indices, Uh, Vh, Bh, 0.0, 0.01, True)
    assert s == s_ref and abs(c - c_ref) <= max(3, int(2e-5 * n))


def test_c3_full_epoch_conserves_item_mass_and_trains(c3):
    """reg = 0 at k = 128 / 1 M items: column sums of V and the sum of B are invariants of a whole 125M-sample epoch (no lost
    update with red.global.add.v4), the skip rate is the matrix's, and a second epoch raises the pairwise accuracy."""
    import torch
    from cornac_b200 import engine
    W, data = c3["W"], c3["data"]
    U, V, B = c3["U"].clone(), c3["V"].clone(), c3["B"].clone()
    col0 = V.double().sum(0)
    stats = torch.zeros(2, dtype=torch.int64, device=c3["dev"])
    engine.bpr_epoch(data, W["n_items"], U, V, B, 0.05, 0.0, True, 7, 0, stats, atomic=True)
    moved = (V.double() - c3["V"].double()).abs().sum().item()
    assert moved > 1e3
    assert (V.double().sum(0) - col0).abs().max().item() < 1e-5 * moved / W["k"] + 1e-3
    assert abs(B.double().sum().item()) < 1e-5 * B.double().abs().sum().item() + 1e-3
    c1, s1 = stats.cpu().tolist()
    assert c1 + s1 <= data.nnz and s1 / data.nnz < 0.01
    stats.zero_()
    engine.bpr_epoch(data, W["n_items"], U, V, B, 0.05, 0.0, True, 7, 1, stats, atomic=True)
    c2_, s2 = stats.cpu().tolist()
    assert c2_ / (data.nnz - s2) > c1 / (data.nnz - s1) + 0.002


def test_c3_fused_rank_equals_exact_path_and_oracle_at_1m_items_k128(c3, monkeypatch):
    """configs[4] shape: tensor-core rank == exact rank (ids AND scores) on 2 048 users x the full 1M-item catalogue at k = 128
    (3 907 MMA stages per user tile, joint threshold raises, 512-entry lists), train positives excluded; 8 of the users are
    also checked against the CPU oracle's f64-accumulated scores and total-order top-k; ranking is idempotent."""
    import torch
    from cornac_b200 import engine
    W = c3["W"]
    dev = c3["dev"]
    g = torch.Generator(device=dev).manual_seed(12)
    U = torch.randn((4096, W["k"]), device=dev, generator=g) * 0.1
    V = torch.randn(c3["V"].shape, device=dev, generator=g) * 0.1
    B = torch.randn(c3["B"].shape, device=dev, generator=g) * 0.1
    first = 7000
    users = torch.arange(1000, 1000 + 2048, device=dev, dtype=torch.int64)
    lo, hi = int(c3["indptr"][first].item()), int(c3["indptr"][first + 2048].item())
    ex_ptr = (c3["indptr"][first:first + 2049].to(torch.int64) - lo).contiguous()
    ex_idx = c3["indices"][lo:hi].contiguous()
    a = engine.rank_topk(U, V, 100, user_idx=users, item_base=B, excl_indptr=ex_ptr, excl_indices=ex_idx)
    a2 = engine.rank_topk(U, V, 100, user_idx=users, item_base=B, excl_indptr=ex_ptr, excl_indices=ex_idx)
    monkeypatch.setenv("B200_RANK_TC", "0")
    b = engine.rank_topk(U, V, 100, user_idx=users, item_base=B, excl_indptr=ex_ptr, excl_indices=ex_idx)
    assert torch.equal(a[0], b[0]) and torch.equal(a[1], b[1])
    assert torch.equal(a[0], a2[0]) and torch.equal(a[1], a2[1])
    ids, sc = a[0].cpu().numpy(), a[1].cpu().numpy()
    assert np.all(sc[:, :-1] >= sc[:, 1:]) and ids.min() >= 0
    exp, exi = ex_ptr.cpu().numpy(), ex_idx.cpu().numpy()
    Uh, Vh, Bh = U.cpu().numpy(), V.cpu().numpy(), B.cpu().numpy()
    for q in (0, 1, 500, 1023, 1024, 1500, 2046, 2047):
        want = O.score_batch(Uh[1000 + q:1001 + q], Vh, Bh)[0]
        wi, wsc, _ = O.topk(want, 100, excl=exi[exp[q]:exp[q + 1]])
        assert np.array_equal(ids[q], wi) and np.array_equal(sc[q], wsc), q
        assert not np.intersect1d(ids[q], exi[exp[q]:exp[q + 1]]).size


def test_c3_trained_model_ranks_identically_on_both_paths(c3, monkeypatch):
    """the rank leg of bench.py: a model trained for two epochs at the target shape, 1 024 of its users with their own train
    positives excluded -- fused == exact."""
    import torch
    from cornac_b200 import engine
    W, data = c3["W"], c3["data"]
    U, V, B = c3["U"].clone(), c3["V"].clone(), c3["B"].clone()
    stats = torch.zeros(2, dtype=torch.int64, device=c3["dev"])
    for e in range(2):
        engine.bpr_epoch(data, W["n_items"], U, V, B, 0.05, 0.01, True, 5, e, stats)
    n = 1024
    ex_ptr = c3["indptr"][: n + 1].to(torch.int64).contiguous()
    ex_idx = c3["indices"]
    a = engine.rank_topk(U[:n], V, 100, item_base=B, excl_indptr=ex_ptr, excl_indices=ex_idx)
    monkeypatch.setenv("B200_RANK_TC", "0")
    b = engine.rank_topk(U[:n], V, 100, item_base=B, excl_indptr=ex_ptr, excl_indices=ex_idx)
    assert torch.equal(a[0], b[0]) and torch.equal(a[1], b[1])


def test_c4_mf_lr0_is_identity_and_loss_matches_at_target_shape(c3):
    """configs[3] item / k shape: one Hogwild MF epoch with lr = 0 over 125M ratings leaves U, V, Bu, Bi bit-identical and
    returns sum(err^2) equal to the f64 evaluation of the same predictions (1e-5 relative); with lr > 0 the loss falls."""
    import torch
    from cornac_b200 import engine
    W, data, dev = c3["W"], c3["data"], c3["dev"]
    k = 128
    g = torch.Generator(device=dev).manual_seed(21)
    n = data.nnz
    rid, cid = data.coo_row, data.indices
    val = torch.randint(1, 6, (n,), generator=g, device=dev).float()
    U = torch.randn((data.n_users, k), generator=g, device=dev) * 0.05
    V = torch.randn((W["n_items"], k), generator=g, device=dev) * 0.05
    Bu = torch.randn(data.n_users, generator=g, device=dev) * 0.1
    Bi = torch.randn(W["n_items"], generator=g, device=dev) * 0.1
    U0, V0, Bu0, Bi0 = U.clone(), V.clone(), Bu.clone(), Bi.clone()
    loss = torch.zeros(1, device=dev)
    engine.mf_epoch(rid, cid, val, U, V, Bu, Bi, 0.0, 0.02, 3.0, True, loss)
    assert torch.equal(U, U0) and torch.equal(V, V0) and torch.equal(Bu, Bu0) and torch.equal(Bi, Bi0)
    want = 0.0
    step = 5_000_000
    for a in range(0, n, step):
        r, c = rid[a:a + step].long(), cid[a:a + step].long()
        pred = 3.0 + Bu[r].double() + Bi[c].double() + (U[r].double() * V[c].double()).sum(1)
        want += float(((val[a:a + step].double() - pred) ** 2).sum().item())
    got = float(loss.item())
    assert abs(got - want) < 1e-5 * want, (got, want)
    first = got
    for _ in range(2):
        engine.mf_epoch(rid, cid, val, U, V, Bu, Bi, 0.01, 0.02, 3.0, True, loss)
    assert float(loss.item()) < 0.995 * first


def test_c3_cache_blocked_order_draws_the_same_law_and_trains_like_the_unblocked_epoch(c3):
    """B200_BPR_BLOCKED at the target shape (plan 16 windows x 13 item blocks): (correct, skipped) of a slice that spans
    several (window, block) runs equal the oracle's counts on the host-evaluated blocked stream (b200_bpr_draw_host2);
    the slice's interactions / negatives really come from the windows / blocks of the plan; a blocked epoch conserves the
    item mass like the unblocked one and reaches the same pairwise accuracy (+-0.01) from the same start."""
    import torch
    from cornac_b200 import engine
    W, data = c3["W"], c3["data"]
    plan = engine.bpr_block_plan(data.n_users, W["n_items"], W["k"])
    assert plan[0] > 1 and plan[1] > 1
    s_sub = -(-data.nnz // (plan[0] * plan[1]))
    U, V, B = c3["U"].clone(), c3["V"].clone(), c3["B"].clone()
    stats = torch.zeros(2, dtype=torch.int64, device=c3["dev"])
    n, base = 1_500_000, 7 * s_sub - 400_000                       # crosses two run boundaries
    engine.bpr_epoch(data, W["n_items"], U, V, B, 0.0, 0.01, True, 99, 3, stats, n_samples=n, sample_base=base, exact_exp=True,
                     blocked=True)
    c, s = stats.cpu().tolist()
    assert torch.equal(U, c3["U"]) and torch.equal(V, c3["V"])
    ii, jj = engine.bpr_draw_host(99, 3, n, data.nnz, W["n_items"], sample_base=base, plan=plan)
    runs = (base + np.arange(n)) // s_sub
    for r in np.unique(runs):
        sel = runs == r
        w, b = int(r) // plan[1], (int(r) % plan[1] + 3) % plan[1]
        assert ii[sel].min() >= w * data.nnz // plan[0] - 1 and ii[sel].max() <= (w + 1) * data.nnz // plan[0] + 1
        assert jj[sel].min() >= b * W["n_items"] // plan[1] - 1 and jj[sel].max() <= (b + 1) * W["n_items"] // plan[1] + 1
    indptr, indices = c3["indptr"].cpu().numpy(), c3["indices"].cpu().numpy()
    Uh, Vh, Bh = c3["U"].cpu().numpy(), c3["V"].cpu().numpy(), c3["B"].cpu().numpy()
    c_ref, s_ref = O.bpr_replay(ii, jj, indptr, indices, Uh, Vh, Bh, 0.0, 0.01, True)
    assert s == s_ref and abs(c - c_ref) <= max(3, int(2e-5 * n))
    acc = {}
    for blocked in (False, True):
        U, V, B = c3["U"].clone(), c3["V"].clone(), c3["B"].clone()
        col0 = V.double().sum(0)
        for e in range(2):
            stats.zero_()
            engine.bpr_epoch(data, W["n_items"], U, V, B, 0.05, 0.0, True, 7, e, stats, blocked=blocked)
        cc, ss = stats.cpu().tolist()
        acc[blocked] = cc / (data.nnz - ss)
        moved = (V.double() - c3["V"].double()).abs().sum().item()
        assert (V.double().sum(0) - col0).abs().max().item() < 1e-5 * moved / W["k"] + 1e-3
        assert abs(ss / data.nnz - 0.00023) < 0.0005
    assert abs(acc[True] - acc[False]) < 0.01, acc
