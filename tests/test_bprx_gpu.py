"""VEBPR and SBPR on the GPU: the replay kernels against the oracle on explicit streams and against the
compiled-reference fixtures through the plug-ins; the Hogwild kernels through their invariants (lr = 0 identity, the
sample law's skip rate, learning planted structure).  Reference: cornac/models/bpr/recom_vebpr.pyx:214-337,
cornac/models/sbpr/recom_sbpr.pyx:193-300."""
from collections import OrderedDict

import numpy as np
import pytest

from conftest import golden, needs_cornac, rel_err, synth_csr

pytestmark = pytest.mark.gpu
TOL = 1e-4


def _view_csr(indptr, indices, n_items, seed, frac_users=0.7, per_user=8):
    """A 'viewed but not purchased' CSR with sorted rows; ~30 % of the users view nothing."""
    rng = np.random.RandomState(seed)
    n_users = len(indptr) - 1
    ptr, idx = [0], []
    for u in range(n_users):
        if rng.rand() < frac_users:
            own = set(indices[indptr[u]:indptr[u + 1]].tolist())
            v = sorted(set(rng.randint(n_items, size=per_user).tolist()) - own)
            idx.extend(v)
        ptr.append(len(idx))
    return np.asarray(ptr, np.int32), np.asarray(idx, np.int32)


def _social_lists(indptr, indices, seed):
    from oracle import oracle as O
    rng = np.random.RandomState(seed)
    n_users = len(indptr) - 1
    fr_ptr, fr = [0], []
    for u in range(n_users):
        if rng.rand() < 0.7 or u == n_users - 1:          # the last user has friends: every position the kernel reads exists
            f = sorted(set(rng.randint(n_users, size=rng.randint(1, 5)).tolist()) - {u})
            fr.extend(f)
        fr_ptr.append(len(fr))
    return O.sbpr_social_items(indptr, indices, np.asarray(fr_ptr, np.int32), np.asarray(fr, np.int32))


def _dev(a, dtype):
    import torch
    return torch.from_numpy(np.ascontiguousarray(a)).to(dtype).cuda()


@pytest.mark.parametrize("k", [8, 10, 64, 130 - 2])
def test_vebpr_replay_kernel_matches_oracle_on_a_stream(k):
    """b200_vebpr_epoch_replay == the sequential oracle on an explicit (i_index, v_id, j_id) stream, both loop bodies."""
    import torch
    from cornac_b200 import _lib
    from cornac_b200._lib import check, current_stream, ptr
    from oracle import oracle as O
    L = _lib.load()
    n_users, n_items, nnz = 200, 150, 4000
    indptr, indices = synth_csr(n_users, n_items, nnz, seed=3)
    nnz = len(indices)
    vptr, vidx = _view_csr(indptr, indices, n_items, seed=4)
    rng = np.random.RandomState(5)
    n = 5000
    ii = rng.randint(nnz, size=n).astype(np.int64)
    coo = O.coo_rows(indptr)
    u = coo[ii]
    nv = (vptr[u + 1] - vptr[u])
    v = np.where(nv > 0, vidx[np.minimum(vptr[u] + rng.randint(1 << 30, size=n) % np.maximum(nv, 1), len(vidx) - 1)], -1).astype(np.int32)
    j = rng.randint(n_items, size=n).astype(np.int32)
    assert (v < 0).any() and (v >= 0).any()
    U0 = ((rng.rand(n_users, k) - 0.5)).astype(np.float32)
    V0 = ((rng.rand(n_items, k) - 0.5)).astype(np.float32)
    Uo, Vo = U0.copy(), V0.copy()
    c, sk = O.vebpr_replay(ii, v, j, indptr, indices, vptr, vidx, Uo, Vo, 0.05, 0.01, 0.3)
    dU, dV = _dev(U0, torch.float32), _dev(V0, torch.float32)
    stats = torch.zeros(2, dtype=torch.int64, device="cuda")
    t = [_dev(ii, torch.int64), _dev(v, torch.int32), _dev(j, torch.int32), _dev(indptr, torch.int32), _dev(indices, torch.int32),
         _dev(coo, torch.int32), _dev(vptr, torch.int32), _dev(vidx, torch.int32)]          # kept alive until the kernel has run
    check(L.b200_vebpr_epoch_replay(ptr(t[0]), ptr(t[1]), ptr(t[2]), n, ptr(t[3]), ptr(t[4]), ptr(t[5]), ptr(t[6]), ptr(t[7]),
                                    ptr(dU), ptr(dV), k, 0.05, 0.01, 0.3, ptr(stats), current_stream()), "replay")
    assert tuple(stats.cpu().tolist()) == (c, sk)
    assert rel_err(dU.cpu().numpy(), Uo) < 1e-5 and rel_err(dV.cpu().numpy(), Vo) < 1e-5


@pytest.mark.parametrize("k,use_bias", [(8, True), (10, False), (64, True)])
def test_sbpr_replay_kernel_matches_oracle_on_a_stream(k, use_bias):
    import torch
    from cornac_b200 import _lib
    from cornac_b200._lib import check, current_stream, ptr
    from oracle import oracle as O
    L = _lib.load()
    n_users, n_items, nnz = 200, 150, 4000
    indptr, indices = synth_csr(n_users, n_items, nnz, seed=7)
    nnz = len(indices)
    s_ids, s_cnts, s_ptr = _social_lists(indptr, indices, seed=8)
    rng = np.random.RandomState(9)
    n = 5000
    ii = rng.randint(nnz, size=n).astype(np.int64)
    coo = O.coo_rows(indptr)
    u = coo[ii]
    ns = s_ptr[u + 1] - s_ptr[u]
    kidx = (s_ptr[u].astype(np.int64) + np.floor(rng.rand(n) * ns).astype(np.int64))
    j = rng.randint(n_items, size=n).astype(np.int32)
    assert (ns == 0).any() and (ns > 0).any() and kidx.max() < len(s_ids)
    U0 = (rng.rand(n_users, k) - 0.5).astype(np.float32)
    V0 = (rng.rand(n_items, k) - 0.5).astype(np.float32)
    B0 = (rng.rand(n_items) - 0.5).astype(np.float32)
    Uo, Vo, Bo = U0.copy(), V0.copy(), B0.copy()
    sk = O.sbpr_replay(ii, j, kidx, indptr, indices, s_ids, s_cnts, s_ptr, Uo, Vo, Bo, 0.05, 0.02, 0.03, 0.04, use_bias)
    dU, dV, dB = _dev(U0, torch.float32), _dev(V0, torch.float32), _dev(B0, torch.float32)
    stats = torch.zeros(2, dtype=torch.int64, device="cuda")
    t = [_dev(ii, torch.int64), _dev(j, torch.int32), _dev(kidx, torch.int64), _dev(indptr, torch.int32), _dev(indices, torch.int32),
         _dev(coo, torch.int32), _dev(s_ptr, torch.int32), _dev(s_ids, torch.int32), _dev(s_cnts, torch.int32)]
    check(L.b200_sbpr_epoch_replay(ptr(t[0]), ptr(t[1]), ptr(t[2]), n, ptr(t[3]), ptr(t[4]), ptr(t[5]), ptr(t[6]), ptr(t[7]), ptr(t[8]),
                                   len(s_ids), ptr(dU), ptr(dV), ptr(dB), k, 0.05, 0.02, 0.03, 0.04, int(use_bias), ptr(stats),
                                   current_stream()), "replay")
    assert int(stats[1].item()) == sk
    assert rel_err(dU.cpu().numpy(), Uo) < 1e-5 and rel_err(dV.cpu().numpy(), Vo) < 1e-5 and rel_err(dB.cpu().numpy(), Bo) < 1e-5


def _identity_dataset(g):
    from cornac.data import Dataset
    n_users, n_items = int(g["total_users"]), int(g["total_items"])
    rows = np.repeat(np.arange(len(g["indptr"]) - 1), np.diff(g["indptr"]))
    uid_map = OrderedDict((str(u), u) for u in range(n_users))
    iid_map = OrderedDict((str(i), i) for i in range(n_items))
    triples = [(str(u), str(i), float(r)) for u, i, r in zip(rows, g["indices"], g["data"])]
    return Dataset.build(triples, global_uid_map=uid_map, global_iid_map=iid_map)


@needs_cornac
def test_vebpr_plugin_reproduces_seeded_reference():
    """cornac_b200.VEBPR(seed) on a PurchaseViewDataset == the compiled reference's VEBPR(seed) (fixture vebpr_mid_k16)."""
    import scipy.sparse as sp
    from cornac.data import PurchaseViewDataset
    from cornac_b200 import VEBPR
    g = golden("vebpr_mid_k16")
    ds = _identity_dataset(g)
    W = sp.csr_matrix((np.ones(len(g["view_indices"]), np.float32), g["view_indices"], g["view_indptr"]), shape=ds.matrix.shape)
    pv = PurchaseViewDataset(ds, W)
    assert np.array_equal(pv.matrix.indices, g["indices"]) and np.array_equal(pv.view_matrix.indices, g["view_indices"])
    m = VEBPR(k=int(g["k"]), max_iter=int(g["max_iter"]), learning_rate=float(g["lr"]), lambda_reg=float(g["reg"]),
              alpha=float(g["alpha"]), seed=int(g["seed"])).fit(pv)
    assert m.u_factor.dtype == np.float32 and m.u_factor.shape == g["U"].shape
    assert rel_err(m.u_factor, g["U"]) < TOL and rel_err(m.i_factor, g["V"]) < TOL
    s = m.score(3)
    assert s.dtype == np.float32 and np.allclose(s, g["V"] @ g["U"][3], rtol=1e-4, atol=1e-6)
    ranked, scores = m.rank(3, k=10)
    assert np.array_equal(np.sort(ranked), np.arange(ds.num_items)) and np.array_equal(scores, s[: ds.num_items])
    with pytest.raises(ValueError):
        VEBPR(k=8, max_iter=1).fit(ds)                      # not a PurchaseViewDataset (recom_vebpr.pyx:176-181)
    c = m.clone()
    assert c.alpha == m.alpha and c.k == m.k and c.seed == m.seed


@needs_cornac
def test_sbpr_plugin_reproduces_seeded_reference():
    """cornac_b200.SBPR(seed).fit == the compiled reference's _fit_sgd driven as its fit() names it (fixture sbpr_mid_k16);
    the social-item lists equal the reference's _prepare_social_data."""
    from cornac.data import GraphModality
    from cornac_b200 import SBPR
    g = golden("sbpr_mid_k16")
    ds = _identity_dataset(g)
    gp, gi = g["graph_indptr"], g["graph_indices"]
    edges = [(str(a), str(int(b)), 1.0) for a in range(len(gp) - 1) for b in gi[gp[a]:gp[a + 1]]]
    gm = GraphModality(data=edges)
    gm.build(id_map=ds.uid_map)
    ds.add_modalities(user_graph=gm)
    m = SBPR(k=int(g["k"]), max_iter=int(g["max_iter"]), learning_rate=float(g["lr"]), lambda_u=float(g["lbd_u"]),
             lambda_v=float(g["lbd_v"]), lambda_b=float(g["lbd_b"]), use_bias=bool(g["use_bias"]), seed=int(g["seed"]))
    ids, cnts, ptr_ = m._prepare_social_data(ds)
    assert np.array_equal(ids, g["social_item_ids"]) and np.array_equal(cnts, g["social_item_counts"])
    assert np.array_equal(ptr_, g["social_indptr"])
    m.fit(ds)
    assert [s[1] for s in m.epoch_stats] == g["skipped"].tolist()
    assert rel_err(m.u_factors, g["U"]) < TOL and rel_err(m.i_factors, g["V"]) < TOL and rel_err(m.i_biases, g["B"]) < TOL
    s = m.score(5)
    assert np.allclose(s, g["B"] + g["V"] @ g["U"][5], rtol=1e-4, atol=1e-6)


def _planted(n_users, n_items, nnz, k_true, seed):
    """Interactions drawn from a planted low-rank preference: a trained model must rank positives above random items."""
    rng = np.random.RandomState(seed)
    P, Q = rng.normal(0, 1, (n_users, k_true)), rng.normal(0, 1, (n_items, k_true))
    u = rng.randint(n_users, size=nnz * 4)
    i = rng.randint(n_items, size=nnz * 4)
    keep = np.einsum("nk,nk->n", P[u], Q[i]) > 1.0
    key = np.unique(u[keep].astype(np.int64) * n_items + i[keep])[:nnz]
    u, i = key // n_items, key % n_items
    indptr = np.zeros(n_users + 1, np.int64)
    np.add.at(indptr, u + 1, 1)
    return np.cumsum(indptr).astype(np.int32), i.astype(np.int32)


def _pair_accuracy(U, V, B, indptr, indices, seed):
    rng = np.random.RandomState(seed)
    rows = np.repeat(np.arange(len(indptr) - 1), np.diff(indptr))
    pick = rng.randint(len(indices), size=20000)
    u, i, j = rows[pick], indices[pick], rng.randint(V.shape[0], size=len(pick))
    b = B if B is not None else np.zeros(V.shape[0], np.float32)
    return float(np.mean(np.einsum("nk,nk->n", U[u], V[i] - V[j]) + b[i] - b[j] > 0))


@pytest.mark.parametrize("k", [16, 30, 128])
def test_vebpr_hogwild_learns_and_keeps_the_sample_law(k):
    from cornac_b200 import engine
    n_users, n_items = 3000, 2000
    indptr, indices = _planted(n_users, n_items, 120000, 6, seed=1)
    vptr, vidx = _view_csr(indptr, indices, n_items, seed=2)
    rng = np.random.RandomState(3)
    U = ((rng.rand(n_users, k) - 0.5) / k).astype(np.float32)
    V = ((rng.rand(n_items, k) - 0.5) / k).astype(np.float32)
    U0, V0 = U.copy(), V.copy()
    # lr = 0: nothing moves; the skip rate is the law's: P(j in purchases(u) or views(u)) over the sampled users
    hist, _ = engine.tri_train_host("vebpr", indptr, indices, (vptr, vidx), n_items, U, V, None, dict(lr=0.0, reg=0.01, alpha=0.5), 2, key=5)
    assert np.array_equal(U, U0) and np.array_equal(V, V0)
    deg = np.diff(indptr).astype(np.float64)
    nv = np.diff(vptr).astype(np.float64)
    expect = float(np.sum(deg * (deg + nv)) / n_items / len(indices))       # users drawn in proportion to their degree
    for c, sk in hist:
        assert abs(sk / len(indices) - expect) < 0.15 * expect + 2e-3
    before = _pair_accuracy(U, V, None, indptr, indices, 4)
    hist, _ = engine.tri_train_host("vebpr", indptr, indices, (vptr, vidx), n_items, U, V, None, dict(lr=0.05, reg=0.001, alpha=0.5), 30, key=5)
    after = _pair_accuracy(U, V, None, indptr, indices, 4)
    assert before < 0.6 and after > 0.7, (before, after)      # Hogwild: measured 0.83-0.95, racy by design
    assert np.isfinite(U).all() and np.isfinite(V).all()


@pytest.mark.parametrize("k", [16, 30, 128])
def test_sbpr_hogwild_learns_and_keeps_the_sample_law(k):
    from cornac_b200 import engine
    n_users, n_items = 3000, 2000
    indptr, indices = _planted(n_users, n_items, 120000, 6, seed=11)
    s_ids, s_cnts, s_ptr = _social_lists(indptr, indices, seed=12)
    rng = np.random.RandomState(13)
    U = ((rng.rand(n_users, k) - 0.5) / k).astype(np.float32)
    V = ((rng.rand(n_items, k) - 0.5) / k).astype(np.float32)
    B = np.zeros(n_items, np.float32)
    U0, V0 = U.copy(), V.copy()
    hy = dict(lr=0.0, lambda_u=0.01, lambda_v=0.01, lambda_b=0.01, use_bias=True)
    hist, _ = engine.tri_train_host("sbpr", indptr, indices, (s_ptr, s_ids, s_cnts), n_items, U, V, B, hy, 2, key=7)
    assert np.array_equal(U, U0) and np.array_equal(V, V0) and not B.any()
    deg = np.diff(indptr).astype(np.float64)
    expect = float(np.sum(deg * deg) / n_items / len(indices))               # + the rare j == k_id
    for c, sk in hist:
        assert abs(sk / len(indices) - expect) < 0.15 * expect + 2e-3
    before = _pair_accuracy(U, V, B, indptr, indices, 14)
    hy["lr"] = 0.05
    hy.update(lambda_u=0.001, lambda_v=0.001, lambda_b=0.001)
    engine.tri_train_host("sbpr", indptr, indices, (s_ptr, s_ids, s_cnts), n_items, U, V, B, hy, 30, key=7)
    after = _pair_accuracy(U, V, B, indptr, indices, 14)
    assert before < 0.6 and after > 0.7, (before, after)      # Hogwild: measured 0.83-0.95, racy by design
    assert np.isfinite(U).all() and np.isfinite(V).all() and np.isfinite(B).all()


@needs_cornac
def test_unsupported_factor_width_is_refused():
    from cornac_b200.recom_bprx import check_tri_factor_width
    check_tri_factor_width(512), check_tri_factor_width(126)
    for k in (0, 130, 516, 1024):
        with pytest.raises(ValueError):
            check_tri_factor_width(k)
