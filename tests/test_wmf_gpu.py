"""WMF on the GPU (b200_wmf_step, cornac_b200.WMF) against the CPU restatement of the reference's TensorFlow-1 graph
(oracle/wmf_oracle.py; cornac/models/wmf/wmf.py:34-55, recom_wmf.py:152-240).  TensorFlow is not installed, so the
reference itself cannot produce golden vectors: PARITY UNPINNED (the restatement is checked for internal consistency in
tests/test_oracle_golden.py).  Tolerance: 1e-4 norm-wise on the trained factors (f32 graph; summation order of the
dense products differs between BLAS and the kernel).  GPU only."""
from collections import OrderedDict

import numpy as np
import pytest
import scipy.sparse as sp

from conftest import needs_cornac, rel_err
from oracle import wmf_oracle as W

pytestmark = pytest.mark.gpu
TOL = 1e-4


def _ratings(n_users, n_items, density, seed):
    rng = np.random.RandomState(seed)
    mask = rng.rand(n_users, n_items) < density
    R = np.where(mask, rng.randint(1, 6, size=(n_users, n_items)), 0).astype(np.float32)
    R[rng.randint(n_users), :] = 0                                  # a user without ratings
    R[:, rng.randint(n_items)] = 0                                  # an item without ratings
    return sp.csc_matrix(R)


@pytest.mark.parametrize("n_users,n_items,k,bs,a,b", [(300, 200, 16, 32, 1.0, 0.01), (1000, 333, 50, 128, 2.0, 0.05),
                                                       (70, 90, 200, 128, 1.0, 0.01), (513, 64, 33, 7, 1.0, 0.0)])
def test_wmf_steps_match_the_restated_graph(n_users, n_items, k, bs, a, b):
    """every step of 3 epochs: same batch loss, and the same U, V at the end (dense Adam on U, non-lazy sparse Adam on V,
    clipping, confidence weights, users spanning several CTA tiles, ragged last batch, empty rows and columns)"""
    from cornac_b200 import engine
    R = _ratings(n_users, n_items, 0.08, seed=k + bs)
    rng = np.random.RandomState(1)
    U0, V0 = W.xavier_uniform((n_users, k), rng), W.xavier_uniform((n_items, k), rng)
    lam_u, lam_v, lr = 0.01, 0.02, 0.003
    order = np.random.RandomState(2)
    epochs = []
    for _ in range(3):
        perm = order.permutation(n_items)
        epochs.append([perm[i:i + bs] for i in range(0, n_items, bs)])
    Uo, Vo = U0.copy(), V0.copy()
    opt, sU, sV = W.Adam(lr), W.AdamState(Uo.shape), W.AdamState(Vo.shape)
    tr = engine.WmfTrainer(R, U0.copy(), V0.copy(), a, b, lam_u, lam_v, lr)
    for batches in epochs:
        for ids in batches:
            R_b, C_b = W.batch_inputs(R, ids, a, b)
            want = float(W.train_step(Uo, Vo, ids, R_b, C_b, lam_u, lam_v, opt, sU, sV))
            got = tr.step(ids)
            assert abs(got - want) <= 2e-4 * max(1.0, abs(want)), (got, want)
    Ug, Vg = tr.download()
    assert rel_err(Ug, Uo) < TOL and rel_err(Vg, Vo) < TOL, (rel_err(Ug, Uo), rel_err(Vg, Vo))
    assert rel_err(tr.mU.cpu().numpy(), sU.m) < 1e-3 and rel_err(tr.vV.cpu().numpy(), sV.v) < 1e-3
    assert int((tr.slot_of != -1).sum().item()) == 0                # scratch map restored after every step


def test_wmf_gradient_clipping_and_large_batch():
    """ratings far from the predictions drive the raw gradients beyond +-5: the clipped step equals the restated one"""
    from cornac_b200 import engine
    R = _ratings(200, 150, 0.3, seed=5) * 40.0
    rng = np.random.RandomState(3)
    U0, V0 = W.xavier_uniform((200, 24), rng) * 5, W.xavier_uniform((150, 24), rng) * 5
    Uo, Vo = U0.copy(), V0.copy()
    opt, sU, sV = W.Adam(0.01), W.AdamState(Uo.shape), W.AdamState(Vo.shape)
    tr = engine.WmfTrainer(R, U0.copy(), V0.copy(), 1.0, 0.01, 0.01, 0.01, 0.01)
    for ids in (np.arange(0, 150, 2), np.arange(1, 150, 2), np.arange(150)):
        R_b, C_b = W.batch_inputs(R, ids, 1.0, 0.01)
        _, gU, _ = W.loss_and_grads(Uo, Vo[ids], R_b, C_b, 0.01, 0.01)
        assert np.abs(gU).max() > 5.0                               # the clip is really exercised
        W.train_step(Uo, Vo, ids, R_b, C_b, 0.01, 0.01, opt, sU, sV)
        tr.step(ids, want_loss=False)
    Ug, Vg = tr.download()
    assert rel_err(Ug, Uo) < TOL and rel_err(Vg, Vo) < TOL


@needs_cornac
def test_wmf_plugin_matches_the_restated_fit_and_serves_scores():
    """cornac_b200.WMF through a cornac Dataset: same constructor arguments as cornac.models.WMF, the reference's own
    item mini-batch schedule (train_set.item_iter(batch_size, shuffle=True)), U / V equal to the restated _fit_cf;
    score / rank / ScoreException like recom_wmf.py:214-240."""
    from cornac.data import Dataset
    from cornac.exception import ScoreException
    from cornac_b200 import WMF
    rng = np.random.RandomState(7)
    n_users, n_items = 120, 80
    triples = [(str(u), str(i), float(rng.randint(1, 6))) for u in range(n_users) for i in range(n_items) if rng.rand() < 0.1]

    def build():
        return Dataset.build(triples, global_uid_map=OrderedDict(), global_iid_map=OrderedDict(), seed=5)
    ds, ds_twin = build(), build()
    m = WMF(k=20, max_iter=4, learning_rate=0.005, lambda_u=0.02, lambda_v=0.03, a=1.5, b=0.02, batch_size=32, seed=9,
            verbose=False).fit(ds)
    U0 = W.xavier_uniform((ds.num_users, 20), np.random.RandomState(9))
    # the reference draws U then V from ONE generator (recom_wmf.py:121-126)
    g = np.random.RandomState(9)
    U0, V0 = W.xavier_uniform((ds.num_users, 20), g), W.xavier_uniform((ds.num_items, 20), g)
    W.fit(ds_twin.csc_matrix, U0, V0, lambda: list(ds_twin.item_iter(32, shuffle=True)), a=1.5, b=0.02, lambda_u=0.02,
          lambda_v=0.03, lr=0.005, max_iter=4)
    assert rel_err(m.U, U0) < TOL and rel_err(m.V, V0) < TOL, (rel_err(m.U, U0), rel_err(m.V, V0))
    # scores: V . U[u] (recom_wmf.py:237-240), full vector from the device kernel, single item on the host
    s = m.score(3)
    assert s.shape == (ds.num_items,) and np.allclose(s, m.V @ m.U[3], rtol=1e-5, atol=1e-6)
    assert abs(m.score(3, 4) - float(m.V[4] @ m.U[3])) < 1e-6
    with pytest.raises(ScoreException):
        m.score(ds.num_users + 5)
    ranked, scores = m.rank(3, k=10)
    assert len(ranked) == ds.num_items and np.array_equal(ranked[:10], np.lexsort((np.arange(ds.num_items), -s.astype(np.float64)))[:10])
    ids, sc = m.rank_batch(np.arange(10), 5)
    for u in range(10):
        su = m.score(u)
        assert np.array_equal(ids[u], np.lexsort((np.arange(ds.num_items), -su.astype(np.float64)))[:5])
    c = m.clone()
    assert (c.k, c.a, c.b, c.batch_size, c.seed) == (20, 1.5, 0.02, 32, 9)
