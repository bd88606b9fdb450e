"""Batched, device-side replacement of the per-user loop of cornac's `ranking_eval`.

Reference: `cornac/eval_methods/base_method.py:108-226` iterates over the test users in Python and,
for each, builds two dense item masks, calls `model.rank(user, item_indices, k)` and feeds the ranked
list to every metric (`cornac/metrics/ranking.py`).  Here the candidate sets are expressed as per-user
exclusion lists (train / validation positives that are not test positives), ALL users are ranked by
`rank_batch_device` (one fused tensor-core kernel per chunk of users) and the per-user metric values are
reduced on the GPU from the device-resident top-k ids by `b200_topk_metrics` (csrc/eval.cu): only
n_metrics doubles per user come back to the host.  SURVEY.md section 8, row (f)2.

Supported metrics: NDCG@k, NCRR@k, Precision@k, Recall@k, FMeasure@k, HitRatio@k with k > 0 -- the ones that
only look at `pd_rank[:k]`.  Anything else (AUC, MAP, MRR, k = -1) needs the full score vector per user and is
delegated to the reference implementation unchanged.  Results are the same numbers the reference loop
produces with the same model (same ids: both use the total order score desc, item id asc).
"""
import numpy as np
import scipy.sparse as sp
import torch

from cornac.eval_methods.base_method import ranking_eval as _reference_ranking_eval
from cornac.metrics import FMeasure, HitRatio, NCRR, NDCG, Precision, Recall

from . import _lib, engine

# most-derived classes first is not needed: none of these derives from another one in the list
_KIND = ((NDCG, _lib.METRIC_NDCG), (NCRR, _lib.METRIC_NCRR), (Precision, _lib.METRIC_PRECISION),
         (Recall, _lib.METRIC_RECALL), (FMeasure, _lib.METRIC_FMEASURE), (HitRatio, _lib.METRIC_HIT))
_TOPK_ONLY = tuple(c for c, _ in _KIND)


def _kind(metric):
    for cls, kind in _KIND:
        if isinstance(metric, cls):
            return kind
    raise TypeError(type(metric).__name__)


def _positives(mat, threshold, n_rows, n_cols):
    """CSR 0/1 matrix of the entries with rating >= threshold, reshaped to [n_rows, n_cols]."""
    m = mat.tocsr()
    keep = m.data >= threshold
    rows = np.repeat(np.arange(m.shape[0]), np.diff(m.indptr))[keep]
    cols = m.indices[keep]
    ok = (rows < n_rows) & (cols < n_cols)
    return sp.csr_matrix((np.ones(int(ok.sum()), dtype=np.int8), (rows[ok], cols[ok])), shape=(n_rows, n_cols))


def ranking_eval(model, metrics, train_set, test_set, val_set=None, rating_threshold=1.0, exclude_unknowns=True,
                 verbose=False, batch_users=75776):
    """Same signature and return value as cornac.eval_methods.base_method.ranking_eval."""
    if len(metrics) == 0:
        return [], []
    supported = (hasattr(model, "rank_batch_device") and exclude_unknowns
                 and all(isinstance(m, _TOPK_ONLY) and 0 < m.k <= 4096 for m in metrics))      # b200_topk_rows: topk <= 4096
    if not supported:
        return _reference_ranking_eval(model, metrics, train_set, test_set, val_set=val_set,
                                       rating_threshold=rating_threshold, exclude_unknowns=exclude_unknowns,
                                       verbose=verbose)
    max_k = max(m.k for m in metrics)
    n_items = train_set.num_items                               # exclude_unknowns: candidates are the train items
    n_rows = max(test_set.csr_matrix.shape[0], train_set.csr_matrix.shape[0])
    test_pos = _positives(test_set.csr_matrix, rating_threshold, n_rows, n_items)
    seen = _positives(train_set.csr_matrix, rating_threshold, n_rows, n_items)
    if val_set is not None:
        seen = seen + _positives(val_set.csr_matrix, rating_threshold, n_rows, n_items)
    excl = (seen - seen.multiply(test_pos)).tocsr()             # train/val positives that are not test positives
    excl.eliminate_zeros()
    excl.sort_indices()
    test_pos.sort_indices()

    # same user order as the reference loop (`for user_idx in set(test_set.uir_tuple[0])`), same skip rule
    full_test = _positives(test_set.csr_matrix, rating_threshold, n_rows, test_set.csr_matrix.shape[1])
    users = np.fromiter((u for u in set(test_set.uir_tuple[0]) if full_test.indptr[u + 1] > full_test.indptr[u]),
                        dtype=np.int64)
    kinds = [_kind(m) for m in metrics]
    ks = [int(m.k) for m in metrics]
    pos_ptr = engine.to_device(test_pos.indptr.astype(np.int64), torch.int64)
    pos_idx = engine.to_device(test_pos.indices.astype(np.int32) if test_pos.nnz else np.zeros(1, np.int32),
                               torch.int32)
    per_metric = np.empty((len(metrics), len(users)), dtype=np.float64)
    for b0 in range(0, len(users), batch_users):
        ub = users[b0:b0 + batch_users]
        ids, _ = model.rank_batch_device(ub, max_k, exclude=excl, n_items=n_items)    # [n, max_k] int32 CUDA, -1 padded
        vals = engine.topk_metrics(ids, pos_ptr, pos_idx, kinds, ks, user_idx=engine.to_device(ub, torch.int64))
        per_metric[:, b0:b0 + len(ub)] = vals.cpu().numpy()
    user_results = [dict(zip(users.tolist(), vals.tolist())) for vals in per_metric]
    avg_results = [sum(r.values()) / len(r) for r in user_results]
    return avg_results, user_results
