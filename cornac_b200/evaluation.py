"""Batched, device-side replacement of the per-user loop of cornac's `ranking_eval`.

Reference: `cornac/eval_methods/base_method.py:108-226` iterates over the test users in Python and,
for each, builds two dense item masks, calls `model.rank(user, item_indices, k)` and feeds the ranked
list to every metric (`cornac/metrics/ranking.py`).  Here the candidate sets are expressed as per-user
exclusion lists (train / validation positives that are not test positives), ALL users are ranked by
`rank_batch` (one fused tensor-core kernel per chunk of users) and the top-k metrics are computed with
vectorised numpy on the returned ids.  SURVEY.md section 8, row (f)2.

Supported metrics: NDCG@k, Precision@k, Recall@k, FMeasure@k, HitRatio@k with k > 0 -- the ones that only
look at `pd_rank[:k]`.  Anything else (AUC, MAP, MRR, k = -1) needs the full score vector per user and is
delegated to the reference implementation unchanged.  Results are the same numbers the reference loop
produces with the same model (same ids: both use the total order score desc, item id asc).
"""
import numpy as np
import scipy.sparse as sp

from cornac.eval_methods.base_method import ranking_eval as _reference_ranking_eval
from cornac.metrics import FMeasure, HitRatio, NDCG, Precision, Recall

_TOPK_ONLY = (NDCG, Precision, Recall, FMeasure, HitRatio)


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
    supported = (hasattr(model, "rank_batch") and exclude_unknowns
                 and all(isinstance(m, _TOPK_ONLY) and m.k > 0 for m in metrics))
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
    n_pos = np.diff(test_pos.indptr)[users].astype(np.float64)

    per_metric = [np.empty(len(users), dtype=np.float64) for _ in metrics]
    inv_disc = 1.0 / np.log2(np.arange(max_k) + 2.0)
    cum_idcg = np.concatenate([[0.0], np.cumsum(inv_disc)])
    # sorted keys (user, item) of the test positives for the membership lookup
    pos_keys = np.repeat(np.arange(n_rows, dtype=np.int64), np.diff(test_pos.indptr)) * n_items + test_pos.indices
    for b0 in range(0, len(users), batch_users):
        ub = users[b0:b0 + batch_users]
        ids, _ = model.rank_batch(ub, max_k, exclude=excl)       # [n, max_k] int32, -1 padded
        pred_keys = ub[:, None] * n_items + ids.astype(np.int64)
        hits = (np.isin(pred_keys, pos_keys, assume_unique=False) & (ids >= 0))
        npos_b = n_pos[b0:b0 + batch_users]
        for mi, m in enumerate(metrics):
            h = hits[:, : m.k]
            tp = h.sum(axis=1).astype(np.float64)
            if isinstance(m, NDCG):
                dcg = (h * inv_disc[: m.k]).sum(axis=1)
                idcg = cum_idcg[np.minimum(npos_b, m.k).astype(np.int64)]
                val = dcg / idcg
            elif isinstance(m, HitRatio):
                val = (tp > 0).astype(np.float64)
            else:
                prec = tp / m.k
                rec = tp / npos_b
                if isinstance(m, Precision):
                    val = prec
                elif isinstance(m, Recall):
                    val = rec
                else:                                              # FMeasure (ranking.py: 2PR/(P+R), 0 when P+R == 0)
                    den = prec + rec
                    val = np.where(den > 0, 2 * prec * rec / np.where(den > 0, den, 1.0), 0.0)
            per_metric[mi][b0:b0 + len(ub)] = val
    user_results = [dict(zip(users.tolist(), vals.tolist())) for vals in per_metric]
    avg_results = [sum(r.values()) / len(r) for r in user_results]
    return avg_results, user_results
