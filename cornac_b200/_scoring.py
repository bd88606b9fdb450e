"""Device-resident scoring/ranking shared by the BPR and MF plug-ins.

Implements the `score()` / `rank()` half of the cornac.models.Recommender contract
(reference: cornac/models/recommender.py:423-441, 476-530) on the GPU, plus the batched
`rank_batch()` that the reference lacks (it ranks one user per Python call).
"""
import numpy as np
import torch

from . import engine
from ._lib import B200Error


class DeviceScoringMixin:
    """Expects the subclass to provide `_b200_host_params()` returning
    (U, V, item_base, user_off_vector_or_None, n_score_items) as numpy arrays."""

    _B200_IGNORED = ("_b200_dev",)

    def _b200_register_ignored(self):
        for a in self._B200_IGNORED:
            if a not in self.ignored_attrs:
                self.ignored_attrs.append(a)
        self._b200_dev = None

    # ---- device cache --------------------------------------------------------------
    def _b200_invalidate(self):
        self._b200_dev = None

    def _b200_device(self):
        dev = getattr(self, "_b200_dev", None)
        if dev is None:
            engine.require_cuda()
            U, V, item_base, user_off, n_items = self._b200_host_params()
            dev = dict(
                U=engine.to_device(U, torch.float32),
                V=engine.to_device(V, torch.float32),
                item_base=None if item_base is None else engine.to_device(item_base, torch.float32),
                user_off=None if user_off is None else engine.to_device(user_off, torch.float32),
                n_items=int(n_items),
            )
            self._b200_dev = dev
        return dev

    def _b200_adopt_device(self, U, V, item_base, user_off, n_items):
        """Keep the freshly trained device tensors as the scoring cache (no re-upload)."""
        self._b200_dev = dict(U=U, V=V, item_base=item_base, user_off=user_off, n_items=int(n_items))

    def _b200_packed_items(self, n_rank):
        """fp16 tile images of the item side for the fused rank, built once per (trained model, candidate count) and kept
        with the device cache: V and the item base are constant until the next fit() / parameter change, which drops
        the whole cache (_b200_invalidate)."""
        d = self._b200_device()
        cache = d.setdefault("packed", {})
        if n_rank not in cache:
            cache.clear()                                   # one candidate count at a time (288 MB at 1 M items)
            cache[n_rank] = engine.rank_pack_items(d["V"], d["item_base"], n_rank)
        return cache[n_rank]

    # ---- scores --------------------------------------------------------------------
    @staticmethod
    def _b200_check_users(user_indices, n_rows):
        """The kernels gather U rows without a bounds check: an index outside [0, n_rows) raises here, like the
        reference's numpy indexing does (IndexError), instead of reading foreign device memory."""
        user_indices = np.asarray(user_indices, dtype=np.int64)
        if user_indices.size and (int(user_indices.min()) < 0 or int(user_indices.max()) >= int(n_rows)):
            bad = user_indices[(user_indices < 0) | (user_indices >= n_rows)]
            raise IndexError("user index %d is out of bounds for the %d user rows of the model" % (int(bad[0]), int(n_rows)))
        return user_indices

    def _b200_scores_dev(self, user_indices):
        """[n_q, n_items] device scores for known users."""
        d = self._b200_device()
        user_indices = self._b200_check_users(user_indices, d["U"].shape[0])
        uidx = torch.as_tensor(user_indices).cuda()
        uoff = None if d["user_off"] is None else d["user_off"][uidx].contiguous()
        return engine.score_batch(d["U"], d["V"], user_idx=uidx, item_base=d["item_base"], user_off=uoff,
                                  n_items=d["n_items"])

    # ---- batched rank (the throughput path) ----------------------------------------
    def rank_batch(self, user_indices, k, exclude=None):
        """Top-k item ids and scores for many users at once.

        user_indices : int array [n_q]
        exclude      : optional scipy CSR matrix (rows = user index) whose stored columns
                       are removed from each user's candidates (e.g. train_set.csr_matrix)
        Returns (ids int32 [n_q, k] (-1 padded), scores float32 [n_q, k]) as numpy arrays,
        ordered by (score desc, item id asc).
        """
        d = self._b200_device()
        user_indices = self._b200_check_users(user_indices, d["U"].shape[0])
        ex_ptr, ex_idx = self._b200_exclusion_rows(user_indices, exclude)
        if d["user_off"] is None and d["n_items"] == d["V"].shape[0]:
            return engine.rank_topk_host(d["U"], d["V"], int(k), user_indices, item_base=d["item_base"],
                                         excl_indptr=ex_ptr, excl_indices=ex_idx, packed_items=self._b200_packed_items(d["n_items"]))
        ids, sc = self.rank_batch_device(user_indices, k, exclude=exclude, _rows=(ex_ptr, ex_idx))
        return ids.cpu().numpy(), sc.cpu().numpy()

    @staticmethod
    def _b200_exclusion_rows(user_indices, exclude):
        if exclude is None:
            return None, None
        n_q = len(user_indices)
        sub = exclude[user_indices] if n_q != exclude.shape[0] or not np.array_equal(
            user_indices, np.arange(exclude.shape[0])) else exclude
        sub = sub.tocsr()
        sub.sort_indices()
        return sub.indptr.astype(np.int64), sub.indices.astype(np.int32)

    def rank_batch_device(self, user_indices, k, exclude=None, _rows=None, n_items=None):
        """`rank_batch` leaving the result on the GPU: (ids int32 [n_q, k], scores f32 [n_q, k]) CUDA tensors
        (what the device-side metric reduction of cornac_b200.evaluation consumes).  `n_items` restricts the candidates
        to the first n_items item rows (ranking_eval with exclude_unknowns: only the train items, base_method.py:200-202)."""
        d = self._b200_device()
        user_indices = self._b200_check_users(user_indices, d["U"].shape[0])
        ex_ptr, ex_idx = _rows if _rows is not None else self._b200_exclusion_rows(user_indices, exclude)
        uidx = engine.to_device(user_indices, torch.int64)
        uoff = None if d["user_off"] is None else d["user_off"][uidx].contiguous()
        ep = None if ex_ptr is None else engine.to_device(ex_ptr, torch.int64)
        ei = None if ex_ptr is None else (engine.to_device(ex_idx, torch.int32) if len(ex_idx) else
                                          torch.zeros(1, dtype=torch.int32, device="cuda"))
        n_rank = d["n_items"] if n_items is None else min(int(n_items), d["n_items"])
        return engine.rank_topk(d["U"], d["V"], int(k), user_idx=uidx, item_base=d["item_base"], user_off=uoff,
                                excl_indptr=ep, excl_indices=ei, n_items=n_rank, packed_items=self._b200_packed_items(n_rank))

    # ---- batched Recommender.recommend ---------------------------------------------
    def recommend_batch(self, batch_users, k=-1, remove_seen=False, train_set=None):
        """Top-k recommendations for many users in one fused kernel call, in ORIGINAL ids: the batched form of
        `Recommender.recommend` (cornac/models/recommender.py:532-580), with the signature of the reference's only batched
        precedent (`ANNMixin.recommend_batch`, cornac/models/ann/recom_ann_base.py:182-235).  Seen items are removed
        BEFORE the top-k (every list has k items, unlike the ANN post-filter).  Returns a list of lists of item ids."""
        user_idx = [self.uid_map.get(uid, -1) for uid in batch_users]
        if any(i == -1 for i in user_idx):
            raise ValueError(f"{batch_users} is unknown to the model.")
        if k < -1 or k > self.total_items:
            raise ValueError(f"k={k} is invalid, there are {self.total_users} users in total.")
        if remove_seen and train_set is None:
            raise ValueError("train_set must be provided to remove seen items.")
        if k == -1 or k > 4096 or any(not self.knows_user(u) for u in user_idx):
            # full rankings / unknown users: not the batched path
            return [self.recommend(uid, k=k, remove_seen=remove_seen, train_set=train_set) for uid in batch_users]
        exclude = None
        if remove_seen:
            exclude = train_set.csr_matrix
            n_rows = max(user_idx) + 1
            if exclude.shape[0] < n_rows:                 # users without a training row have nothing to remove
                import scipy.sparse as sp
                exclude = sp.vstack([exclude, sp.csr_matrix((n_rows - exclude.shape[0], exclude.shape[1]), dtype=exclude.dtype)]).tocsr()
        ids, _ = self.rank_batch(np.asarray(user_idx, dtype=np.int64), int(k), exclude=exclude)
        item_ids = self.item_ids
        return [[item_ids[i] for i in row if i >= 0] for row in ids]

    # ---- Recommender.rank ------------------------------------------------------------
    def _b200_rank(self, all_scores_dev, item_indices, k):
        """Reference semantics of Recommender.rank (recommender.py:513-530) given the
        device score vector [1, total]: returns (ranked_items, item_scores)."""
        total = all_scores_dev.shape[1]
        all_item_scores = all_scores_dev[0].cpu().numpy()
        if item_indices is None:
            item_indices = np.arange(self.num_items)
            n_cand_all = self.num_items == total
        else:
            item_indices = np.asarray(item_indices)
            n_cand_all = False
        item_scores = all_item_scores[item_indices]
        n_cand = len(item_indices)
        if k == -1 or k >= n_cand or k > 4096:
            # full ordering requested: not the hot path (MRR-style metrics); host sort with the
            # same total order (score desc, id asc)
            order = np.lexsort((item_indices, -item_scores.astype(np.float64)))
            return item_indices[order], item_scores
        if n_cand_all:
            ex_ptr = ex_idx = None
        else:
            mask = np.ones(total, dtype=bool)
            mask[item_indices] = False
            excl = np.flatnonzero(mask).astype(np.int32)
            ex_ptr = engine.to_device(np.array([0, len(excl)], dtype=np.int64), torch.int64, pinned=False)
            ex_idx = engine.to_device(excl if len(excl) else np.zeros(1, np.int32), torch.int32, pinned=False)
        ids, _ = engine.topk_rows(all_scores_dev, int(k), ex_ptr, ex_idx)
        top = ids[0].cpu().numpy().astype(item_indices.dtype)
        in_top = np.zeros(total, dtype=bool)
        in_top[top] = True
        rest = item_indices[~in_top[item_indices]]
        return np.concatenate([top, rest]), item_scores
