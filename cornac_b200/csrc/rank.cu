// b200_rank_topk: scores + exclusion + top-k for a batch of users.
//
// Replaces the per-user `score -> argpartition -> argsort` of Recommender.rank
// (reference: cornac/models/recommender.py:476-530, driven once per test user by
// ranking_eval, cornac/eval_methods/base_method.py:177-220) with batched device work.
//
// This translation unit holds the exact path: chunks of queries are scored with
// score_batch_kernel into the workspace and reduced with topk_rows_kernel.  The
// tensor-core candidate pass (rank_tc.cu) plugs in in front of it when shapes allow.
#include "common.cuh"

namespace b200 {
int rank_tc_supported(int64_t n_q, int64_t n_items, int k, int topk);
int64_t rank_tc_workspace_bytes(int64_t n_q, int64_t n_items, int k, int topk);
int rank_tc(const float* U, const int64_t* user_idx, int64_t n_q, const float* V, int64_t n_items, int k,
            const float* item_base, const float* user_off, const int64_t* excl_indptr, const int32_t* excl_indices,
            int topk, int32_t* out_ids, float* out_scores, void* workspace, int64_t workspace_bytes, const void* packed_items,
            cudaStream_t st);
int64_t rank_tc_items_bytes(int64_t n_items, int k);
int rank_tc_pack_items(const float* V, int64_t n_items, int k, const float* item_base, void* packed, int64_t packed_bytes,
                       cudaStream_t st);
}  // namespace b200

using namespace b200;

static int64_t exact_chunk_queries(int64_t n_q, int64_t n_items)
{
    // keep the score slab around 256 MB so it stays largely L2-resident between the passes
    int64_t c = (256ll << 20) / (n_items * 4 > 0 ? n_items * 4 : 1);
    if (c < 8) c = 8;
    if (c > n_q) c = n_q;
    return c;
}

extern "C" int64_t b200_rank_topk_workspace_bytes(int64_t n_q, int64_t n_items, int k, int topk)
{
    if (n_q <= 0 || n_items <= 0) return 0;
    if (rank_tc_supported(n_q, n_items, k, topk)) return rank_tc_workspace_bytes(n_q, n_items, k, topk);
    return exact_chunk_queries(n_q, n_items) * n_items * (int64_t)sizeof(float);
}

extern "C" int64_t b200_rank_items_bytes(int64_t n_items, int k)
{
    if (n_items <= 0 || !rank_tc_supported(1, n_items, k, 1)) return 0;       // shapes the tensor-core pass does not take
    return rank_tc_items_bytes(n_items, k);
}

extern "C" int b200_rank_pack_items(const float* V, int64_t n_items, int k, const float* item_base,
                                    void* packed, int64_t packed_bytes, void* stream)
{
    B200_REQUIRE(V && n_items >= 1 && k >= 1, "b200_rank_pack_items: bad argument");
    B200_REQUIRE(b200_rank_items_bytes(n_items, k) > 0, "b200_rank_pack_items: shape not taken by the tensor-core pass (k=%d, n_items=%lld)",
                 k, (long long)n_items);
    return rank_tc_pack_items(V, n_items, k, item_base, packed, packed_bytes, (cudaStream_t)stream);
}

extern "C" int b200_rank_topk(const float* U, const int64_t* user_idx, int64_t n_q,
                              const float* V, int64_t n_items, int k,
                              const float* item_base, const float* user_off,
                              const int64_t* excl_indptr, const int32_t* excl_indices,
                              int topk, int32_t* out_ids, float* out_scores,
                              void* workspace, int64_t workspace_bytes, void* stream)
{
    return b200_rank_topk_packed(U, user_idx, n_q, V, n_items, k, item_base, user_off, excl_indptr, excl_indices, topk,
                                 out_ids, out_scores, nullptr, workspace, workspace_bytes, stream);
}

extern "C" int b200_rank_topk_packed(const float* U, const int64_t* user_idx, int64_t n_q,
                                     const float* V, int64_t n_items, int k,
                                     const float* item_base, const float* user_off,
                                     const int64_t* excl_indptr, const int32_t* excl_indices,
                                     int topk, int32_t* out_ids, float* out_scores,
                                     const void* packed_items, void* workspace, int64_t workspace_bytes, void* stream)
{
    B200_REQUIRE(U && V && out_ids && out_scores, "b200_rank_topk: null pointer argument");
    B200_REQUIRE(n_q >= 0 && n_items >= 1 && k >= 1 && topk >= 1, "b200_rank_topk: bad sizes");
    if (n_q == 0) return B200_OK;
    const int64_t need = b200_rank_topk_workspace_bytes(n_q, n_items, k, topk);
    B200_REQUIRE(workspace && workspace_bytes >= need, "b200_rank_topk: workspace too small (%lld < %lld bytes)",
                 (long long)workspace_bytes, (long long)need);
    cudaStream_t st = (cudaStream_t)stream;
    if (rank_tc_supported(n_q, n_items, k, topk))
        return rank_tc(U, user_idx, n_q, V, n_items, k, item_base, user_off, excl_indptr, excl_indices, topk, out_ids,
                       out_scores, workspace, workspace_bytes, packed_items, st);
    const int64_t chunk = exact_chunk_queries(n_q, n_items);
    float* slab = static_cast<float*>(workspace);
    for (int64_t q0 = 0; q0 < n_q; q0 += chunk) {
        const int64_t nq = (n_q - q0 < chunk) ? n_q - q0 : chunk;
        // rows q0.. of U when no index list is given
        const float* Uq = user_idx ? U : U + (size_t)q0 * k;
        int rc = b200_score_batch(Uq, user_idx ? user_idx + q0 : nullptr, nq, V, n_items, k, item_base,
                                  user_off ? user_off + q0 : nullptr, slab, stream);
        if (rc) return rc;
        rc = b200_topk_rows(slab, nq, n_items, excl_indptr ? excl_indptr + q0 : nullptr, excl_indices, topk,
                            out_ids + (size_t)q0 * topk, out_scores + (size_t)q0 * topk, stream);
        if (rc) return rc;
    }
    (void)st;
    return B200_OK;
}
