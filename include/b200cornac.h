/*
 * b200cornac.h -- C ABI of libb200cornac.so: the B200 (sm_100a) implementation of
 * Cornac's embedding train-and-rank hot path (BPR / MF SGD, per-user score + rank).
 *
 * The reference has no C ABI for this path: its kernels are Cython `def`/`cpdef`
 * functions taking typed memoryviews (paths relative to the reference root):
 *     BPR._fit_sgd             cornac/models/bpr/recom_bpr.pyx:208-269
 *     RNGVector                cornac/models/bpr/recom_bpr.pyx:54-62 (+ recom_bpr.pxd:26-41)
 *     backend_cpu.fit_sgd      cornac/models/mf/backend_cpu.pyx:35-97
 *     fast_dot                 cornac/utils/fast_dot.pyx:40-43
 *     Recommender.rank         cornac/models/recommender.py:476-530
 * Each entry point below names the reference interface it replaces.  The Python
 * plug-in classes in cornac_b200/ (same constructor arguments as the reference's
 * BPR / MF) call these through ctypes; see INTEGRATION.md for the binding.
 *
 * Conventions
 *  - plain C types only; every pointer documented "device" is a CUDA device pointer
 *    (row-major, contiguous), every pointer documented "host" is ordinary host memory;
 *  - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream);
 *  - calls are asynchronous on `stream` unless stated otherwise and never allocate
 *    device memory; scratch comes in through explicit workspace arguments;
 *  - return value: 0 = ok, otherwise a B200_ERR_* code; b200_last_error() returns the
 *    message of the calling thread's last failure;
 *  - there is NO CPU fallback anywhere in this library.
 */
#ifndef B200CORNAC_H_
#define B200CORNAC_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B200_API __attribute__((visibility("default")))

#define B200_OK 0
#define B200_ERR_INVALID 1   /* bad argument (message says which) */
#define B200_ERR_CUDA 2      /* a CUDA runtime call or kernel launch failed */
#define B200_ERR_UNSUPPORTED 3

/* flags for the SGD epochs */
#define B200_SGD_ATOMIC 1u   /* scatter with red.global.add.f32 (no lost updates) instead of plain stores */
#define B200_SGD_EXACT_EXP 2u /* z = 1/(1+exp(double)) like the reference instead of the fast f32 path */
#define B200_BPR_NEG_WEIGHTED 8u /* WBPR (recom_wbpr.pyx:125-136): j = item of a uniformly drawn INTERACTION */
#define B200_BPR_LOSS_HINGE 16u  /* MMMF (cornac/models/mmmf/recom_mmmf.pyx:129-154): hinge loss, biases always trained */
#define B200_SGD_UNBOUNDED 4u /* do not cap the number of concurrently running samples (see b200_bpr_epoch) */
#define B200_BPR_BLOCKED 32u /* cache-blocked sample ORDER (same per-epoch law, see b200_bpr_block_plan): the epoch visits the
                              * interaction list window by window and the items block by block so that the rows in use stay in
                              * the L2; a no-op for matrices whose factors already fit (plan 1 x 1) */

B200_API const char* b200_last_error(void);
B200_API int b200_abi_version(void);
/* Number of CUDA kernels this library has launched in the calling process so far (all entry points, all
 * streams): what bench.py reports as `gpu_launches` around its timed region. */
B200_API int64_t b200_kernel_launches(void);
/* multiProcessorCount / compute capability of the current device (host call). */
B200_API int b200_device_info(int* sm_count, int* cc_major, int* cc_minor);

/* ------------------------------------------------------------------------------------
 * BPR, throughput mode.
 *
 * b200_bpr_prepare: analogue of BPR._prepare_data (recom_bpr.pyx:154-161).  Turns the CSR
 * training matrix into the two device structures the epoch kernel gathers from:
 *   pairs  device int32[nnz, 2]  (user, item) of every stored interaction, CSR order
 *                                (= the reference's user_ids[] and X.indices[] interleaved);
 *   table  device uint64[table_slots], table_slots = b200_bpr_table_slots(nnz): an
 *                                open-addressing hash set of (user << 32 | item) with 4-slot
 *                                buckets, answering has_non_zero(u, j) (recom_bpr.pyx:46-51)
 *                                with one 32-byte gather instead of a binary search.
 *   indptr device int32[n_users+1], indices device int32[nnz] (train_set.matrix).           */
B200_API int64_t b200_bpr_table_slots(int64_t nnz);
B200_API int b200_bpr_prepare(const int32_t* indptr, const int32_t* indices, int64_t n_users, int64_t nnz,
                              int32_t* pairs, uint64_t* table, int64_t table_slots, void* stream);

/* b200_bpr_epoch replaces one call of BPR._fit_sgd (recom_bpr.pyx:208-269) run Hogwild
 * over all cores: `n_samples` triplets, each drawing i_index uniformly from [0, nnz) and j
 * uniformly from [0, n_neg) ON DEVICE (Philox4x32-10 keyed by `seed`, counter =
 * (sample_base + s, epoch)), skipping (not redrawing) a sample when user u already has
 * item j (recom_bpr.pyx:241-243), otherwise applying the update of recom_bpr.pyx:249-267
 * to U[u], V[i], V[j], B[i], B[j].  Updates are scattered with red.global.add (B200_SGD_ATOMIC,
 * no lost updates; recommended and ~3x faster on B200) or plain stores (the reference's racy
 * Hogwild).  At most min(n_users, n_neg)/4 samples run concurrently so that small matrices
 * are not trained from hopelessly stale rows (B200_SGD_UNBOUNDED lifts the cap).
 *   U device f32[*, k], V device f32[*, k], B device f32[*]
 *   stats   device int64[2]: {correct, skipped} are ADDED to it (caller zeroes)           */
B200_API int b200_bpr_epoch(const int32_t* pairs, const uint64_t* table, int64_t table_slots,
                            int64_t nnz, int64_t n_users, int64_t n_neg, int64_t n_samples,
                            float* U, float* V, float* B, int k,
                            float lr, float reg, int use_bias,
                            uint64_t seed, uint64_t epoch, uint64_t sample_base,
                            unsigned flags, int64_t* stats, void* stream);

/* The plan of the cache-blocked order for a factor matrix pair: windows of the interaction list (user side) and blocks
 * of the items such that one window's user rows and one block's item rows stay L2-resident.  Sample s of an epoch
 * (s = sample_base + local index) belongs to run (s / ceil(nnz / (windows * blocks))) mod (windows * blocks); run
 * (w, b) draws i_index uniformly from the w-th window of [0, nnz) and j uniformly from item block (b + epoch) mod blocks:
 * every interaction is still drawn once per epoch in expectation and every negative is uniform over the items,
 * independently of the interaction (the law of recom_bpr.pyx:237-239); only the ORDER of the epoch's triplets changes. */
B200_API int b200_bpr_block_plan(int64_t n_users, int64_t n_neg, int k, uint32_t* n_windows, uint32_t* n_blocks);

/* The sample law of b200_bpr_epoch, evaluated on the HOST (no CUDA): writes, for
 * s = 0..n-1, the i_index and j_id that sample `sample_base + s` of `epoch` draws.  Lets a
 * caller replay / audit the exact stream the throughput kernel consumed.
 *   out_i_index host int64[n], out_j_id host int32[n]                                      */
B200_API int b200_bpr_draw_host(uint64_t seed, uint64_t epoch, uint64_t sample_base, int64_t n,
                                int64_t nnz, int64_t n_neg, int64_t* out_i_index, int32_t* out_j_id);
/* the same for an epoch run with B200_BPR_BLOCKED under the plan (n_windows, n_blocks) of b200_bpr_block_plan */
B200_API int b200_bpr_draw_host2(uint64_t seed, uint64_t epoch, uint64_t sample_base, int64_t n,
                                 int64_t nnz, int64_t n_neg, uint32_t n_windows, uint32_t n_blocks,
                                 int64_t* out_i_index, int32_t* out_j_id);

/* BPR, parity mode.  Applies an explicit sample stream (i_index[s], j_id[s]),
 * s = 0..n_samples-1, with the SAME RESULT AS APPLYING IT SEQUENTIALLY in stream order,
 * i.e. the seeded single-thread reference (recom_bpr.pyx:132-133).  The stream normally
 * comes from b200_mt_sampler_* below.  flags: B200_BPR_LOSS_HINGE selects the MMMF loop body.
 *   i_index device int64[n_samples], j_id device int32[n_samples]                          */
B200_API int b200_bpr_epoch_replay(const int64_t* i_index, const int32_t* j_id, int64_t n_samples,
                                   const int32_t* indptr, const int32_t* indices, const int32_t* coo_row,
                                   float* U, float* V, float* B, int k,
                                   float lr, float reg, int use_bias, unsigned flags,
                                   int64_t* stats, void* stream);

/* The same with the row counts of U and V given (n_users x k, n_items x k, B n_items): when the whole model fits the
 * shared memory of one SM (ML-100K sized problems) the epoch runs on an on-chip copy of the factors.  Same result. */
B200_API int b200_bpr_epoch_replay2(const int64_t* i_index, const int32_t* j_id, int64_t n_samples,
                                    const int32_t* indptr, const int32_t* indices, const int32_t* coo_row,
                                    int64_t n_users, int64_t n_items,
                                    float* U, float* V, float* B, int k,
                                    float lr, float reg, int use_bias, unsigned flags,
                                    int64_t* stats, void* stream);

/* Host-side restatement of RNGVector (recom_bpr.pyx:54-62): boost::random::mt19937 seeded
 * with `seed` + boost::random::uniform_int_distribution<long>(0, hi).  Pure host code, no
 * CUDA.  `fill` writes n consecutive draws from [0, hi] INCLUSIVE into host memory.        */
typedef struct b200_mt_sampler b200_mt_sampler;
B200_API b200_mt_sampler* b200_mt_sampler_create(uint32_t seed);
B200_API void b200_mt_sampler_destroy(b200_mt_sampler* s);
B200_API int b200_mt_sampler_fill_i64(b200_mt_sampler* s, int64_t hi, int64_t n, int64_t* out_host);
B200_API int b200_mt_sampler_fill_i32(b200_mt_sampler* s, int64_t hi, int64_t n, int32_t* out_host);

/* ------------------------------------------------------------------------------------
 * BPR siblings with a third item per sample, csrc/bprx.cu.  Common arguments:
 *   indptr / indices  device int32 CSR of the (purchase) interactions, rows sorted;  coo_row device int32[nnz] = row of
 *                     every interaction (BPR._prepare_data, recom_bpr.pyx:154-161)
 *   stats             device int64[2], accumulated: [0] correct (VEBPR only), [1] skipped
 *   *_epoch           one Hogwild epoch of n_samples samples drawn on the device (Philox keyed by seed, epoch) in the
 *                     reference's law; scatter with red.global.add
 *   *_epoch_replay    an explicit sample stream (device arrays, from *_draw_host) applied with the SERIAL result, arithmetic
 *                     in the reference's operation order and types: trained factors match the seeded single-thread
 *                     reference within 1e-4
 *   *_draw_host       (host, no CUDA) the seeded streams in the reference's RNG order; all pointers are HOST pointers
 *
 * VEBPR: replaces VEBPR._fit_sgd_viewloss (bpr/recom_vebpr.pyx:214-337).  view_indptr / view_indices = CSR of the
 * "viewed but not purchased" matrix (PurchaseViewDataset.view_matrix, sorted rows, same shape as the purchase matrix);
 * v_id = the sampled viewed item, -1 for a user without viewed items (BPR fall-back branch).  No item biases.        */
B200_API int b200_vebpr_epoch(const int32_t* indptr, const int32_t* indices, const int32_t* coo_row, int64_t n_users,
                              int64_t n_items, int64_t nnz, const int32_t* view_indptr, const int32_t* view_indices,
                              float* U, float* V, int k, float lr, float reg, float alpha, uint64_t seed, uint64_t epoch,
                              int64_t n_samples, int64_t* stats, void* stream);
B200_API int b200_vebpr_epoch_replay(const int64_t* i_index, const int32_t* v_id, const int32_t* j_id, int64_t n_samples,
                                     const int32_t* indptr, const int32_t* indices, const int32_t* coo_row,
                                     const int32_t* view_indptr, const int32_t* view_indices,
                                     float* U, float* V, int k, float lr, float reg, float alpha, int64_t* stats, void* stream);
B200_API int b200_vebpr_draw_host(b200_mt_sampler* pos, b200_mt_sampler* view, b200_mt_sampler* neg, int64_t nnz, int64_t n_items,
                                  const int32_t* coo_row, const int32_t* view_indptr, const int32_t* view_indices,
                                  int64_t n_samples, int64_t* i_index_out, int32_t* v_id_out, int32_t* j_id_out);
/* SBPR: replaces SBPR._fit_sgd (sbpr/recom_sbpr.pyx:193-300).  social_indptr / social_item_ids / social_item_counts =
 * the output of SBPR._prepare_social_data (:119-145): per user the items her friends have and she has not, and how many
 * friends have each; n_social = len(social_item_ids).  k_index = the sampled POSITION in social_item_ids (the entry is
 * read, and compared with j, also for users without social items, as in the reference; positions past the end compare
 * unequal).  lambda_u / lambda_v / lambda_b regularise users / items / biases; use_bias gates the bias updates of the
 * SBPR-2 branch only (the BPR fall-back branch always trains them, :263-264).                                        */
B200_API int b200_sbpr_epoch(const int32_t* indptr, const int32_t* indices, const int32_t* coo_row, int64_t n_users,
                             int64_t n_items, int64_t nnz, const int32_t* social_indptr, const int32_t* social_item_ids,
                             const int32_t* social_item_counts, int64_t n_social,
                             float* U, float* V, float* B, int k, float lr, float lambda_u, float lambda_v, float lambda_b,
                             int use_bias, uint64_t seed, uint64_t epoch, int64_t n_samples, int64_t* stats, void* stream);
B200_API int b200_sbpr_epoch_replay(const int64_t* i_index, const int32_t* j_id, const int64_t* k_index, int64_t n_samples,
                                    const int32_t* indptr, const int32_t* indices, const int32_t* coo_row,
                                    const int32_t* social_indptr, const int32_t* social_item_ids,
                                    const int32_t* social_item_counts, int64_t n_social,
                                    float* U, float* V, float* B, int k, float lr, float lambda_u, float lambda_v, float lambda_b,
                                    int use_bias, int64_t* stats, void* stream);
B200_API int b200_sbpr_draw_host(b200_mt_sampler* pos, b200_mt_sampler* neg, int64_t nnz, int64_t n_items,
                                 const int32_t* coo_row, const int32_t* social_indptr, int64_t n_samples,
                                 int64_t* i_index_out, int32_t* j_id_out, int64_t* k_index_out);

/* ------------------------------------------------------------------------------------
 * MF.  Replaces one epoch of backend_cpu.fit_sgd (mf/backend_cpu.pyx:58-83).
 *   rid, cid device int64[n] (the reference's INT64_t layout) or int32[n] when ids_are_i32
 *   val device f32[n]; U f32[n_users,k]; V f32[n_items,k]; Bu f32[n_users], Bi f32[n_items]
 *   ordered = 1: ratings are applied with the same result as the stored-order sequential
 *                loop (seeded reference); ordered = 0: Hogwild over the whole GPU.
 *   loss device f32[1]: receives sum(err^2) of the epoch (caller multiplies by 0.5,
 *                backend_cpu.pyx:85); it is overwritten, not accumulated.
 *   k = 0: the bias-only model -- one epoch of BaselineOnly._fit_sgd
 *                (baseline_only/recom_bo.pyx:121-131: r_pred = mu + Bu[u] + Bi[i]); U and V are
 *                not read and may be NULL, n_users / n_items are then given explicitly.     */
B200_API int b200_mf_epoch(const void* rid, const void* cid, const float* val, int64_t n, int ids_are_i32,
                           int64_t n_users, int64_t n_items, float* U, float* V, float* Bu, float* Bi, int k,
                           float lr, float reg, float mu, int use_bias, int ordered,
                           unsigned flags, float* loss, void* stream);

/* ------------------------------------------------------------------------------------
 * WMF.  One optimisation step of the reference's TensorFlow-1 graph (cornac/models/wmf/wmf.py:34-55, fed by
 * cornac/models/wmf/recom_wmf.py:186-199) for a mini-batch of `b` item ids:
 *   loss = sum(C (R_b - U V_b^T)^2) + lambda_u |U|^2/2 + lambda_v |V_b|^2/2,  C = a_conf where R_b != 0 else b_conf;
 *   gradients clipped elementwise to [-5, 5]; Adam with TF-1 semantics (dense step on U; the sparse step on V decays
 *   the moments of ALL rows and moves ALL rows).  The caller advances the beta powers and passes
 *   lr_t = lr * sqrt(1 - beta2^t) / (1 - beta1^t).
 *   csc_indptr / csc_rows / csc_vals  device int32[n_items+1] / int32[nnz] / f32[nnz]: train_set.csc_matrix
 *                (user indices sorted within a column); ids device int32[b], distinct
 *   U f32[n_users,k], V f32[n_items,k] and their Adam slots mU, vU, mV, vV (same shapes), all device, updated in place
 *   slot_of      device int32[n_items], all -1 on entry and on exit (scratch: item -> position in the batch)
 *   gV_scratch   device f32[b*k]; loss device f64[1]: receives the batch loss                                          */
B200_API int b200_wmf_step(const int32_t* csc_indptr, const int32_t* csc_rows, const float* csc_vals,
                           const int32_t* ids, int b, int64_t n_users, int64_t n_items, int k,
                           float* U, float* V, float* mU, float* vU, float* mV, float* vV,
                           float a_conf, float b_conf, float lambda_u, float lambda_v,
                           float lr_t, float beta1, float beta2, float epsilon,
                           int32_t* slot_of, float* gV_scratch, double* loss, void* stream);

/* ------------------------------------------------------------------------------------
 * Scores.  Replaces `out = base; fast_dot(U[u], V, out)` (fast_dot.pyx:40-43 as used by
 * BPR.score recom_bpr.pyx:290-293 and MF.score mf/recom_mf.py:272-278) for a BATCH of
 * query users:  out[q, i] = (item_base[i] + user_off[q]) + dot(U[user_idx[q]], V[i]).
 * The dot is accumulated in f64 in index order and rounded once to f32 (the defined
 * summation order shared with the oracle), so results are reproducible bit-for-bit.
 *   user_idx device int64[n_q] (NULL = rows 0..n_q-1 of U); item_base, user_off may be NULL
 *   out device f32[n_q, n_items]                                                            */
B200_API int b200_score_batch(const float* U, const int64_t* user_idx, int64_t n_q,
                              const float* V, int64_t n_items, int k,
                              const float* item_base, const float* user_off,
                              float* out, void* stream);

/* One user: out[i] = (item_base[i] + user_off) + dot(U[user_idx], V[i]) -- one call of fast_dot as BPR.score / MF.score
 * make it (fast_dot.pyx:25-43; user_off = mu + Bu[u] for MF, 0 for BPR).  Same arithmetic as b200_score_batch.          */
B200_API int b200_score(const float* U, int64_t user_idx, const float* V, int64_t n_items, int k,
                        const float* item_base, float user_off, float* out, void* stream);

/* Top-k of precomputed score rows.  Replaces the argpartition/argsort of
 * Recommender.rank (recommender.py:521-528) with a TOTAL order (score desc, id asc).
 *   scores device f32[n_q, n_items]; excl_indptr device int64[n_q+1] / excl_indices device
 *   int32 (sorted per row) list item ids removed from row q's candidates (NULL = none);
 *   out_ids device int32[n_q, topk] (-1 padded), out_scores device f32[n_q, topk].         */
B200_API int b200_topk_rows(const float* scores, int64_t n_q, int64_t n_items,
                            const int64_t* excl_indptr, const int32_t* excl_indices,
                            int topk, int32_t* out_ids, float* out_scores, void* stream);

/* Fused rank: scores (as b200_score_batch) + exclusion + top-k (as b200_topk_rows) for a
 * batch of users without materialising the [n_q, n_items] score matrix.  Tensor-core
 * (tcgen05) candidate pass + exact f64 re-score of the candidates; ids and scores are
 * identical to b200_score_batch followed by b200_topk_rows.
 *   workspace: device scratch of b200_rank_topk_workspace_bytes(...) bytes.                */
B200_API int64_t b200_rank_topk_workspace_bytes(int64_t n_q, int64_t n_items, int k, int topk);
B200_API int b200_rank_topk(const float* U, const int64_t* user_idx, int64_t n_q,
                            const float* V, int64_t n_items, int k,
                            const float* item_base, const float* user_off,
                            const int64_t* excl_indptr, const int32_t* excl_indices,
                            int topk, int32_t* out_ids, float* out_scores,
                            void* workspace, int64_t workspace_bytes, void* stream);

/* The item side of the fused rank, packed once: fp16 tile images of V (with the item base folded in) and the scaling
 * scalars.  V and item_base are constant across an evaluation / a serving session, so a caller that ranks many batches
 * builds this once (b200_rank_pack_items, ~0.4 ms at 1 M items x k = 128) and passes it to b200_rank_topk_packed, which
 * then skips the two passes over V every b200_rank_topk call makes.  The caller owns the buffer and must rebuild it
 * whenever V or item_base change (e.g. after a training epoch).  b200_rank_items_bytes returns 0 for shapes the
 * tensor-core pass does not take (then pass packed_items = NULL).
 *   packed device, 128-byte aligned, b200_rank_items_bytes(n_items, k) bytes                                              */
B200_API int64_t b200_rank_items_bytes(int64_t n_items, int k);
B200_API int b200_rank_pack_items(const float* V, int64_t n_items, int k, const float* item_base,
                                  void* packed, int64_t packed_bytes, void* stream);
B200_API int b200_rank_topk_packed(const float* U, const int64_t* user_idx, int64_t n_q,
                                   const float* V, int64_t n_items, int k,
                                   const float* item_base, const float* user_off,
                                   const int64_t* excl_indptr, const int32_t* excl_indices,
                                   int topk, int32_t* out_ids, float* out_scores,
                                   const void* packed_items, void* workspace, int64_t workspace_bytes, void* stream);

/* Validation hook of the tensor-core pass: dense APPROXIMATE scores (operands scaled by powers of
 * two and rounded to fp16, f32 accumulation, + item_base, scaled back) as the candidate pass of
 * b200_rank_topk sees them; padding items (>= n_items) read -inf.
 *   out device f32[ceil(n_q/128)*128, ceil(n_items/256)*256] row-major; workspace as above. */
B200_API int b200_rank_tc_debug_scores(const float* U, int64_t n_q, const float* V, int64_t n_items, int k,
                                       const float* item_base, float* out, int64_t out_elems,
                                       void* workspace, int64_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------------------
 * Top-k ranking metrics of a batch of ranked lists (the metric half of the per-user loop of
 * `ranking_eval`, cornac/eval_methods/base_method.py:200-220, for the metrics that only read
 * pd_rank[:k]: cornac/metrics/ranking.py:67-123 NDCG, :126-178 NCRR, :240-275 MeasureAtK ->
 * HitRatio / Precision / Recall / FMeasure).  A hit is `ids[q, r] in positives of user q`.
 *   ids          device int32[n_q, ids_stride], the first `topk` of each row are the ranked
 *                list (b200_rank_topk output; -1 = padding, never a hit)
 *   user_idx     device int64[n_q] row of the positives CSR for list q (NULL = q)
 *   pos_indptr / pos_indices  device int64[.] / int32[.] CSR of the test positives, ids sorted
 *                per row; every listed user must have >= 1 positive (base_method.py:180-182)
 *   metric_kind / metric_k    device int32[n_metrics]: B200_METRIC_* and its k (1 <= k)
 *   out          device f64[n_metrics, n_q]: the value metric.compute() returns for each user */
#define B200_METRIC_NDCG 0
#define B200_METRIC_PRECISION 1
#define B200_METRIC_RECALL 2
#define B200_METRIC_FMEASURE 3
#define B200_METRIC_HIT 4
#define B200_METRIC_NCRR 5
B200_API int b200_topk_metrics(const int32_t* ids, int64_t n_q, int topk, int64_t ids_stride,
                               const int64_t* user_idx, const int64_t* pos_indptr, const int32_t* pos_indices,
                               const int32_t* metric_kind, const int32_t* metric_k, int n_metrics,
                               double* out, void* stream);

/* The counts behind the full-vector ranking metrics of the same loop -- AUC (cornac/metrics/ranking.py:473-485), MAP
 * (:522-525), MRR (:213-222) -- for a batch of users whose score rows are on the device (b200_score_batch output):
 *   scores       device f32[n_q, n_items]; MODIFIED: the entries listed in excl_* are overwritten with NaN (not candidates)
 *   excl_indptr / excl_indices   device int64[n_q+1] / int32: per ROW q, the item ids that are not candidates (NULL = none)
 *   user_idx / pos_indptr / pos_indices   as for b200_topk_metrics: the test positives of row q are row user_idx[q] of the CSR
 *   less         device int64, indexed like pos_indices: number of candidates of the user scoring strictly BELOW that positive
 *   pos_score    device f32, indexed like pos_indices: the positive's score
 *   n_cand       device int64[n_q]: candidates of the user (items minus exclusions)
 *   before_first device int64[n_q]: candidates ranked ahead of the user's best positive in the total order
 *                (score desc, id asc) -> MRR = 1 / (1 + before_first)
 * From these: rank_p = n_cand - less_p (rankdata "max"), AUC = sum_p (less_p - #{positives below p}) / (|P| (n_cand - |P|)). */
B200_API int b200_rank_counts(float* scores, int64_t n_q, int64_t n_items,
                              const int64_t* excl_indptr, const int32_t* excl_indices,
                              const int64_t* user_idx, const int64_t* pos_indptr, const int32_t* pos_indices,
                              int64_t* less, float* pos_score, int64_t* n_cand, int64_t* before_first, void* stream);

/* ------------------------------------------------------------------------------------
 * Multi-GPU item-factor exchange (no reference counterpart: the reference is a single
 * process).  Each rank trains its user shard against a replica of V / B; at the epoch
 * boundary   b200_delta_make:  delta[i] = x[i] - snapshot[i]
 * the caller all-reduces (sum) `delta` over NCCL, then
 *            b200_delta_apply: x[i] = snapshot[i] + delta[i];  snapshot[i] = x[i]
 * so every replica ends the epoch with x_start + sum over ranks of the local changes -- or, the default of the Python layer,
 * x_start + (sum of the changes) / (number of ranks that changed the element): the caller also all-reduces the indicator
 * (delta != 0) and divides.  The plain sum is the single-process step count only while the ranks change different rows; a row
 * every rank trains (a popular item) moves `world` times too far and the epochs oscillate with growing amplitude from 4 ranks
 * on (tools/sim_localsgd.py: pairwise accuracy 0.81 -> 0.30 at 4 ranks); the mean over the ranks that changed an element is a
 * convex combination of their local results: stable at any world size, equal to the sum where one rank alone touched it.   */
B200_API int b200_delta_make(const float* x, const float* snapshot, float* delta, int64_t n, void* stream);
B200_API int b200_delta_apply(float* x, float* snapshot, const float* delta, int64_t n, void* stream);

/* The same exchange as ONE kernel over NVLink peer memory (one process per GPU, replicas mapped into each other with
 * CUDA IPC): rank r owns the slice b200_item_exchange_slice(r, world, n) of the vector, reads that slice of EVERY
 * replica over NVLink, forms  snapshot + sum_r (x_r - snapshot)  in rank order (deterministic, bit-equal everywhere) and
 * stores it into every replica and into its snapshot -- delta, reduce-scatter, apply and all-gather fused; each byte
 * crosses NVLink once per direction and there is no delta buffer.
 *   b200_ipc_export   (host) 64-byte CUDA IPC handle of the allocation containing dev_ptr + the pointer's offset in it
 *   b200_ipc_open     (host) map a peer's exported allocation; returns the peer pointer (peer access enabled lazily)
 *   x_peers / flag_peers  host arrays of `world` DEVICE pointers: every rank's replica (f32[n]) and flag buffer
 *                     (u32[32], zeroed once by its owner before the first exchange), own pointers at [rank]
 *   snapshot_slice    device f32[hi - lo]: the epoch-start values of the owned slice (= the replica's after each exchange)
 *   seq               1, 2, 3, ... : the same value on every rank for the same exchange
 *   mean_touched      1: snapshot + (sum_r d_r) / #{r : d_r != 0} per element (see above; the default of the Python layer); 0: the sum
 * Every rank must call it once per exchange; the kernel returns when all peers have finished writing this rank's
 * replica.  A peer that never arrives sets flag word [17] after ~4 s instead of hanging the GPU. */
B200_API int b200_ipc_export(const void* dev_ptr, void* handle64_out, int64_t* offset_out);
B200_API int b200_ipc_open(const void* handle64, int64_t offset, void** mapped_out);
B200_API int b200_ipc_close(void* mapped, int64_t offset);
B200_API int b200_item_exchange_slice(int rank, int world, int64_t n, int64_t* lo_out, int64_t* hi_out);
B200_API int b200_item_exchange(int rank, int world, void* const* x_peers, void* const* flag_peers, float* snapshot_slice,
                                int64_t n, uint32_t seq, int mean_touched, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* B200CORNAC_H_ */
