// BPR SGD epochs for sm_100a.
//
// Replaces BPR._fit_sgd (reference: cornac/models/bpr/recom_bpr.pyx:208-269) with
//   * bpr_hogwild_kernel : throughput mode.  Persistent grid; every G-lane group draws its
//     own (u, i+, j-) triplets from the CSR matrix with a counter-based RNG, gathers the
//     three factor rows with 128-bit L2-only loads, reduces the pairwise dot with warp
//     shuffles and scatters the update back (plain 128-bit stores = the reference's
//     lock-free Hogwild, or red.global.add.v4.f32 when B200_SGD_ATOMIC is set).
//   * bpr_replay_kernel  : parity mode.  One warp applies an explicit sample stream with
//     the same result as the sequential seeded reference (num_threads = 1,
//     recom_bpr.pyx:132-133): sample metadata (u, i, j, skip test) is resolved 32 samples
//     at a time in parallel (read-only data), the updates are applied strictly in order.
//
// HBM-bound integer/gather work: no tensor cores here by design.
#include "sgd_common.cuh"

namespace b200 {

struct BprParams {
    const int32_t* __restrict__ indptr;
    const int32_t* __restrict__ indices;
    const int32_t* __restrict__ coo_row;
    int64_t nnz;
    int64_t n_neg;
    int64_t n_samples;
    float* U;
    float* V;
    float* B;
    int k;
    float lr, reg;
    int use_bias;
    uint32_t seed_lo, seed_hi;
    uint32_t epoch_lo, epoch_hi;
    uint64_t sample_base;
    unsigned long long* stats;   // {correct, skipped}
};

// z = 1 / (1 + exp(score))   (recom_bpr.pyx:252)
template <bool EXACT>
__device__ __forceinline__ float bpr_z(float score)
{
    if (EXACT) return (float)(1.0 / (1.0 + exp((double)score)));
    return __frcp_rn(1.f + __expf(score));
}

template <int G, int NPL, bool VEC, bool ATOMIC, bool EXACT, int S>
__global__ void __launch_bounds__(256) bpr_hogwild_kernel(const BprParams p)
{
    using Frag = RowFrag<NPL, VEC>;
    constexpr int E = NPL * Frag::W;
    const int lg = threadIdx.x & (G - 1);
    const int n_units = VEC ? p.k / 4 : p.k;
    const int64_t groups_per_block = blockDim.x / G;
    const int64_t n_groups = (int64_t)gridDim.x * groups_per_block;
    const int64_t gid = (int64_t)blockIdx.x * groups_per_block + threadIdx.x / G;
    const size_t k = (size_t)p.k;

    unsigned int n_correct = 0, n_skipped = 0;

    for (int64_t s0 = gid * S; s0 < p.n_samples; s0 += n_groups * S) {
        int32_t u[S], it[S], jt[S];
        int64_t lo[S], hi[S];
        bool live[S];
        // ---- phase A: draw the triplets (every lane of the group computes the same values)
#pragma unroll
        for (int t = 0; t < S; ++t) {
            const uint64_t s = p.sample_base + (uint64_t)(s0 + t);
            live[t] = (s0 + t) < p.n_samples;
            Philox4 r = philox4x32_10((uint32_t)s, (uint32_t)(s >> 32), p.epoch_lo, p.epoch_hi, p.seed_lo, p.seed_hi);
            const int64_t ii = (int64_t)range64(r.x, r.y, (uint64_t)p.nnz);
            jt[t] = (int32_t)range64(r.z, r.w, (uint64_t)p.n_neg);
            u[t] = __ldg(p.coo_row + ii);
            it[t] = __ldg(p.indices + ii);
        }
        // ---- phase B: row bounds + all factor-row gathers in flight before the skip test resolves
        Frag fu[S], fi[S], fj[S];
        float bi[S], bj[S];
#pragma unroll
        for (int t = 0; t < S; ++t) {
            lo[t] = __ldg(p.indptr + u[t]);
            hi[t] = __ldg(p.indptr + u[t] + 1);
            row_load<G, NPL, VEC>(fu[t], p.U + (size_t)u[t] * k, lg, n_units);
            row_load<G, NPL, VEC>(fi[t], p.V + (size_t)it[t] * k, lg, n_units);
            row_load<G, NPL, VEC>(fj[t], p.V + (size_t)jt[t] * k, lg, n_units);
            bi[t] = __ldcg(p.B + it[t]);
            bj[t] = __ldcg(p.B + jt[t]);
        }
        // ---- phase C: has_non_zero(u, j)  (recom_bpr.pyx:241-243)
#pragma unroll
        for (int t = 0; t < S; ++t) {
            if (live[t] && row_contains(p.indices, lo[t], hi[t], jt[t])) {
                live[t] = false;
                ++n_skipped;
            }
        }
        // ---- phase D: score, z, update (recom_bpr.pyx:249-267)
#pragma unroll
        for (int t = 0; t < S; ++t) {
            float part = 0.f;
#pragma unroll
            for (int e = 0; e < E; ++e) part = fmaf(fu[t].v[e], fi[t].v[e] - fj[t].v[e], part);
            const float score = (bi[t] - bj[t]) + group_sum<G>(part);
            if (!live[t]) continue;     // group-uniform
            const float z = bpr_z<EXACT>(score);
            n_correct += (z < .5f);
            const float lr = p.lr, reg = p.reg;
            float* pu = p.U + (size_t)u[t] * k;
            float* pi = p.V + (size_t)it[t] * k;
            float* pj = p.V + (size_t)jt[t] * k;
            if (ATOMIC) {
                Frag du, di, dj;
#pragma unroll
                for (int e = 0; e < E; ++e) {
                    const float uf = fu[t].v[e], vi = fi[t].v[e], vj = fj[t].v[e];
                    du.v[e] = lr * (z * (vi - vj) - reg * uf);
                    di.v[e] = lr * (z * uf - reg * vi);
                    dj.v[e] = lr * (-z * uf - reg * vj);
                }
                row_red_add<G, NPL, VEC>(du, pu, lg, n_units);
                row_red_add<G, NPL, VEC>(di, pi, lg, n_units);
                row_red_add<G, NPL, VEC>(dj, pj, lg, n_units);
                if (p.use_bias && lg == 0) {
                    red_add_f32(p.B + it[t], lr * (z - reg * bi[t]));
                    red_add_f32(p.B + jt[t], lr * (-z - reg * bj[t]));
                }
            } else {
#pragma unroll
                for (int e = 0; e < E; ++e) {
                    const float uf = fu[t].v[e], vi = fi[t].v[e], vj = fj[t].v[e];
                    fu[t].v[e] = uf + lr * (z * (vi - vj) - reg * uf);
                    fi[t].v[e] = vi + lr * (z * uf - reg * vi);
                    fj[t].v[e] = vj + lr * (-z * uf - reg * vj);
                }
                row_store<G, NPL, VEC>(fu[t], pu, lg, n_units);
                row_store<G, NPL, VEC>(fi[t], pi, lg, n_units);
                row_store<G, NPL, VEC>(fj[t], pj, lg, n_units);
                if (p.use_bias && lg == 0) {
                    __stcg(p.B + it[t], bi[t] + lr * (z - reg * bi[t]));
                    __stcg(p.B + jt[t], bj[t] + lr * (-z - reg * bj[t]));
                }
            }
        }
    }

    // ---- epoch statistics: one count per group (its lane 0), block-reduced, two atomics per block
    __shared__ unsigned int sh_stats[2];
    if (threadIdx.x < 2) sh_stats[threadIdx.x] = 0;
    __syncthreads();
    unsigned int c = (lg == 0) ? n_correct : 0u, sk = (lg == 0) ? n_skipped : 0u;
    c = __reduce_add_sync(0xffffffffu, c);
    sk = __reduce_add_sync(0xffffffffu, sk);
    if ((threadIdx.x & 31) == 0) {
        atomicAdd(&sh_stats[0], c);
        atomicAdd(&sh_stats[1], sk);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        atomicAdd(p.stats + 0, (unsigned long long)sh_stats[0]);
        atomicAdd(p.stats + 1, (unsigned long long)sh_stats[1]);
    }
}

// ---------------------------------------------------------------------------------------
// Parity mode: one warp, serial-equivalent.  Unfused f32 arithmetic in the operation order
// of recom_bpr.pyx:249-267 (the dot is a lane-strided partial sum + shuffle tree).
struct ReplayParams {
    const int64_t* __restrict__ i_index;
    const int32_t* __restrict__ j_id;
    int64_t n_samples;
    const int32_t* __restrict__ indptr;
    const int32_t* __restrict__ indices;
    const int32_t* __restrict__ coo_row;
    float* U;
    float* V;
    float* B;
    int k;
    float lr, reg;
    int use_bias;
    unsigned long long* stats;
};

__global__ void __launch_bounds__(32) bpr_replay_kernel(const ReplayParams p)
{
    const int lane = threadIdx.x;
    const size_t k = (size_t)p.k;
    unsigned long long n_correct = 0, n_skipped = 0;
    for (int64_t base = 0; base < p.n_samples; base += 32) {
        // metadata of 32 consecutive samples, one per lane (read-only inputs => order-free)
        const int64_t s = base + lane;
        int32_t mu = 0, mi = 0, mj = 0;
        bool mskip = true;
        if (s < p.n_samples) {
            const int64_t ii = p.i_index[s];
            mj = p.j_id[s];
            mu = __ldg(p.coo_row + ii);
            mi = __ldg(p.indices + ii);
            mskip = row_contains(p.indices, __ldg(p.indptr + mu), __ldg(p.indptr + mu + 1), mj);
        }
        const int n_here = (int)min((int64_t)32, p.n_samples - base);
        for (int t = 0; t < n_here; ++t) {
            const bool skip = __shfl_sync(0xffffffffu, (int)mskip, t) != 0;
            if (skip) { ++n_skipped; continue; }
            const int32_t u = __shfl_sync(0xffffffffu, mu, t);
            const int32_t i = __shfl_sync(0xffffffffu, mi, t);
            const int32_t j = __shfl_sync(0xffffffffu, mj, t);
            float* pu = p.U + (size_t)u * k;
            float* pi = p.V + (size_t)i * k;
            float* pj = p.V + (size_t)j * k;
            const float bi = __ldcg(p.B + i), bj = __ldcg(p.B + j);
            float part = 0.f;
            for (int f = lane; f < p.k; f += 32)
                part = __fadd_rn(part, __fmul_rn(__ldcg(pu + f), __fsub_rn(__ldcg(pi + f), __ldcg(pj + f))));
            const float score = __fadd_rn(__fsub_rn(bi, bj), group_sum<32>(part));
            const float z = (float)(1.0 / (1.0 + exp((double)score)));
            n_correct += (z < .5f);
            const float lr = p.lr, reg = p.reg;
            for (int f = lane; f < p.k; f += 32) {
                const float uf = __ldcg(pu + f), vi = __ldcg(pi + f), vj = __ldcg(pj + f);
                __stcg(pu + f, __fadd_rn(uf, __fmul_rn(lr, __fsub_rn(__fmul_rn(z, __fsub_rn(vi, vj)), __fmul_rn(reg, uf)))));
                __stcg(pi + f, __fadd_rn(vi, __fmul_rn(lr, __fsub_rn(__fmul_rn(z, uf), __fmul_rn(reg, vi)))));
                __stcg(pj + f, __fadd_rn(vj, __fmul_rn(lr, __fsub_rn(__fmul_rn(-z, uf), __fmul_rn(reg, vj)))));
            }
            if (p.use_bias && lane == 0) {
                __stcg(p.B + i, __fadd_rn(bi, __fmul_rn(lr, __fsub_rn(z, __fmul_rn(reg, bi)))));
                __stcg(p.B + j, __fadd_rn(bj, __fmul_rn(lr, __fsub_rn(-z, __fmul_rn(reg, bj)))));
            }
            __syncwarp();   // order lane 0's bias stores before the next sample's reads
        }
    }
    if (lane == 0) {
        atomicAdd(p.stats + 0, n_correct);
        atomicAdd(p.stats + 1, n_skipped);
    }
}

template <int G, int NPL, bool VEC, bool ATOMIC, bool EXACT>
static int launch_hogwild(const BprParams& p, cudaStream_t st)
{
    // samples in flight per group: keep the register footprint of the 3*S row fragments moderate
    constexpr int E = NPL * (VEC ? 4 : 1);
    constexpr int S = (E <= 4) ? 2 : 1;
    auto kern = bpr_hogwild_kernel<G, NPL, VEC, ATOMIC, EXACT, S>;
    const int threads = 256;
    int occ = 0;
    B200_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, threads, 0));
    if (occ < 1) occ = 1;
    const int64_t groups_per_block = threads / G;
    int64_t want = (p.n_samples + groups_per_block * S - 1) / (groups_per_block * S);
    int64_t grid = (int64_t)sm_count() * occ;
    if (want < grid) grid = want < 1 ? 1 : want;
    kern<<<(unsigned)grid, threads, 0, st>>>(p);
    B200_CUDA(cudaGetLastError());
    return B200_OK;
}

}  // namespace b200

using namespace b200;

extern "C" int b200_bpr_epoch(const int32_t* indptr, const int32_t* indices, const int32_t* coo_row,
                              int64_t nnz, int64_t n_neg, int64_t n_samples,
                              float* U, float* V, float* B, int k,
                              float lr, float reg, int use_bias,
                              uint64_t seed, uint64_t epoch, uint64_t sample_base,
                              unsigned flags, int64_t* stats, void* stream)
{
    B200_REQUIRE(indptr && indices && coo_row && U && V && B && stats, "b200_bpr_epoch: null pointer argument");
    B200_REQUIRE(k >= 1 && k <= 1024, "b200_bpr_epoch: k=%d out of range [1, 1024]", k);
    B200_REQUIRE(nnz >= 0 && n_neg >= 1 && n_samples >= 0, "b200_bpr_epoch: bad sizes nnz=%lld n_neg=%lld n_samples=%lld",
                 (long long)nnz, (long long)n_neg, (long long)n_samples);
    if (n_samples == 0 || nnz == 0) return B200_OK;
    const RowLayout L = pick_layout(k);
    B200_REQUIRE(L.npl <= 8, "b200_bpr_epoch: k=%d not supported (scalar rows are limited to k <= 256)", k);
    if (L.vec) {
        B200_REQUIRE((((uintptr_t)U | (uintptr_t)V) & 15) == 0, "b200_bpr_epoch: U/V must be 16-byte aligned");
    }
    BprParams p;
    p.indptr = indptr; p.indices = indices; p.coo_row = coo_row;
    p.nnz = nnz; p.n_neg = n_neg; p.n_samples = n_samples;
    p.U = U; p.V = V; p.B = B; p.k = k; p.lr = lr; p.reg = reg; p.use_bias = use_bias;
    p.seed_lo = (uint32_t)seed; p.seed_hi = (uint32_t)(seed >> 32);
    p.epoch_lo = (uint32_t)epoch; p.epoch_hi = (uint32_t)(epoch >> 32);
    p.sample_base = sample_base;
    p.stats = reinterpret_cast<unsigned long long*>(stats);
    cudaStream_t st = (cudaStream_t)stream;
    const bool atomic = flags & B200_SGD_ATOMIC, exact = flags & B200_SGD_EXACT_EXP;
#define CALL(G_, NPL_, VEC_)                                                                      \
    do {                                                                                          \
        int rc;                                                                                   \
        if (atomic) rc = exact ? launch_hogwild<G_, NPL_, VEC_, true, true>(p, st)                \
                               : launch_hogwild<G_, NPL_, VEC_, true, false>(p, st);              \
        else        rc = exact ? launch_hogwild<G_, NPL_, VEC_, false, true>(p, st)               \
                               : launch_hogwild<G_, NPL_, VEC_, false, false>(p, st);             \
        if (rc) return rc;                                                                        \
    } while (0)
    B200_DISPATCH_LAYOUT(L, CALL);
#undef CALL
    return B200_OK;
}

extern "C" int b200_bpr_epoch_replay(const int64_t* i_index, const int32_t* j_id, int64_t n_samples,
                                     const int32_t* indptr, const int32_t* indices, const int32_t* coo_row,
                                     float* U, float* V, float* B, int k,
                                     float lr, float reg, int use_bias,
                                     int64_t* stats, void* stream)
{
    B200_REQUIRE(i_index && j_id && indptr && indices && coo_row && U && V && B && stats,
                 "b200_bpr_epoch_replay: null pointer argument");
    B200_REQUIRE(k >= 1, "b200_bpr_epoch_replay: k=%d", k);
    if (n_samples <= 0) return B200_OK;
    ReplayParams p;
    p.i_index = i_index; p.j_id = j_id; p.n_samples = n_samples;
    p.indptr = indptr; p.indices = indices; p.coo_row = coo_row;
    p.U = U; p.V = V; p.B = B; p.k = k; p.lr = lr; p.reg = reg; p.use_bias = use_bias;
    p.stats = reinterpret_cast<unsigned long long*>(stats);
    bpr_replay_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(p);
    B200_CUDA(cudaGetLastError());
    return B200_OK;
}
