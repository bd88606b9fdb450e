// Top-k ranking metrics of a batch of ranked lists, on the device.
//
// Reference: the body of the per-user loop of `ranking_eval`
// (cornac/eval_methods/base_method.py:169-220) calls `metric.compute(gt_pos, pd_rank)` for every metric;
// for the @k metrics (cornac/metrics/ranking.py:67-123 NDCG, :126-178 NCRR, :240-275 MeasureAtK and its
// subclasses HitRatio / Precision / Recall / FMeasure) only `np.isin(pd_rank[:k], gt_pos)` matters.
// Here: one warp per ranked list; the hit bitmap of the list against the user's sorted test positives is
// built once in shared memory (coalesced id loads, binary search per lane), then every metric reduces a
// prefix of the bitmap.  HBM traffic = the ids (4 B each) + the positives the searches touch; integer work
// except for the final f64 ratios.
#include "common.cuh"

namespace {

constexpr int EV_THREADS = 256;
constexpr int EV_WARPS = EV_THREADS / 32;
constexpr int EV_MAX_TOPK = 4096;
constexpr int EV_MAX_METRICS = 32;

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

__global__ void __launch_bounds__(EV_THREADS)
topk_metrics_kernel(const int32_t* __restrict__ ids, long long n_q, int topk, long long ids_stride,
                    const long long* __restrict__ user_idx, const long long* __restrict__ pos_indptr,
                    const int32_t* __restrict__ pos_indices, const int32_t* __restrict__ kinds,
                    const int32_t* __restrict__ ks, int n_metrics, double* __restrict__ out) {
    extern __shared__ double disc[];                       // [topk] 1/log2(r+2), then per-warp bitmaps
    uint32_t* bitmap = reinterpret_cast<uint32_t*>(disc + topk) + (threadIdx.x >> 5) * ((topk + 31) / 32);
    __shared__ int s_kind[EV_MAX_METRICS], s_k[EV_MAX_METRICS];
    for (int r = threadIdx.x; r < topk; r += EV_THREADS) disc[r] = 1.0 / log2((double)(r + 2));
    if (threadIdx.x < n_metrics) {
        s_kind[threadIdx.x] = kinds[threadIdx.x];
        int k = ks[threadIdx.x];
        s_k[threadIdx.x] = k < topk ? k : topk;
    }
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const int words = (topk + 31) / 32;
    for (long long q = (long long)blockIdx.x * EV_WARPS + (threadIdx.x >> 5); q < n_q;
         q += (long long)gridDim.x * EV_WARPS) {
        const long long u = user_idx ? user_idx[q] : q;
        const long long lo = pos_indptr[u];
        const int npos = (int)(pos_indptr[u + 1] - lo);
        const int32_t* pos = pos_indices + lo;
        for (int w = 0; w < words; ++w) {
            int r = w * 32 + lane;
            int id = r < topk ? ids[q * ids_stride + r] : -1;
            bool hit = false;
            if (id >= 0) {
                int a = 0, b = npos;                       // lower bound in the sorted positives
                while (a < b) {
                    int m = (a + b) >> 1;
                    if (pos[m] < id) a = m + 1; else b = m;
                }
                hit = a < npos && pos[a] == id;
            }
            uint32_t bits = __ballot_sync(0xffffffffu, hit);
            if (lane == 0) bitmap[w] = bits;
        }
        __syncwarp();
        for (int m = 0; m < n_metrics; ++m) {
            const int kind = s_kind[m], k = s_k[m];
            const int k_nominal = ks[m];                   // tp_fp uses the requested k (ranking.py:272)
            int tp = 0;
            double dcg = 0.0, crr = 0.0;
            for (int r = lane; r < k; r += 32) {
                if ((bitmap[r >> 5] >> (r & 31)) & 1u) {
                    ++tp;
                    dcg += disc[r];
                    crr += 1.0 / (double)(r + 1);
                }
            }
            tp = __reduce_add_sync(0xffffffffu, tp);
            double val = 0.0;
            if (kind == B200_METRIC_NDCG || kind == B200_METRIC_NCRR) {
                const bool nd = kind == B200_METRIC_NDCG;
                double num = warp_sum(nd ? dcg : crr);
                int ideal = npos < k ? npos : k;
                double den = 0.0;
                for (int r = lane; r < ideal; r += 32) den += nd ? disc[r] : 1.0 / (double)(r + 1);
                den = warp_sum(den);
                val = num / den;                           // npos >= 1 is the caller's contract (base_method.py:180)
            } else {
                double prec = (double)tp / (double)k_nominal;
                double rec = (double)tp / (double)npos;
                if (kind == B200_METRIC_PRECISION) val = prec;
                else if (kind == B200_METRIC_RECALL) val = rec;
                else if (kind == B200_METRIC_HIT) val = tp > 0 ? 1.0 : 0.0;
                else val = (prec + rec) > 0.0 ? 2.0 * prec * rec / (prec + rec) : 0.0;   // ranking.py FMeasure
            }
            if (lane == 0) out[(long long)m * n_q + q] = val;
        }
        __syncwarp();
    }
}

}  // namespace

extern "C" int b200_topk_metrics(const int32_t* ids, int64_t n_q, int topk, int64_t ids_stride,
                                 const int64_t* user_idx, const int64_t* pos_indptr, const int32_t* pos_indices,
                                 const int32_t* metric_kind, const int32_t* metric_k, int n_metrics,
                                 double* out, void* stream) {
    B200_REQUIRE(ids && pos_indptr && pos_indices && metric_kind && metric_k && out, "null pointer");
    B200_REQUIRE(topk >= 1 && topk <= EV_MAX_TOPK, "topk out of range [1, 4096]");
    B200_REQUIRE(n_metrics >= 1 && n_metrics <= EV_MAX_METRICS, "n_metrics out of range [1, 32]");
    B200_REQUIRE(ids_stride >= topk, "ids_stride < topk");
    if (n_q == 0) return 0;
    const int sms = ::b200::sm_count();
    size_t smem = (size_t)topk * sizeof(double) + (size_t)EV_WARPS * ((topk + 31) / 32) * sizeof(uint32_t);
    long long blocks = (n_q + EV_WARPS - 1) / EV_WARPS;
    long long cap = (long long)sms * 8;
    if (blocks > cap) blocks = cap;
    topk_metrics_kernel<<<(unsigned)blocks, EV_THREADS, smem, (cudaStream_t)stream>>>(
        ids, (long long)n_q, topk, (long long)ids_stride, (const long long*)user_idx,
        (const long long*)pos_indptr, pos_indices, metric_kind, metric_k, n_metrics, out); ::b200::count_launch();
    B200_CUDA(cudaGetLastError());
    return 0;
}
