// Host-side sampler for the seeded (deterministic) BPR mode.
//
// Restates RNGVector (reference: cornac/models/bpr/recom_bpr.pyx:54-62), i.e.
// boost::random::mt19937 + boost::random::uniform_int_distribution<long>(0, hi)
// (cornac/utils/external/boost/random/uniform_int_distribution.hpp:49-228): the engine is
// the standard 32-bit Mersenne Twister; a draw from [0, hi] with hi < 2^32-1 divides the
// engine output by a bucket size and rejects values above hi, so the number of engine
// outputs consumed per draw is data dependent and the stream is inherently sequential.
// The reference's sampler is compiled C++, so this one is C++ as well (not CUDA): the
// draws of an epoch are produced here and shipped to the GPU replay kernel.
#include <stdint.h>

#include <new>

#include "../../include/b200cornac.h"

namespace b200 {
void set_error(const char* fmt, ...);
}

struct b200_mt_sampler {
    uint32_t state[624];
    int pos;

    explicit b200_mt_sampler(uint32_t seed)
    {
        state[0] = seed;
        for (uint32_t i = 1; i < 624; ++i) state[i] = 1812433253u * (state[i - 1] ^ (state[i - 1] >> 30)) + i;
        pos = 624;
    }

    void regenerate()
    {
        constexpr uint32_t UP = 0x80000000u, LO = 0x7fffffffu, A = 0x9908b0dfu;
        int i = 0;
        for (; i < 624 - 397; ++i) {
            const uint32_t y = (state[i] & UP) | (state[i + 1] & LO);
            state[i] = state[i + 397] ^ (y >> 1) ^ ((y & 1u) ? A : 0u);
        }
        for (; i < 623; ++i) {
            const uint32_t y = (state[i] & UP) | (state[i + 1] & LO);
            state[i] = state[i + 397 - 624] ^ (y >> 1) ^ ((y & 1u) ? A : 0u);
        }
        const uint32_t y = (state[623] & UP) | (state[0] & LO);
        state[623] = state[396] ^ (y >> 1) ^ ((y & 1u) ? A : 0u);
        pos = 0;
    }

    inline uint32_t next()
    {
        if (pos >= 624) regenerate();
        uint32_t y = state[pos++];
        y ^= y >> 11;
        y ^= (y << 7) & 0x9d2c5680u;
        y ^= (y << 15) & 0xefc60000u;
        y ^= y >> 18;
        return y;
    }

    // uniform integer in [0, range], range as an unsigned 64-bit span
    uint64_t draw(uint64_t range)
    {
        const uint64_t brange = 0xFFFFFFFFull;
        if (range == 0) return 0;
        if (range == brange) return next();
        if (range < brange) {
            const uint32_t r32 = (uint32_t)range;
            uint32_t bucket = 0xFFFFFFFFu / (r32 + 1u);
            if (0xFFFFFFFFu % (r32 + 1u) == r32) ++bucket;
            for (;;) {
                const uint32_t v = next() / bucket;
                if (v <= r32) return v;
            }
        }
        // span wider than the engine: base-2^32 digits + one recursive top digit, with rejection
        for (;;) {
            uint64_t limit;
            if (range == UINT64_MAX) {
                limit = range / (brange + 1);
                if (range % (brange + 1) == brange) ++limit;
            } else {
                limit = (range + 1) / (brange + 1);
            }
            uint64_t result = 0, mult = 1;
            bool exact_power = false;
            while (mult <= limit) {
                result += (uint64_t)next() * mult;
                if (mult * brange == range - mult + 1) { exact_power = true; break; }
                mult *= brange + 1;
            }
            if (exact_power) return result;
            uint64_t top = draw(range / mult);
            if (UINT64_MAX / mult < top) continue;
            top *= mult;
            result += top;
            if (result < top) continue;
            if (result > range) continue;
            return result;
        }
    }
};

extern "C" b200_mt_sampler* b200_mt_sampler_create(uint32_t seed)
{
    return new (std::nothrow) b200_mt_sampler(seed);
}

extern "C" void b200_mt_sampler_destroy(b200_mt_sampler* s) { delete s; }

extern "C" int b200_mt_sampler_fill_i64(b200_mt_sampler* s, int64_t hi, int64_t n, int64_t* out)
{
    if (!s || hi < 0 || n < 0 || (n > 0 && !out)) {
        b200::set_error("b200_mt_sampler_fill_i64: bad argument");
        return B200_ERR_INVALID;
    }
    if ((uint64_t)hi < 0xFFFFFFFFull && hi > 0) {   // hot case: hoist the bucket computation
        const uint32_t r32 = (uint32_t)hi;
        uint32_t bucket = 0xFFFFFFFFu / (r32 + 1u);
        if (0xFFFFFFFFu % (r32 + 1u) == r32) ++bucket;
        for (int64_t t = 0; t < n; ++t) {
            uint32_t v;
            do { v = s->next() / bucket; } while (v > r32);
            out[t] = (int64_t)v;
        }
        return B200_OK;
    }
    for (int64_t t = 0; t < n; ++t) out[t] = (int64_t)s->draw((uint64_t)hi);
    return B200_OK;
}

extern "C" int b200_mt_sampler_fill_i32(b200_mt_sampler* s, int64_t hi, int64_t n, int32_t* out)
{
    if (!s || hi < 0 || hi > 0x7fffffffLL || n < 0 || (n > 0 && !out)) {
        b200::set_error("b200_mt_sampler_fill_i32: bad argument");
        return B200_ERR_INVALID;
    }
    for (int64_t t = 0; t < n; ++t) out[t] = (int32_t)s->draw((uint64_t)hi);
    return B200_OK;
}

// ---------------------------------------------------------------------------------------
// Seeded sample streams of the BPR siblings whose draws depend on the data (so they cannot be produced by three
// independent fill calls): the streams are drawn here, sample by sample, in the reference's order.

// VEBPR._fit_sgd_viewloss (cornac/models/bpr/recom_vebpr.pyx:240-280): per sample  i_index <- pos % nnz;  then, ONLY when
// the user has viewed items,  v <- view_indices[view_indptr[u] + view_draw % num_view]  (else v = -1);  then  j <- neg.
extern "C" int b200_vebpr_draw_host(b200_mt_sampler* pos, b200_mt_sampler* view, b200_mt_sampler* neg, int64_t nnz, int64_t n_items,
                                    const int32_t* coo_row, const int32_t* view_indptr, const int32_t* view_indices,
                                    int64_t n_samples, int64_t* i_index_out, int32_t* v_id_out, int32_t* j_id_out)
{
    if (!pos || !view || !neg || nnz < 1 || n_items < 1 || n_items > 0x7fffffffLL || n_samples < 0 || !coo_row || !view_indptr ||
        (n_samples > 0 && (!i_index_out || !v_id_out || !j_id_out))) {
        b200::set_error("b200_vebpr_draw_host: bad argument");
        return B200_ERR_INVALID;
    }
    for (int64_t s = 0; s < n_samples; ++s) {
        const int64_t ii = (int64_t)(pos->draw((uint64_t)(nnz - 1)) % (uint64_t)nnz);
        const int32_t u = coo_row[ii];
        const int32_t num_view = view_indptr[u + 1] - view_indptr[u];
        int32_t v = -1;
        if (num_view > 0) v = view_indices[view_indptr[u] + (int64_t)(view->draw((uint64_t)(n_items - 1)) % (uint64_t)num_view)];
        i_index_out[s] = ii;
        v_id_out[s] = v;
        j_id_out[s] = (int32_t)neg->draw((uint64_t)(n_items - 1));
    }
    return B200_OK;
}

// SBPR._fit_sgd (cornac/models/sbpr/recom_sbpr.pyx:225-232): per sample  i_index <- pos;  j <- neg;  k_rand <- neg (second draw
// of the same stream) / num_items;  k_index = social_indptr[u] + (int)floor(k_rand * n_social).  k_rand is the float
// product draw * (1 / num_items): the reference's build (-O3 -ffast-math) hoists the reciprocal out of the loop, and only
// this form reproduces the compiled reference's samples (fixture tests/golden/sbpr_mid_k16.npz; the division does not).
extern "C" int b200_sbpr_draw_host(b200_mt_sampler* pos, b200_mt_sampler* neg, int64_t nnz, int64_t n_items,
                                   const int32_t* coo_row, const int32_t* social_indptr, int64_t n_samples,
                                   int64_t* i_index_out, int32_t* j_id_out, int64_t* k_index_out)
{
    if (!pos || !neg || nnz < 1 || n_items < 1 || n_items > 0x7fffffffLL || n_samples < 0 || !coo_row || !social_indptr ||
        (n_samples > 0 && (!i_index_out || !j_id_out || !k_index_out))) {
        b200::set_error("b200_sbpr_draw_host: bad argument");
        return B200_ERR_INVALID;
    }
    volatile float inv_items = 1.0f / (float)n_items;
    for (int64_t s = 0; s < n_samples; ++s) {
        const int64_t ii = (int64_t)pos->draw((uint64_t)(nnz - 1));
        const int32_t u = coo_row[ii];
        const int32_t j = (int32_t)neg->draw((uint64_t)(n_items - 1));
        const float k_rand = (float)(long)neg->draw((uint64_t)(n_items - 1)) * inv_items;
        const int32_t n_social = social_indptr[u + 1] - social_indptr[u];
        i_index_out[s] = ii;
        j_id_out[s] = j;
        k_index_out[s] = (int64_t)social_indptr[u] + (int)__builtin_floor((double)(k_rand * (float)n_social));
    }
    return B200_OK;
}
