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
