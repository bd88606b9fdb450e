// Shared helpers for the sm_100a kernels of the BPR / MF / score+rank hot path.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/b200cornac.h"

namespace b200 {

// ---------------------------------------------------------------------------
// error plumbing: every C-ABI entry returns 0 or a non-zero status; the text
// is kept per thread and returned by b200_last_error().
void set_error(const char* fmt, ...);
int cuda_fail(cudaError_t e, const char* what, const char* file, int line);

#define B200_CUDA(call)                                                        \
    do {                                                                       \
        cudaError_t _e = (call);                                               \
        if (_e != cudaSuccess) return ::b200::cuda_fail(_e, #call, __FILE__, __LINE__); \
    } while (0)

#define B200_REQUIRE(cond, ...)                                                \
    do {                                                                       \
        if (!(cond)) {                                                         \
            ::b200::set_error(__VA_ARGS__);                                    \
            return B200_ERR_INVALID;                                           \
        }                                                                      \
    } while (0)

int sm_count();   // cached multiProcessorCount of the current device
void count_launch(int n = 1);   // one tick per kernel this library launches (b200_kernel_launches)

// ---------------------------------------------------------------------------
// Philox4x32-10 counter-based RNG (Salmon et al., SC'11).  Stateless: the
// (u,i,j) triplet of sample s in epoch e is a pure function of (seed, e, s),
// so any grid shape / shard layout draws the same stream.
struct Philox4 {
    uint32_t x, y, z, w;
};

__host__ __device__ __forceinline__ uint32_t mulhi32(uint32_t a, uint32_t b)
{
#ifdef __CUDA_ARCH__
    return __umulhi(a, b);
#else
    return (uint32_t)(((uint64_t)a * b) >> 32);
#endif
}

__host__ __device__ __forceinline__ Philox4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                                          uint32_t k0, uint32_t k1)
{
    const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        uint32_t hi0 = mulhi32(M0, c0), lo0 = M0 * c0;
        uint32_t hi1 = mulhi32(M1, c2), lo1 = M1 * c2;
        uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += W0; k1 += W1;
    }
    return Philox4{c0, c1, c2, c3};
}

// unbiased-enough range reduction: floor(r64 * n / 2^64), bias <= n / 2^64
__host__ __device__ __forceinline__ uint64_t range64(uint32_t lo, uint32_t hi, uint64_t n)
{
    uint64_t r = ((uint64_t)hi << 32) | lo;
#ifdef __CUDA_ARCH__
    return __umul64hi(r, n);
#else
    return (uint64_t)(((unsigned __int128)r * n) >> 64);
#endif
}

// ---------------------------------------------------------------------------
// sub-warp (G-lane group) sum; G is a power of two <= 32, groups are aligned.  The shuffle
// mask names only the group's own lanes, so groups of one warp may diverge (different trip
// counts, skipped samples) without dead-locking each other.
template <int G>
__device__ __forceinline__ unsigned group_mask()
{
    if constexpr (G >= 32) {
        return 0xffffffffu;
    } else {
        const unsigned lane = threadIdx.x & 31u;
        return ((1u << G) - 1u) << (lane & ~(unsigned)(G - 1));
    }
}
template <int G>
__device__ __forceinline__ float group_sum(float v)
{
    const unsigned m = group_mask<G>();
#pragma unroll
    for (int o = G / 2; o > 0; o >>= 1) v += __shfl_xor_sync(m, v, o);
    return v;
}

// has_non_zero: lower_bound over the sorted CSR row [lo, hi)
__device__ __forceinline__ bool row_contains(const int32_t* __restrict__ indices, int64_t lo, int64_t hi, int32_t col)
{
    const int64_t end = hi;
    while (lo < hi) {
        int64_t mid = lo + ((hi - lo) >> 1);
        int32_t v = __ldg(indices + mid);
        if (v < col) lo = mid + 1; else hi = mid;
    }
    return lo < end && __ldg(indices + lo) == col;
}

// order-preserving float -> uint32 key (larger float => larger key; -0.0 maps to +0.0)
__host__ __device__ __forceinline__ uint32_t float_key(float f)
{
#ifdef __CUDA_ARCH__
    uint32_t b = __float_as_uint(f);
#else
    union { float f; uint32_t u; } cv; cv.f = f; uint32_t b = cv.u;
#endif
    if (b == 0x80000000u) b = 0u;   // -0.0 ties with +0.0, as in a float compare
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}

}  // namespace b200
