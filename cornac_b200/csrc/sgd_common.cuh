// Row-fragment helpers shared by the BPR and MF SGD kernels.
//
// A sample is owned by a GROUP of G lanes (G = 4..32, power of two, aligned inside the
// warp).  A factor row of k floats is cut into "units" (a float4 when k % 4 == 0, else a
// float); lane `lg` of the group owns units lg, lg+G, ... (NPL of them), so one row
// access is a single fully coalesced request of 128-bit loads (k=128, G=32: 512 B).
//
// Factor rows are read with ld.global.cg / written with st.global.cg (L2 only): rows are
// concurrently rewritten by other SMs (Hogwild), and an L1-resident copy of a hot item row
// would stay stale for the whole persistent kernel.
#pragma once
#include "common.cuh"

namespace b200 {

template <int NPL, bool VEC>
struct RowFrag {
    static constexpr int W = VEC ? 4 : 1;
    float v[NPL * W];
};

template <int G, int NPL, bool VEC>
__device__ __forceinline__ void row_load(RowFrag<NPL, VEC>& r, const float* __restrict__ row, int lg, int n_units)
{
#pragma unroll
    for (int t = 0; t < NPL; ++t) {
        const int e = lg + t * G;
        if (VEC) {
            float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
            if (e < n_units) x = __ldcg(reinterpret_cast<const float4*>(row) + e);
            r.v[t * 4 + 0] = x.x; r.v[t * 4 + 1] = x.y; r.v[t * 4 + 2] = x.z; r.v[t * 4 + 3] = x.w;
        } else {
            r.v[t] = (e < n_units) ? __ldcg(row + e) : 0.f;
        }
    }
}

template <int G, int NPL, bool VEC>
__device__ __forceinline__ void row_store(const RowFrag<NPL, VEC>& r, float* __restrict__ row, int lg, int n_units)
{
#pragma unroll
    for (int t = 0; t < NPL; ++t) {
        const int e = lg + t * G;
        if (e < n_units) {
            if (VEC) {
                __stcg(reinterpret_cast<float4*>(row) + e,
                       make_float4(r.v[t * 4 + 0], r.v[t * 4 + 1], r.v[t * 4 + 2], r.v[t * 4 + 3]));
            } else {
                __stcg(row + e, r.v[t]);
            }
        }
    }
}

// relaxed, fire-and-forget atomic accumulate of a delta fragment (no lost updates)
__device__ __forceinline__ void red_add_v4(float* addr, float a, float b, float c, float d)
{
    asm volatile("red.relaxed.gpu.global.add.v4.f32 [%0], {%1, %2, %3, %4};"
                 :: "l"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
__device__ __forceinline__ void red_add_f32(float* addr, float a)
{
    asm volatile("red.relaxed.gpu.global.add.f32 [%0], %1;" :: "l"(addr), "f"(a) : "memory");
}

template <int G, int NPL, bool VEC>
__device__ __forceinline__ void row_red_add(const RowFrag<NPL, VEC>& d, float* __restrict__ row, int lg, int n_units)
{
#pragma unroll
    for (int t = 0; t < NPL; ++t) {
        const int e = lg + t * G;
        if (e < n_units) {
            if (VEC) red_add_v4(row + e * 4, d.v[t * 4 + 0], d.v[t * 4 + 1], d.v[t * 4 + 2], d.v[t * 4 + 3]);
            else red_add_f32(row + e, d.v[t]);
        }
    }
}

// choose the lane layout for a factor width k
struct RowLayout {
    bool vec;
    int g;       // lanes per sample
    int npl;     // units per lane
    int n_units;
};
inline RowLayout pick_layout(int k)
{
    RowLayout L;
    L.vec = (k % 4 == 0);
    L.n_units = L.vec ? k / 4 : k;
    int g = 4;
    while (g < 32 && g < L.n_units) g <<= 1;
    L.g = g;
    L.npl = (L.n_units + g - 1) / g;
    return L;
}

// Compile-time dispatch over (G, NPL, VEC); NPL in {1,2,4,8} covers k <= 1024 (vec) / k <= 256 (scalar).
#define B200_DISPATCH_LAYOUT(L, CALL)                                                     \
    do {                                                                                  \
        const int _npl = (L).npl <= 1 ? 1 : (L).npl <= 2 ? 2 : (L).npl <= 4 ? 4 : 8;     \
        if ((L).vec) {                                                                    \
            switch ((L).g * 16 + _npl) {                                                  \
                case 4 * 16 + 1: CALL(4, 1, true); break;                                 \
                case 8 * 16 + 1: CALL(8, 1, true); break;                                 \
                case 16 * 16 + 1: CALL(16, 1, true); break;                               \
                case 32 * 16 + 1: CALL(32, 1, true); break;                               \
                case 32 * 16 + 2: CALL(32, 2, true); break;                               \
                case 32 * 16 + 4: CALL(32, 4, true); break;                               \
                case 32 * 16 + 8: CALL(32, 8, true); break;                               \
                default: ::b200::set_error("unsupported factor width"); return B200_ERR_UNSUPPORTED; \
            }                                                                             \
        } else {                                                                          \
            switch ((L).g * 16 + _npl) {                                                  \
                case 4 * 16 + 1: CALL(4, 1, false); break;                                \
                case 8 * 16 + 1: CALL(8, 1, false); break;                                \
                case 16 * 16 + 1: CALL(16, 1, false); break;                              \
                case 32 * 16 + 1: CALL(32, 1, false); break;                              \
                case 32 * 16 + 2: CALL(32, 2, false); break;                              \
                case 32 * 16 + 4: CALL(32, 4, false); break;                              \
                case 32 * 16 + 8: CALL(32, 8, false); break;                              \
                default: ::b200::set_error("unsupported factor width"); return B200_ERR_UNSUPPORTED; \
            }                                                                             \
        }                                                                                 \
    } while (0)

}  // namespace b200
