// Tensor-core (tcgen05 + TMA) candidate pass of the fused rank kernel.  Not enabled yet:
// rank_tc_supported() returns 0 so b200_rank_topk takes the exact path in rank.cu.
#include "common.cuh"

namespace b200 {

int rank_tc_supported(int64_t, int64_t, int, int) { return 0; }
int64_t rank_tc_workspace_bytes(int64_t, int64_t, int, int) { return 0; }
int rank_tc(const float*, const int64_t*, int64_t, const float*, int64_t, int, const float*, const float*,
            const int64_t*, const int32_t*, int, int32_t*, float*, void*, int64_t, cudaStream_t)
{
    set_error("rank_tc: tensor-core path not built");
    return B200_ERR_UNSUPPORTED;
}

}  // namespace b200
