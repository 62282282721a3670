// zs_match_l2.cu -- L2 top-2 for 128-dim integer-valued descriptors (cv::SIFT) -- placeholder that routes to
// the exact CUDA-core dp4a kernel until the tcgen05 kernel lands.
#include "zs_common.cuh"

zs_status zs_l2_cuda_core_top2(zs_context* ctx, const uint8_t* q8, const int* nq, const uint8_t* t8, const int* nt, int pairs,
                               int cap_q, int cap_t, int dim, int* idx, int* dist);

zs_status zs_l2_tensor_top2(zs_context* ctx, const uint8_t* q8, const int* nq, const uint8_t* t8, const int* nt, int pairs,
                            int cap_q, int cap_t, int dim, int* idx, int* dist)
{
    return zs_l2_cuda_core_top2(ctx, q8, nq, t8, nt, pairs, cap_q, cap_t, dim, idx, dist);
}
