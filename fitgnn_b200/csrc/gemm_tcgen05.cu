// tcgen05 / TMA dense transform (FITGNN_GEMM_BF16X3) — placeholder until the tensor-core kernel lands.
#include "common.cuh"
namespace fitgnn {
int gemm_bf16x3(const void*, const void*, int64_t, const void*, const void*, int64_t, const float*, int64_t, int, int,
                int, int, float*, int64_t, cudaStream_t) {
  set_error("gemm: FITGNN_GEMM_BF16X3 is not built yet");
  return FITGNN_EUNSUP;
}
}  // namespace fitgnn
