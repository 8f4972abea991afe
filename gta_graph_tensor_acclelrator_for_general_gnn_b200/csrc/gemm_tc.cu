// tcgen05 3xTF32 GEMM (placeholder until the kernel lands): reports "unsupported" so
// gta_gemm_f32 takes the FFMA path.
#include "common.cuh"

namespace gta {
int gemm_tc_launch(const float*, int64_t, const float*, int64_t, float*, int64_t, int64_t, int, int, const float*,
                   const float*, int, float*, float*, cudaStream_t) {
  return GTA_ERR_UNSUPPORTED;
}
}  // namespace gta
