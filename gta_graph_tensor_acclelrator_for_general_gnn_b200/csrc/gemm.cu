// gta_gemm_f32: COMP_MM applynode entry point.  Chooses the tcgen05 3xTF32 kernel
// (gemm_tc.cu) when its tile rules hold, else the FFMA kernel (gemm_simt.cu).
#include <stdlib.h>

#include "common.cuh"

namespace gta {
int gemm_simt_launch(const float* x, int64_t ldx, const float* w, int64_t ldw, float* z, int64_t ldz, int64_t num_rows,
                     int k, int f, cudaStream_t st);
int attn_project_launch(const float* z, int64_t ldz, int64_t num_rows, int f, const float* al, const float* ar,
                        int heads, float* el, float* er, int64_t lder, cudaStream_t st);
// returns GTA_ERR_UNSUPPORTED when the shape is outside the tensor-core kernel's rules
int gemm_tc_launch(const float* x, int64_t ldx, const float* w, int64_t ldw, float* z, int64_t ldz, int64_t num_rows,
                   int k, int f, const float* al, const float* ar, int heads, float* el, float* er, int64_t lder,
                   void* workspace, size_t workspace_bytes, cudaStream_t st, int z_bf16 = 0);
size_t gemm_tc_workspace(int k, int f);
}  // namespace gta

using namespace gta;

extern "C" {

// 0 = auto (tensor cores when eligible), 1 = force FFMA, 2 = force tcgen05 (error if ineligible)
static int g_gemm_mode = 0;
int gta_gemm_set_mode(int mode) {
  GTA_REQUIRE(mode >= 0 && mode <= 2, "gta_gemm_set_mode: mode %d is not 0 (auto), 1 (simt) or 2 (tc)", mode);
  g_gemm_mode = mode;
  return GTA_OK;
}
int gta_gemm_get_mode(void) { return g_gemm_mode; }

size_t gta_gemm_workspace(int32_t k, int32_t f) { return (k > 0 && f > 0) ? gemm_tc_workspace(k, f) : 0; }

int gta_gemm_f32(const float* x, int64_t ldx, const float* w, int64_t ldw, float* z, int64_t ldz, int64_t num_rows,
                 int32_t k, int32_t f, const float* al, const float* ar, int32_t heads, float* el, float* er,
                 int64_t lder, void* workspace, size_t workspace_bytes, void* stream_) {
  cudaStream_t st = static_cast<cudaStream_t>(stream_);
  if (num_rows == 0) return GTA_OK;
  GTA_REQUIRE(x && w && z, "gta_gemm_f32: null pointer");
  GTA_REQUIRE(num_rows > 0 && k > 0 && f > 0, "gta_gemm_f32: non-positive shape");
  GTA_REQUIRE(ldx >= k && ldw >= f && ldz >= f, "gta_gemm_f32: leading dimension smaller than the row");
  const bool want_attn = (el && al) || (er && ar);
  GTA_REQUIRE(!want_attn || heads >= 1, "gta_gemm_f32: heads must be >= 1 when el/er are requested");
  if (lder <= 0) lder = heads;
  GTA_REQUIRE(!want_attn || lder >= heads, "gta_gemm_f32: er row stride %lld < heads %d", (long long)lder, heads);
  if (g_gemm_mode != 1) {
    int rc = gemm_tc_launch(x, ldx, w, ldw, z, ldz, num_rows, k, f, al, ar, heads, el, er, lder, workspace, workspace_bytes, st);
    if (rc == GTA_OK) return GTA_OK;
    if (rc != GTA_ERR_UNSUPPORTED) return rc;
    if (g_gemm_mode == 2) {
      set_error("gta_gemm_f32: shape k=%d f=%d ldx=%lld (or a missing workspace) is outside the tcgen05 kernel's rules",
                k, f, (long long)ldx);
      return rc;
    }
  }
  int rc = gemm_simt_launch(x, ldx, w, ldw, z, ldz, num_rows, k, f, st);
  if (rc != GTA_OK) return rc;
  if (want_attn) return attn_project_launch(z, ldz, num_rows, f, al, ar, heads, el, er, lder, st);
  return GTA_OK;
}

// bf16 storage mode: the same fp32-accurate product, Z rounded to bf16 once in the
// epilogue of the tcgen05 kernel (ldz counts bf16 elements, a multiple of 8); el / er from the fp32 accumulators.
int gta_gemm_f32_zbf16(const float* x, int64_t ldx, const float* w, int64_t ldw, void* z, int64_t ldz, int64_t num_rows,
                       int32_t k, int32_t f, const float* al, const float* ar, int32_t heads, float* el, float* er,
                       int64_t lder, void* workspace, size_t workspace_bytes, void* stream_) {
  cudaStream_t st = static_cast<cudaStream_t>(stream_);
  if (num_rows == 0) return GTA_OK;
  GTA_REQUIRE(x && w && z, "gta_gemm_f32_zbf16: null pointer");
  GTA_REQUIRE(num_rows > 0 && k > 0 && f > 0, "gta_gemm_f32_zbf16: non-positive shape");
  GTA_REQUIRE(ldx >= k && ldw >= f && ldz >= f, "gta_gemm_f32_zbf16: leading dimension smaller than the row");
  const bool want_attn = (el && al) || (er && ar);
  GTA_REQUIRE(!want_attn || heads >= 1, "gta_gemm_f32_zbf16: heads must be >= 1 when el/er are requested");
  if (lder <= 0) lder = heads;
  int rc = gemm_tc_launch(x, ldx, w, ldw, static_cast<float*>(z), ldz, num_rows, k, f, al, ar, heads, el, er, lder, workspace,
                          workspace_bytes, st, 1);
  if (rc == GTA_ERR_UNSUPPORTED)
    set_error("gta_gemm_f32_zbf16: shape k=%d f=%d ldx=%lld ldz=%lld (or a missing workspace) is outside the tcgen05 kernel's "
              "rules; the bf16 storage mode has no FFMA fallback", k, f, (long long)ldx, (long long)ldz);
  return rc;
}

}  // extern "C"
