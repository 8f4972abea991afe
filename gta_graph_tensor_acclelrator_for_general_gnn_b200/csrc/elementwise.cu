// Generic single-instruction kernels so that ANY legal plan executes, not only the fused
// patterns: COMP_ADD / COMP_MUL / COMP_SF on edges (interpreter.py:85-106 applyedge tiles)
// with FETCH-eliminated scatters as virtual operands (interpreter.py:768-806), and the
// applynode elementwise ops.  Pure streaming work: one warp per destination row, lanes over
// the feature columns of each edge, coalesced.
//
// Indexing: destination-side node operands by (row - row_begin), source-side node operands
// by indices[k], edge tensors by the CSR position k.
#include "common.cuh"

namespace gta {

__device__ __forceinline__ float fetch_operand(const float* __restrict__ a, int kind, int64_t ld, int64_t k,
                                               int64_t lrow, int src, int col) {
  int64_t r = kind == GTA_OPND_EDGE ? k : (kind == GTA_OPND_DST ? lrow : int64_t(src));
  return a[r * ld + col];
}

__device__ __forceinline__ float binary_op(int op, float a, float b) {
  if (op == GTA_BIN_ADD) return a + b;
  if (op == GTA_BIN_MUL) return a * b;
  return b != 0.f ? a / b : 0.f;   // rows without edges: sum 0 / sum 0 is pinned to 0 (oracle gat_layer)
}

__device__ __forceinline__ float unary_op(int op, float slope, float a) {
  if (op == GTA_UN_EXP_LEAKY_RELU) return expf(leaky(a, slope));
  if (op == GTA_UN_ELU) return elu1(a);
  if (op == GTA_UN_RELU) return fmaxf(a, 0.f);
  return a;
}

__global__ void __launch_bounds__(256)
edge_binary_kernel(const int64_t* __restrict__ indptr, const int32_t* __restrict__ indices, int64_t row_begin,
                   int64_t row_end, int op, const float* __restrict__ a, int kind_a, int wa, int64_t lda,
                   const float* __restrict__ b, int kind_b, int wb, int64_t ldb, float* __restrict__ out, int wo,
                   int64_t ldo) {
  const int lane = threadIdx.x & 31;
  const int64_t r = row_begin + ((blockIdx.x * int64_t(blockDim.x) + threadIdx.x) >> 5);
  if (r >= row_end) return;
  const int64_t kb = indptr[r], ke = indptr[r + 1];
  const int da = wo / wa, db = wo / wb;
  if (wo >= 32) {
    for (int64_t k = kb; k < ke; ++k) {
      int src = indices[k];
      for (int c = lane; c < wo; c += 32) {
        float va = fetch_operand(a, kind_a, lda, k, r - row_begin, src, c / da);
        float vb = fetch_operand(b, kind_b, ldb, k, r - row_begin, src, c / db);
        out[k * ldo + c] = binary_op(op, va, vb);
      }
    }
  } else {
    // narrow tensors ([E,H]): flatten (edge, column) over the lanes
    const int64_t total = (ke - kb) * wo;
    for (int64_t t = lane; t < total; t += 32) {
      int64_t k = kb + t / wo;
      int c = int(t % wo);
      int src = indices[k];
      float va = fetch_operand(a, kind_a, lda, k, r - row_begin, src, c / da);
      float vb = fetch_operand(b, kind_b, ldb, k, r - row_begin, src, c / db);
      out[k * ldo + c] = binary_op(op, va, vb);
    }
  }
}

__global__ void __launch_bounds__(256)
edge_unary_kernel(const int64_t* __restrict__ indptr, const int32_t* __restrict__ indices, int64_t row_begin,
                  int64_t row_end, int op, float slope, const float* __restrict__ a, int kind_a, int wa, int64_t lda,
                  float* __restrict__ out, int64_t ldo) {
  const int lane = threadIdx.x & 31;
  const int64_t r = row_begin + ((blockIdx.x * int64_t(blockDim.x) + threadIdx.x) >> 5);
  if (r >= row_end) return;
  const int64_t kb = indptr[r], ke = indptr[r + 1];
  const int64_t total = (ke - kb) * wa;
  for (int64_t t = lane; t < total; t += 32) {
    int64_t k = kb + t / wa;
    int c = int(t % wa);
    float va = fetch_operand(a, kind_a, lda, k, r - row_begin, indices[k], c);
    out[k * ldo + c] = unary_op(op, slope, va);
  }
}

__global__ void __launch_bounds__(256)
node_binary_kernel(int op, const float* __restrict__ a, int wa, int64_t lda, const float* __restrict__ b, int wb,
                   int64_t ldb, float* __restrict__ out, int wo, int64_t ldo, int64_t num_rows) {
  int64_t t = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
  int64_t stride = int64_t(gridDim.x) * blockDim.x;
  const int da = wo / wa, db = wo / wb;
  for (; t < num_rows * wo; t += stride) {
    int64_t r = t / wo;
    int c = int(t % wo);
    out[r * ldo + c] = binary_op(op, a[r * lda + c / da], b[r * ldb + c / db]);
  }
}

__global__ void __launch_bounds__(256)
node_unary_kernel(int op, float slope, const float* __restrict__ a, int64_t lda, float* __restrict__ out, int64_t ldo,
                  int width, int64_t num_rows) {
  int64_t t = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
  int64_t stride = int64_t(gridDim.x) * blockDim.x;
  for (; t < num_rows * width; t += stride) {
    int64_t r = t / width;
    int c = int(t % width);
    out[r * ldo + c] = unary_op(op, slope, a[r * lda + c]);
  }
}

static unsigned stream_grid(int64_t n) {
  int64_t g = (n + 255) / 256;
  int64_t cap = int64_t(kNumSMs) * 32;
  if (g > cap) g = cap;
  return (unsigned)(g < 1 ? 1 : g);
}

}  // namespace gta

using namespace gta;

extern "C" {

int gta_edge_binary_f32(const int64_t* indptr, const int32_t* indices, int64_t row_begin, int64_t row_end, int32_t op,
                        const float* a, int32_t kind_a, int32_t wa, int64_t lda, const float* b, int32_t kind_b,
                        int32_t wb, int64_t ldb, float* out, int32_t wo, int64_t ldo, void* stream_) {
  cudaStream_t st = static_cast<cudaStream_t>(stream_);
  GTA_REQUIRE(indptr && indices && a && b && out, "gta_edge_binary_f32: null pointer");
  GTA_REQUIRE(op >= GTA_BIN_ADD && op <= GTA_BIN_DIV, "gta_edge_binary_f32: bad op %d", op);
  GTA_REQUIRE(kind_a >= 0 && kind_a <= 2 && kind_b >= 0 && kind_b <= 2, "gta_edge_binary_f32: bad operand kind");
  GTA_REQUIRE(wa >= 1 && wb >= 1 && wo >= 1 && wo % wa == 0 && wo % wb == 0, "gta_edge_binary_f32: operand widths %d,%d must divide %d", wa, wb, wo);
  int64_t rows = row_end - row_begin;
  if (rows <= 0) return GTA_OK;
  edge_binary_kernel<<<(unsigned)((rows * 32 + 255) / 256), 256, 0, st>>>(indptr, indices, row_begin, row_end, op, a,
                                                                          kind_a, wa, lda, b, kind_b, wb, ldb, out, wo, ldo);
  GTA_CHECK_LAUNCH("edge_binary_kernel");
  return GTA_OK;
}

int gta_edge_unary_f32(const int64_t* indptr, const int32_t* indices, int64_t row_begin, int64_t row_end, int32_t op,
                       float slope, const float* a, int32_t kind_a, int32_t wa, int64_t lda, float* out, int64_t ldo,
                       void* stream_) {
  cudaStream_t st = static_cast<cudaStream_t>(stream_);
  GTA_REQUIRE(indptr && indices && a && out, "gta_edge_unary_f32: null pointer");
  GTA_REQUIRE(op >= GTA_UN_EXP_LEAKY_RELU && op <= GTA_UN_COPY, "gta_edge_unary_f32: bad op %d", op);
  GTA_REQUIRE(kind_a >= 0 && kind_a <= 2 && wa >= 1, "gta_edge_unary_f32: bad operand");
  int64_t rows = row_end - row_begin;
  if (rows <= 0) return GTA_OK;
  edge_unary_kernel<<<(unsigned)((rows * 32 + 255) / 256), 256, 0, st>>>(indptr, indices, row_begin, row_end, op, slope,
                                                                         a, kind_a, wa, lda, out, ldo);
  GTA_CHECK_LAUNCH("edge_unary_kernel");
  return GTA_OK;
}

int gta_node_binary_f32(int32_t op, const float* a, int32_t wa, int64_t lda, const float* b, int32_t wb, int64_t ldb,
                        float* out, int32_t wo, int64_t ldo, int64_t num_rows, void* stream_) {
  cudaStream_t st = static_cast<cudaStream_t>(stream_);
  GTA_REQUIRE(a && b && out, "gta_node_binary_f32: null pointer");
  GTA_REQUIRE(op >= GTA_BIN_ADD && op <= GTA_BIN_DIV, "gta_node_binary_f32: bad op %d", op);
  GTA_REQUIRE(wa >= 1 && wb >= 1 && wo >= 1 && wo % wa == 0 && wo % wb == 0, "gta_node_binary_f32: operand widths %d,%d must divide %d", wa, wb, wo);
  if (num_rows <= 0) return GTA_OK;
  node_binary_kernel<<<stream_grid(num_rows * wo), 256, 0, st>>>(op, a, wa, lda, b, wb, ldb, out, wo, ldo, num_rows);
  GTA_CHECK_LAUNCH("node_binary_kernel");
  return GTA_OK;
}

int gta_node_unary_f32(int32_t op, float slope, const float* a, int64_t lda, float* out, int64_t ldo, int32_t width,
                       int64_t num_rows, void* stream_) {
  cudaStream_t st = static_cast<cudaStream_t>(stream_);
  GTA_REQUIRE(a && out, "gta_node_unary_f32: null pointer");
  GTA_REQUIRE(op >= GTA_UN_EXP_LEAKY_RELU && op <= GTA_UN_COPY && width >= 1, "gta_node_unary_f32: bad op/width");
  if (num_rows <= 0) return GTA_OK;
  node_unary_kernel<<<stream_grid(num_rows * width), 256, 0, st>>>(op, slope, a, lda, out, ldo, width, num_rows);
  GTA_CHECK_LAUNCH("node_unary_kernel");
  return GTA_OK;
}

}  // extern "C"
