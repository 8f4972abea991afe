// COMP_MM applynode on the FP32 pipe (FFMA): the IEEE-fp32 fallback of gta_gemm_f32, used
// when the tcgen05 path does not apply (K or F outside its tile rules) and as the in-library
// cross-check of the tensor-core kernel.  Z[N,F] = X[N,K] . W[K,F], row-major.
//
// CTA tile 128 x BN x 16, 256 threads as 16 x 16, 8 x (BN/16) outputs per thread, register
// prefetch of the next k-slab while the current one is multiplied out of shared memory.
#include "common.cuh"

namespace gta {

constexpr int kBM = 128;
constexpr int kBK = 16;

template <int BN>
__global__ void __launch_bounds__(256)
gemm_simt_kernel(const float* __restrict__ x, int64_t ldx, const float* __restrict__ w, int64_t ldw,
                 float* __restrict__ z, int64_t ldz, int64_t num_rows, int k_dim, int f, int vec_ok) {
  constexpr int TN = BN / 16;
  __shared__ __align__(16) float As[kBK][kBM + 4];
  __shared__ __align__(16) float Bs[kBK][BN];
  const int tid = threadIdx.x;
  const int tx = tid & 15;    // column group
  const int ty = tid >> 4;    // row group
  const int64_t row0 = int64_t(blockIdx.x) * kBM;
  const int col0 = blockIdx.y * BN;

  float acc[8][TN];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

  // A loader: 128 rows x 16 k = 512 float4; thread t -> rows (t>>2) and (t>>2)+64, k offset 4*(t&3)
  const int a_row = tid >> 2;
  const int a_k = (tid & 3) * 4;
  float4 a_reg[2];
  // B loader: 16 x BN scalars, BN*16/256 per thread
  constexpr int B_PER = BN * kBK / 256;
  float b_reg[B_PER];

  auto load_a = [&](int k0) {
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      int64_t r = row0 + a_row + 64 * h;
      int kk = k0 + a_k;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (r < num_rows && kk < k_dim) {
        const float* p = x + r * ldx + kk;
        if (vec_ok && kk + 3 < k_dim) {
          v = *reinterpret_cast<const float4*>(p);
        } else {
          v.x = p[0];
          if (kk + 1 < k_dim) v.y = p[1];
          if (kk + 2 < k_dim) v.z = p[2];
          if (kk + 3 < k_dim) v.w = p[3];
        }
      }
      a_reg[h] = v;
    }
  };
  auto load_b = [&](int k0) {
#pragma unroll
    for (int q = 0; q < B_PER; ++q) {
      int e = tid + 256 * q;
      int kk = k0 + e / BN;
      int c = col0 + e % BN;
      b_reg[q] = (kk < k_dim && c < f) ? __ldg(w + int64_t(kk) * ldw + c) : 0.f;
    }
  };
  auto store_smem = [&]() {
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      int r = a_row + 64 * h;
      As[a_k + 0][r] = a_reg[h].x;
      As[a_k + 1][r] = a_reg[h].y;
      As[a_k + 2][r] = a_reg[h].z;
      As[a_k + 3][r] = a_reg[h].w;
    }
#pragma unroll
    for (int q = 0; q < B_PER; ++q) {
      int e = tid + 256 * q;
      Bs[e / BN][e % BN] = b_reg[q];
    }
  };

  load_a(0);
  load_b(0);
  for (int k0 = 0; k0 < k_dim; k0 += kBK) {
    store_smem();
    __syncthreads();
    if (k0 + kBK < k_dim) {
      load_a(k0 + kBK);
      load_b(k0 + kBK);
    }
#pragma unroll
    for (int kk = 0; kk < kBK; ++kk) {
      float a[8], b[TN];
      float4 a0 = *reinterpret_cast<const float4*>(&As[kk][ty * 8]);
      float4 a1 = *reinterpret_cast<const float4*>(&As[kk][ty * 8 + 4]);
      a[0] = a0.x; a[1] = a0.y; a[2] = a0.z; a[3] = a0.w;
      a[4] = a1.x; a[5] = a1.y; a[6] = a1.z; a[7] = a1.w;
#pragma unroll
      for (int j = 0; j < TN; ++j) b[j] = Bs[kk][tx + 16 * j];
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    int64_t r = row0 + ty * 8 + i;
    if (r >= num_rows) continue;
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      int c = col0 + tx + 16 * j;
      if (c < f) z[r * ldz + c] = acc[i][j];
    }
  }
}

// GAT ops 1 and 2 (applynode MM with [F,H] weights): el = Z.Al, er = Z.Ar; warp per row.
__global__ void __launch_bounds__(256)
attn_project_kernel(const float* __restrict__ z, int64_t ldz, int64_t num_rows, int f, const float* __restrict__ al,
                    const float* __restrict__ ar, int heads, float* __restrict__ el, float* __restrict__ er,
                    int64_t lder) {
  const int lane = threadIdx.x & 31;
  const int64_t r = (blockIdx.x * int64_t(blockDim.x) + threadIdx.x) >> 5;
  if (r >= num_rows) return;
  for (int h = 0; h < heads; ++h) {
    float sl = 0.f, sr = 0.f;
    for (int c = lane; c < f; c += 32) {
      float v = z[r * ldz + c];
      if (al) sl = fmaf(v, __ldg(al + int64_t(c) * heads + h), sl);
      if (ar) sr = fmaf(v, __ldg(ar + int64_t(c) * heads + h), sr);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      sl += __shfl_xor_sync(0xffffffffu, sl, o);
      sr += __shfl_xor_sync(0xffffffffu, sr, o);
    }
    if (lane == 0) {
      if (el) el[r * heads + h] = sl;
      if (er) er[r * lder + h] = sr;
    }
  }
}

int gemm_simt_launch(const float* x, int64_t ldx, const float* w, int64_t ldw, float* z, int64_t ldz, int64_t num_rows,
                     int k, int f, cudaStream_t st) {
  int vec_ok = (ldx % 4 == 0) && ((reinterpret_cast<uintptr_t>(x) & 15) == 0);
  unsigned gx = (unsigned)((num_rows + kBM - 1) / kBM);
  if (f > 64) {
    gemm_simt_kernel<128><<<dim3(gx, (f + 127) / 128), 256, 0, st>>>(x, ldx, w, ldw, z, ldz, num_rows, k, f, vec_ok);
  } else if (f > 32) {
    gemm_simt_kernel<64><<<dim3(gx, 1), 256, 0, st>>>(x, ldx, w, ldw, z, ldz, num_rows, k, f, vec_ok);
  } else if (f > 16) {
    gemm_simt_kernel<32><<<dim3(gx, 1), 256, 0, st>>>(x, ldx, w, ldw, z, ldz, num_rows, k, f, vec_ok);
  } else {
    gemm_simt_kernel<16><<<dim3(gx, 1), 256, 0, st>>>(x, ldx, w, ldw, z, ldz, num_rows, k, f, vec_ok);
  }
  GTA_CHECK_LAUNCH("gemm_simt_kernel");
  return GTA_OK;
}

int attn_project_launch(const float* z, int64_t ldz, int64_t num_rows, int f, const float* al, const float* ar,
                        int heads, float* el, float* er, int64_t lder, cudaStream_t st) {
  attn_project_kernel<<<(unsigned)((num_rows * 32 + 255) / 256), 256, 0, st>>>(z, ldz, num_rows, f, al, ar, heads, el, er, lder);
  GTA_CHECK_LAUNCH("attn_project_kernel");
  return GTA_OK;
}

}  // namespace gta
