// Shared helpers for the gta_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/gta_b200.h"

namespace gta {

void set_error(const char* fmt, ...);
void count_launch(int n = 1);

inline int check_cuda(cudaError_t e, const char* what) {
  if (e != cudaSuccess) {
    set_error("%s: %s", what, cudaGetErrorString(e));
    return GTA_ERR_CUDA;
  }
  return GTA_OK;
}

#define GTA_CHECK_LAUNCH(what)                                  \
  do {                                                          \
    gta::count_launch();                                        \
    int _rc = gta::check_cuda(cudaGetLastError(), what);        \
    if (_rc != GTA_OK) return _rc;                              \
  } while (0)

#define GTA_CUDA(call)                                          \
  do {                                                          \
    int _rc = gta::check_cuda((call), #call);                   \
    if (_rc != GTA_OK) return _rc;                              \
  } while (0)

#define GTA_REQUIRE(cond, ...)                                  \
  do {                                                          \
    if (!(cond)) {                                              \
      gta::set_error(__VA_ARGS__);                              \
      return GTA_ERR_INVALID;                                   \
    }                                                           \
  } while (0)

constexpr int kNumSMs = 148;  // B200

inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

__host__ __device__ inline int64_t ceil_div64(int64_t a, int64_t b) { return (a + b - 1) / b; }

// ---- cache-hinted loads ------------------------------------------------------------
// L2 eviction priority travels in a 64-bit cache-policy register (createpolicy); the direct
// .L2::evict_* qualifier only exists for 256-bit loads on sm_100.
__device__ __forceinline__ uint64_t policy_evict_first() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ uint64_t policy_evict_last() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
  return p;
}
// The same 64-bit policy words, fetched once per process (a one-thread kernel runs createpolicy) and then
// handed to the gather kernels as KERNEL PARAMETERS: a parameter lives in the constant bank, so the
// cache-hint descriptor of every load sits in a uniform register (a per-thread createpolicy result costs
// two R2UR per load; profiles/r01_gat_aggregate_v2: 16 of 85 instructions per 8 edges).
struct CachePolicies {
  uint64_t stream;   // L2 evict_first: CSR indices, edge weights
  uint64_t keep;     // L2 evict_last: gathered source rows
};
int cache_policies(CachePolicies* out);

// streamed once (CSR indices, edge weights): do not pollute L1, evict first from L2 so the
// gathered feature table stays resident in the 126 MB L2.
__device__ __forceinline__ int ld_stream_i32(const int* p, uint64_t pol) {
  int v;
  asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.s32 %0, [%1], %2;" : "=r"(v) : "l"(p), "l"(pol));
  return v;
}
__device__ __forceinline__ float ld_stream_f32(const float* p, uint64_t pol) {
  float v;
  asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.f32 %0, [%1], %2;" : "=f"(v) : "l"(p), "l"(pol));
  return v;
}
// gathered feature rows: read-only path, keep in L2 (evict_last policy), skip L1 allocation
// (random 512 B rows have no L1 reuse and would thrash it).
__device__ __forceinline__ float4 ld_gather_f32x4(const float* p, uint64_t pol) {
  float4 v;
  asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v4.f32 {%0,%1,%2,%3}, [%4], %5;"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p), "l"(pol));
  return v;
}
// output rows: written once, never re-read by this kernel
__device__ __forceinline__ void st_stream_f32x4(float* p, float4 v) {
  asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};"
               :: "l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

__device__ __forceinline__ float elu1(float x) { return x > 0.f ? x : expm1f(x); }
__device__ __forceinline__ float leaky(float x, float slope) { return x > 0.f ? x : x * slope; }

__device__ __forceinline__ float apply_epilogue(float x, int epi) {
  if (epi == GTA_EPI_ELU) return elu1(x);
  if (epi == GTA_EPI_RELU) return fmaxf(x, 0.f);
  return x;
}

}  // namespace gta
