// Peer-to-peer exchange of the source-side tables over NVLink WITHOUT streaming multiprocessors.
//
// The per-layer exchange of a destination-partitioned run is an all-gather of [Z | er].  As an NCCL
// collective it occupies SMs, so it cannot hide under the aggregation kernel (measured on 8 B200:
// chunked NCCL all-gathers overlapped with the gather kernel made the step SLOWER, 1.01 -> 1.18 ms).
// Here every rank publishes its slot buffer through CUDA IPC once at setup; each step it PULLS its
// peers' slots with plain device-to-device copies on a copy stream -- copy engines, NVLink, no SMs --
// chunk by chunk, so column block q of the work list can start while chunk q+1 is in flight.
//
//   gta_ipc_alloc / gta_ipc_free     slot buffers (cudaMalloc'ed so the IPC handle names a base pointer)
//   gta_ipc_export / gta_ipc_open    64-byte handle out / peer mapping in (lazy peer access)
//   gta_copy_many                    n asynchronous copies on one stream
#include <string.h>

#include "common.cuh"

using namespace gta;

extern "C" {

int gta_ipc_alloc(size_t bytes, void** ptr) {
  GTA_REQUIRE(ptr && bytes > 0, "gta_ipc_alloc: bad arguments");
  GTA_CUDA(cudaMalloc(ptr, bytes));
  GTA_CUDA(cudaMemset(*ptr, 0, bytes));
  return GTA_OK;
}

int gta_ipc_free(void* ptr) {
  if (ptr) GTA_CUDA(cudaFree(ptr));
  return GTA_OK;
}

int gta_ipc_export(void* ptr, uint8_t* handle64) {
  GTA_REQUIRE(ptr && handle64, "gta_ipc_export: null pointer");
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle is 64 bytes");
  cudaIpcMemHandle_t h;
  GTA_CUDA(cudaIpcGetMemHandle(&h, ptr));
  memcpy(handle64, &h, 64);
  return GTA_OK;
}

int gta_ipc_open(const uint8_t* handle64, void** mapped) {
  GTA_REQUIRE(handle64 && mapped, "gta_ipc_open: null pointer");
  cudaIpcMemHandle_t h;
  memcpy(&h, handle64, 64);
  GTA_CUDA(cudaIpcOpenMemHandle(mapped, h, cudaIpcMemLazyEnablePeerAccess));
  return GTA_OK;
}

int gta_ipc_close(void* mapped) {
  if (mapped) GTA_CUDA(cudaIpcCloseMemHandle(mapped));
  return GTA_OK;
}

// dst[i] <- src[i] (bytes[i]) for i < n, all on `stream`; pointers may be peer mappings
int gta_copy_many(void* const* dst, const void* const* src, const int64_t* bytes, int32_t n, void* stream_) {
  cudaStream_t st = static_cast<cudaStream_t>(stream_);
  GTA_REQUIRE(n >= 0 && (n == 0 || (dst && src && bytes)), "gta_copy_many: bad arguments");
  for (int32_t i = 0; i < n; ++i) {
    if (bytes[i] <= 0) continue;
    GTA_CUDA(cudaMemcpyAsync(dst[i], src[i], size_t(bytes[i]), cudaMemcpyDefault, st));
  }
  return GTA_OK;
}

}  // extern "C"
