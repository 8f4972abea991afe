// Peer-to-peer plumbing of the in-kernel source exchange (see gta_exchange_t in include/gta_b200.h).
//
// The per-layer exchange of a destination-partitioned run is an all-gather of [Z | er] (SURVEY.md
// section 8e).  As an NCCL collective it is a launch of its own between the GEMM and the aggregation
// and it occupies SMs: on 8 B200 it was 0.21 ms of a 1.0 ms layer, un-overlapped, and the round-1
// attempts to overlap it (chunked NCCL all-gathers; copy-engine pulls gated by events) both lost to
// launch boundaries and SM contention.  Here the transfer happens INSIDE the aggregation launch
// (aggregate.cu: a few CTAs pull the peers' slots with NVLink loads while the others already reduce the
// rank's own column block); this file holds what is left outside it:
//
//   gta_ipc_alloc / gta_ipc_free     tables and signal blocks (cudaMalloc'ed so the IPC handle names a base pointer)
//   gta_ipc_export / gta_ipc_open    64-byte handle out / peer mapping in (lazy peer access)
//   gta_exchange_publish             "my slot is written": er range (from gta_er_stats) + step number into every
//                                    peer's signal block
#include <string.h>

#include "common.cuh"
#include "exchange.cuh"

namespace gta {

// One small CTA.  `stats` = this rank's er range per head as gta_er_stats writes it for one column block
// ([max codes (heads) | max -er codes (heads)], NULL: none); it goes, with the step number, to the signal
// block of every rank q at the slot index q uses for this rank.
__global__ void __launch_bounds__(128)
exchange_publish_kernel(const uint32_t* __restrict__ stats, int heads, int rank, int world, int step,
                        SignalPointers peers) {
  const int t = threadIdx.x;
  const int parity = step & 1;
  for (int i = t; i < world * 2 * heads; i += blockDim.x) {
    const int q = i / (2 * heads), j = i % (2 * heads);
    const int slot = rank >= q ? rank - q : rank - q + world;    // where q keeps this rank
    peers.sig[q]->stats[parity][slot][j] = stats[j];
  }
  __threadfence_system();
  __syncthreads();
  if (t < world) {
    const int slot = rank >= t ? rank - t : rank - t + world;
    asm volatile("st.release.sys.global.s32 [%0], %1;" ::"l"(&peers.sig[t]->ready[slot]), "r"(step) : "memory");
  }
}

}  // namespace gta

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

size_t gta_exchange_signal_bytes(void) { return sizeof(ExchangeSignals); }

int gta_exchange_publish(const uint32_t* stats, int32_t heads, int32_t rank, int32_t world, int32_t step,
                         void* const* h_peer_signals, void* stream_) {
  cudaStream_t st = static_cast<cudaStream_t>(stream_);
  GTA_REQUIRE(h_peer_signals && world >= 1 && world <= GTA_MAX_RANKS && rank >= 0 && rank < world && step >= 1,
              "gta_exchange_publish: bad arguments (world %d, rank %d, step %d)", world, rank, step);
  GTA_REQUIRE(heads >= 0 && heads <= 32 && (heads == 0 || stats), "gta_exchange_publish: %d heads without statistics", heads);
  SignalPointers peers{};
  for (int q = 0; q < world; ++q) {
    GTA_REQUIRE(h_peer_signals[q], "gta_exchange_publish: signal block of rank %d is not mapped", q);
    peers.sig[q] = static_cast<ExchangeSignals*>(h_peer_signals[q]);
  }
  exchange_publish_kernel<<<1, 128, 0, st>>>(stats, heads, rank, world, step, peers);
  GTA_CHECK_LAUNCH("exchange_publish_kernel");
  return GTA_OK;
}

}  // extern "C"
