// Error string, ABI version and launch counter of libgta_b200.so.
#include <atomic>
#include <mutex>
#include <stdarg.h>

#include "common.cuh"

namespace gta {

static thread_local char g_error[512] = "";
static std::atomic<int64_t> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_error, sizeof(g_error), fmt, ap);
  va_end(ap);
}

void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

__global__ void cache_policy_kernel(uint64_t* out) {
  out[0] = policy_evict_first();
  out[1] = policy_evict_last();
}

// createpolicy's encoding is opaque, so it is asked of the device once (first aggregation call of the
// process; synchronises, hence outside any stream capture) instead of being hard-coded.
int cache_policies(CachePolicies* out) {
  static std::mutex mu;
  static bool ready = false;
  static CachePolicies cached{};
  std::lock_guard<std::mutex> lock(mu);
  if (!ready) {
    uint64_t* d = nullptr;
    uint64_t h[2] = {0, 0};
    GTA_CUDA(cudaMalloc(&d, sizeof(h)));
    cache_policy_kernel<<<1, 1>>>(d);
    cudaError_t e = cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
    cudaFree(d);
    GTA_CUDA(e);
    cached.stream = h[0];
    cached.keep = h[1];
    ready = true;
  }
  *out = cached;
  return GTA_OK;
}

}  // namespace gta

extern "C" {

const char* gta_last_error(void) { return gta::g_error; }
int gta_abi_version(void) { return 1; }
int64_t gta_launch_count(void) { return gta::g_launches.load(); }
void gta_launch_count_reset(void) { gta::g_launches.store(0); }

}  // extern "C"
