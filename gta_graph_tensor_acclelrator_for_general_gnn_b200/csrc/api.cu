// Error string, ABI version and launch counter of libgta_b200.so.
#include <atomic>
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

}  // namespace gta

extern "C" {

const char* gta_last_error(void) { return gta::g_error; }
int gta_abi_version(void) { return 1; }
int64_t gta_launch_count(void) { return gta::g_launches.load(); }
void gta_launch_count_reset(void) { gta::g_launches.store(0); }

}  // extern "C"
