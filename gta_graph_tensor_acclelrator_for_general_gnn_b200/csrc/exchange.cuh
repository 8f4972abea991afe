// Device side of the in-kernel source exchange (gta_exchange_t, include/gta_b200.h): the signal block
// peers write into, the pull loop run by the first CTAs of an aggregation launch, and the gate the
// work-list CTAs pass before they gather from a slot that comes from another GPU.
#pragma once
#include "common.cuh"

namespace gta {

// ---- ordered-int code of a float: unsigned order == float order, 0 = "nothing seen" ---------------
__device__ __forceinline__ uint32_t ordered_code(float f) {
  const uint32_t b = __float_as_uint(f);
  return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float ordered_decode(uint32_t c) {
  return __uint_as_float((c & 0x80000000u) ? (c & 0x7fffffffu) : ~c);
}

// One per rank, mapped by every peer.  ready[k] = last step the owner of slot k has published;
// stats[step & 1][k] = er range codes of slot k ([max er per head (32) | max -er per head (32)], the first
// `heads` of each half are used -- the layout gta_er_stats writes per column block when heads == 32,
// and what block_bound() in aggregate.cu reads with a row pitch of 2*heads words).
struct ExchangeSignals {
  int32_t ready[GTA_MAX_RANKS];
  uint32_t stats[2][GTA_MAX_RANKS][64];
};
struct SignalPointers {
  ExchangeSignals* sig[GTA_MAX_RANKS];
};

// What an aggregation kernel needs of a gta_exchange_t (passed by value: constant bank, uniform).
struct Exchange {
  int32_t world;              // 0: no exchange, the table is complete
  int32_t copy_ctas;
  int32_t step;
  uint32_t row_bytes;
  int64_t slot_rows;
  char* table;
  const ExchangeSignals* signals;
  int32_t* arrived;           // [world] copy CTAs done with slot k (zeroed before the launch)
  const char* peer[GTA_MAX_RANKS];
  int32_t valid_rows[GTA_MAX_RANKS];
};

// bounded spin on a counter another agent releases; a protocol bug traps instead of hanging the GPU
template <bool SYSTEM>
__device__ __forceinline__ void wait_at_least(const int32_t* p, int32_t want) {
  // relaxed polls, one acquire fence on success (an acquire load per iteration would invalidate L1 every time)
  int32_t v = 0;
#pragma unroll 1
  for (uint32_t spin = 0; spin < (1u << 27); ++spin) {
    if (SYSTEM) asm volatile("ld.relaxed.sys.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    else asm volatile("ld.relaxed.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    if (v >= want) {
      if (SYSTEM) asm volatile("fence.acq_rel.sys;" ::: "memory");
      else asm volatile("fence.acq_rel.gpu;" ::: "memory");
      return;
    }
    __nanosleep(SYSTEM ? 200 : 100);
  }
  __trap();
}

// Pull loop of copy CTA `cta` (of ex.copy_ctas): slot k = 1 .. world-1 in ring order -- every rank reads
// from a different peer at any moment -- each slot cut evenly over the copy CTAs.  16-byte peer loads, 8 in
// flight per thread (peer latency is about 2 us: 64 CTAs x 128 threads x 128 B keep about 1 MB in flight,
// enough for the ~770 GB/s of one NVLink direction), plain stores into the local table, then one release
// increment of arrived[k] per CTA.
__device__ __forceinline__ void exchange_pull(const Exchange& ex, int cta) {
  constexpr int kInFlight = 8;
  for (int k = 1; k < ex.world; ++k) {
    if (threadIdx.x == 0) wait_at_least<true>(&ex.signals->ready[k], ex.step);
    __syncthreads();
    const int64_t units = int64_t(ex.valid_rows[k]) * (ex.row_bytes / 16);
    const int64_t lo = units * cta / ex.copy_ctas, hi = units * (cta + 1) / ex.copy_ctas;
    const uint4* src = reinterpret_cast<const uint4*>(ex.peer[k]);
    uint4* dst = reinterpret_cast<uint4*>(ex.table + int64_t(k) * ex.slot_rows * ex.row_bytes);
    for (int64_t i = lo + threadIdx.x; i < hi; i += int64_t(blockDim.x) * kInFlight) {
      uint4 v[kInFlight];
#pragma unroll
      for (int u = 0; u < kInFlight; ++u) {
        const int64_t j = i + int64_t(u) * blockDim.x;
        if (j < hi)
          asm volatile("ld.relaxed.sys.global.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                       : "=r"(v[u].x), "=r"(v[u].y), "=r"(v[u].z), "=r"(v[u].w) : "l"(src + j) : "memory");
      }
#pragma unroll
      for (int u = 0; u < kInFlight; ++u) {
        const int64_t j = i + int64_t(u) * blockDim.x;
        if (j < hi) dst[j] = v[u];
      }
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) asm volatile("red.release.gpu.global.add.s32 [%0], 1;" ::"l"(ex.arrived + k) : "memory");
  }
}

// Gate of a work-list group: `last_src` is the item's LAST (= highest) source id.  The copy CTAs finish the
// slots in ring order, so once the slot of the last source has landed every earlier slot the item touches
// has too.  Slot 0 is the rank's own rows.  Called by ONE lane of the group, followed by a group-wide
// __syncwarp by the caller.
// `landed` (per calling lane, 0 at kernel start) is the highest slot this lane has already seen complete: slots land in
// order and the work list walks them in order, so all but P-1 of a warp's items return on the first compare -- no
// division, no L2 poll, no fence.
__device__ __forceinline__ void exchange_gate(const Exchange& ex, int32_t last_src, int32_t& landed) {
  if (int64_t(last_src) < int64_t(landed + 1) * ex.slot_rows) return;
  const int64_t k = int64_t(last_src) / ex.slot_rows;
  wait_at_least<false>(ex.arrived + k, ex.copy_ctas);
  landed = int32_t(k);
}

}  // namespace gta
