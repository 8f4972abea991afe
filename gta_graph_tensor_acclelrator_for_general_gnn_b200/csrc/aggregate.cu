// Segmented scatter-reduce kernels: the HBM/L2-bound heart of the hot path.
//
//   aggregate_kernel      COMP_MUL_COMP_ADD / COMP_ADD gather  (interpreter.py:575-638, 85-106)
//   gat_aggregate_kernel  GAT ops 3-13 in one pass, online softmax (genGraphOP.py:52-62)
//   gat_logits_kernel     GAT block [4,5,6,7,8] alone (STORE_E p, STORE_N S)
//
// Mapping.  A work item (<= chunk edges of one destination row inside one column block, see
// schedule.cu) is owned by a GROUP of LANES = min(F,128)/4 lanes; each lane owns 4 consecutive
// features, so one gathered source row is ONE 128-bit load per lane and a full 512 B row per 32
// lanes.  Wider rows (F > 128) are covered by blockIdx.y feature windows of 128.  Source ids (and
// scalar edge weights) are read once per group, coalesced and streaming (L1 no-allocate, L2
// evict_first), then handed round the group by shuffle / shared memory.  Gathered rows use the
// read-only path without L1 allocation (no reuse inside an SM); L2 residency comes from the
// column-block order of the work list.  kUnroll independent row loads are in flight per lane and a
// full batch runs without a single predicate.
//
// Rows that own several items (long rows cut at `chunk` edges, and every row when the table is walked
// in column blocks) form a CHAIN in slot order: an item reduces its own edges first, then waits for
// its predecessor's published state (a release/acquire flag per slot), folds it in, and either
// publishes the folded state or -- last item of the row -- normalises and writes the output.  The
// wait is at the END of an item and whoever holds the predecessor took it from the item counter earlier
// and is running, so it almost never spins and cannot deadlock; there is no separate merge launch and the
// last partial of every row is never written (round 1: a combine kernel re-read 507 MB in 0.24 ms).
//
// Determinism.  Every item is reduced by exactly one group in ascending source order and the chain
// is a left fold in slot order: a fixed-shape reduction, bitwise reproducible run to run, no atomics.
#include <cuda_bf16.h>
#include <string.h>

#include "common.cuh"
#include "exchange.cuh"

namespace gta {

#ifndef GTA_AGG_THREADS
#define GTA_AGG_THREADS 128
#endif
// Resident CTAs per SM (caps the registers) and row loads in flight per lane, per kernel family.  Measured on
// B200, Reddit shape, persistent launch (tools/agg_probe.py, gpurun_out/p2_probe.log):
//   weighted aggregate   6x8: 3.36 ms   8x8: 3.29   8x4: 3.29   10x4: 3.13   (48 registers, no spills)
//   GAT staged (H <= 4)  6x8: 3.89 ms   7x8: 3.88   8x8: 3.75   8x4: 3.49   10x4: 4.11 (spills)
// More resident warps beat a deeper unroll: the kernels wait on L2 latency (long scoreboard), not on issue.
#ifndef GTA_AGG_MINBLOCKS
#define GTA_AGG_MINBLOCKS 10
#endif
#ifndef GTA_AGG_UNROLL
#define GTA_AGG_UNROLL 4
#endif
#ifndef GTA_GAT_MINBLOCKS
#define GTA_GAT_MINBLOCKS 8
#endif
#ifndef GTA_GAT_UNROLL
#define GTA_GAT_UNROLL 4
#endif
#ifndef GTA_LLH_MINBLOCKS
#define GTA_LLH_MINBLOCKS 6
#endif
#ifndef GTA_LLH_UNROLL
#define GTA_LLH_UNROLL 8
#endif
#ifndef GTA_AGG_FASTEXP
#define GTA_AGG_FASTEXP 1
#endif
#ifndef GTA_GAT_FORCE_LLH
#define GTA_GAT_FORCE_LLH 0       // experiment: run the lane-local-head kernel for every head count
#endif
#ifndef GTA_ITEM_PREFETCH
#define GTA_ITEM_PREFETCH 1       // stage the NEXT item's record and first ids / el while the current one is folded
#endif
#ifndef GTA_PUBLISH_FENCE
#define GTA_PUBLISH_FENCE 0       // 1: an extra fence.sc in front of the release store of a chain publish (round-2 form)
#endif
constexpr int kAggThreads = GTA_AGG_THREADS;
constexpr int kAggWarps = kAggThreads / 32;

// floats per partial slot of the GAT kernel: acc[f] | per 128-feature window: max[H] | sum[H], padded to 16 bytes
__host__ __device__ inline int gat_stats_stride(int heads) { return (2 * heads + 3) & ~3; }
__host__ __device__ inline int gat_partial_stride(int f, int heads) { return f + ((f + 127) / 128) * gat_stats_stride(heads); }

// ---- slot chain of a multi-item row ----------------------------------------------------------
// flag[slot] becomes 1 once the state folded over slots [first, slot] is in partials[slot].
__device__ __forceinline__ void chain_wait(const int32_t* flag) {
  // Poll with a RELAXED load and fence once on success.  An acquire load in the loop costs an L1 invalidation per
  // iteration (ptxas emits CCTL.IVALL behind every acquire at gpu scope): with rows split into consecutive items,
  // thousands of polling warps kept every SM's L1 empty and the 8-GPU step went from 0.90 to 2.5 ms.
  int32_t v = 0;
#pragma unroll 1
  for (uint32_t spin = 0; spin < (1u << 27); ++spin) {
    asm volatile("ld.relaxed.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(flag) : "memory");
    if (v != 0) {
      asm volatile("fence.acq_rel.gpu;" ::: "memory");
      return;
    }
    __nanosleep(100);
  }
  __trap();      // the predecessor never published: a protocol bug must not hang the GPU
}
// Called by ONE lane after a __syncwarp of its group: the barrier orders the other lanes' state stores before
// this lane, and its fence + release store make them visible, cumulatively, to whoever acquires the flag -- the
// idiom of a cooperative grid barrier (block barrier, then one thread fences and signals).  One fence per
// item instead of one per lane: a membar.gl is the most expensive instruction of a short item.
__device__ __forceinline__ void chain_publish(int32_t* flag) {
  // st.release is cumulative over what the barrier ordered before this lane (the idiom of CUTLASS's Semaphore::release:
  // barrier, then one thread's st.release.gpu); a __threadfence() in front of it is a second, sequentially consistent
  // fence (membar.gl) per item and buys nothing
#if GTA_PUBLISH_FENCE
  __threadfence();
#endif
  asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(flag), "r"(1) : "memory");
}
// predecessor state was written by another SM during this launch: read it at L2, never from L1
__device__ __forceinline__ float4 ld_state_f32x4(const float* p) { return __ldcg(reinterpret_cast<const float4*>(p)); }
__device__ __forceinline__ float ld_state_f32(const float* p) { return __ldcg(p); }
template <int LANES>
__device__ __forceinline__ uint32_t group_mask(int lane) {
  if constexpr (LANES == 32) return 0xffffffffu;
  else return ((1u << LANES) - 1u) << (lane & ~(LANES - 1));
}
// Run `body` once per group of the warp, groups in ascending order (one pass when the warp is a single
// group or no group of it sits in a chain): a predecessor that lives in the SAME warp has then
// published before its successor waits.
template <int LANES, typename F>
__device__ __forceinline__ void for_groups_in_order(int lane, bool chained, F&& body) {
  if (LANES == 32 || !__any_sync(0xffffffffu, chained)) {
    body();
  } else {
#pragma unroll 1
    for (int g = 0; g < 32 / LANES; ++g) {
      if (lane / LANES == g) body();
      __syncwarp();
    }
  }
}

// ---- the next item, staged while the current one runs ------------------------------------------
// An item costs a chain of dependent loads before its first gather can issue: the item record, then its first source
// ids (streamed from DRAM) and el row, then the er rows of those ids; and two row_slots reads in front of the chain fold.
// Measured on the Reddit shape (agg_probe, items of 20 edges against items of 164): about 5 us of warp time per item
// whatever its length, a quarter of the kernel.  The record of the NEXT item and the slot range of the CURRENT row are
// therefore copied into shared memory asynchronously at the top of an item (no registers held across the gather loop),
// and the next item's el row and first two id batches are requested right after the gather loop, so that they travel
// while the chain fold of the current item waits for its predecessor.
struct NextItem {
  int4 item;
  int32_t s0, s1, pad0, pad1;
};
__device__ __forceinline__ void cp_async_16(void* smem, const void* gmem) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(uint32_t(__cvta_generic_to_shared(smem))), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_4(void* smem, const void* gmem) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(uint32_t(__cvta_generic_to_shared(smem))), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

template <int LANES>
__device__ __forceinline__ float group_max(float v) {
#pragma unroll
  for (int o = LANES / 2; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
template <int LANES>
__device__ __forceinline__ float group_sum(float v) {
#pragma unroll
  for (int o = LANES / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ int warp_max_i32(int v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = max(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
// gathered feature row, 128 bits per lane.  GTA_AGG_GATHER picks the cache policy (measured on B200,
// see DESIGN.md): 0 = L1 no-allocate, 1 = default, 2 = L1 no-allocate + L2 evict_last, 3 = L2 evict_last,
// 4 = as 2 without .nc (coherent path)
#ifndef GTA_AGG_GATHER
#define GTA_AGG_GATHER 4
#endif
__device__ __forceinline__ float4 ld_row_f32x4(const float* p, uint64_t pol_keep) {
  float4 v;
#if GTA_AGG_GATHER == 0
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
#elif GTA_AGG_GATHER == 1
  asm volatile("ld.global.nc.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
#elif GTA_AGG_GATHER == 2
  asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v4.f32 {%0,%1,%2,%3}, [%4], %5;"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p), "l"(pol_keep));
#elif GTA_AGG_GATHER == 3
  asm volatile("ld.global.nc.L2::cache_hint.v4.f32 {%0,%1,%2,%3}, [%4], %5;"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p), "l"(pol_keep));
#else      // 4: as 2 but through the coherent path (no .nc)
  asm volatile("ld.global.L1::no_allocate.L2::cache_hint.v4.f32 {%0,%1,%2,%3}, [%4], %5;"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p), "l"(pol_keep));
#endif
  return v;
}
// address of a gathered row: base + id * row_bytes.  Written as a 64-bit multiply-add of two 32-bit
// values so ptxas emits ONE IMAD.WIDE.U32 with the lane's base pointer as the addend (the round-1 inline
// mad.wide.u32 was split into IMAD.WIDE + IADD3 + IADD3.X once the base pair was not register-aligned).
__device__ __forceinline__ const float* row_ptr(const float* base, uint32_t id, uint32_t row_bytes) {
  return reinterpret_cast<const float*>(reinterpret_cast<const char*>(base) + uint64_t(id) * row_bytes);
}
// exp of a non-positive softmax exponent.  GTA_AGG_FASTEXP=1: ex2.approx path (relative error about
// 2e-7 + |x| 1e-7; terms that matter have small |x|), two instructions instead of about ten.
__device__ __forceinline__ float softmax_exp(float x) {
#if GTA_AGG_FASTEXP
  return __expf(x);
#else
  return expf(x);
#endif
}
__device__ __forceinline__ void fma4(float4& acc, float w, const float4& v) {
  acc.x = fmaf(w, v.x, acc.x);
  acc.y = fmaf(w, v.y, acc.y);
  acc.z = fmaf(w, v.z, acc.z);
  acc.w = fmaf(w, v.w, acc.w);
}
__device__ __forceinline__ float4 epilogue4(float4 a, float scale, int epi) {
  a.x = apply_epilogue(a.x * scale, epi);
  a.y = apply_epilogue(a.y * scale, epi);
  a.z = apply_epilogue(a.z * scale, epi);
  a.w = apply_epilogue(a.w * scale, epi);
  return a;
}

// ----------------------------------------------------------------------------------------
// the work list, its chain state and the dynamic item counter every aggregation launch takes
// ----------------------------------------------------------------------------------------
struct WorkList {
  const int4* items;
  int64_t num_items;
  const int32_t* row_slots;
  int64_t num_slots;
  const int32_t* indices;
  float* partials;
  int32_t* chain_flags;      // [windows][num_slots]
  int32_t* work_counter;     // [windows], zeroed before every launch
  int32_t take;              // 1: warps take items from the counter; 0: static striding (long lists of tiny items)
  uint64_t pol_stream;
  uint64_t pol_keep;
};

// Persistent launch: every warp takes the next 32/LANES items from a global counter until the list is
// empty.  (Round 1 launched one CTA per 4 items: a CTA slot stayed occupied until its longest item was
// done and only 18 of the 24 resident warps per SM were active.)  The grab for the NEXT items is issued
// before the current ones are processed and its result is only read afterwards, so the atomic's round
// trip hides under the gathers.  Items are still started in work-list order, which keeps the CTAs on one
// column block at a time and keeps the chain invariant: whoever holds a predecessor slot started earlier
// and is running, so a wait can never deadlock, whatever the grid size.
// Long lists of tiny items (RMAT: millions of items of a few edges) do not need the balancing and would
// hammer the counter: with wl.take == 0 the warps stride through the list statically (warp w takes groups
// w, w + W, ...; the host then sizes the grid so that all W warps are resident, which the chain argument now
// needs).  Consecutive items still go to different warps -- taking several consecutive items per grab
// instead was measured 3x slower on RMAT-20: a hub row's chain of 1024-edge items then serialises, every
// warp sitting on its predecessor's publish while that warp works through the rest of its batch.
struct ItemCursor {
  int32_t first;       // first item of the warp's current group-step
  int32_t pending;     // dynamic: the next grab (lane 0), in flight
  int32_t stride;      // static: items between two steps of this warp
};
template <int LANES>
__device__ __forceinline__ ItemCursor cursor_begin(const WorkList& wl, const Exchange& ex, int32_t* counter, int lane) {
  constexpr int kGroups = 32 / LANES;
  ItemCursor c;
  if (wl.take > 0) {
    int32_t v = 0;
    if (lane == 0) v = atomicAdd(counter, kGroups);
    c.first = __shfl_sync(0xffffffffu, v, 0);
    c.pending = 0;
    if (lane == 0) c.pending = atomicAdd(counter, kGroups);
    c.stride = 0;
  } else {
    const int copy = ex.world > 1 ? ex.copy_ctas : 0;
    c.first = ((int32_t(blockIdx.x) - copy) * kAggWarps + int32_t(threadIdx.x >> 5)) * kGroups;
    c.stride = (int32_t(gridDim.x) - copy) * kAggWarps * kGroups;
    c.pending = 0;
  }
  return c;
}
template <int LANES>
__device__ __forceinline__ void cursor_next(ItemCursor& c, const WorkList& wl, int32_t* counter, int lane) {
  if (wl.take > 0) {
    c.first = __shfl_sync(0xffffffffu, c.pending, 0);
    if (c.first < wl.num_items && lane == 0) c.pending = atomicAdd(counter, 32 / LANES);
  } else {
    c.first = (c.first > 0x7fffffff - c.stride) ? 0x7fffffff : c.first + c.stride;
  }
}

// ----------------------------------------------------------------------------------------
// storage type of the gathered table: fp32, or bf16 with fp32 accumulation (SURVEY.md section 8d "bf16 mode";
// the reference's IR declares data_format FP16, template/IR_defination.yaml:10-27).  A lane always moves
// 16-byte pieces of a row: 4 fp32 or 8 bf16 features.
// ----------------------------------------------------------------------------------------
// A PIECE is what one lane moves of one gathered row: its storage type, how many features, how many bytes.
//   F32x4   4 fp32 in 16 bytes            (512-byte rows at 32 lanes: the fp32 mode)
//   Bf16x8  8 bf16 in 16 bytes            (rows wider than 128 features)
//   Bf16x4  4 bf16 in  8 bytes            (rows of up to 128 features keep all 32 lanes on ONE item: the per-batch
//                                          work -- staging, softmax -- is then spread over 32 edges, not 16; measured
//                                          on the Reddit shape: Bf16x8 at 16 lanes per item was no faster than fp32)
struct F32x4 {
  using T = float;
  using Raw = uint4;
  static constexpr int kPer = 4, kBytes = 16;
  static __device__ __forceinline__ Raw load(const char* p, uint64_t pol) {
    Raw v;
    asm volatile("ld.global.L1::no_allocate.L2::cache_hint.v4.u32 {%0,%1,%2,%3}, [%4], %5;"
                 : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p), "l"(pol));
    return v;
  }
  static __device__ __forceinline__ Raw zero() { return make_uint4(0u, 0u, 0u, 0u); }
  static __device__ __forceinline__ void unpack(const Raw& r, float (&f)[4]) {
    f[0] = __uint_as_float(r.x); f[1] = __uint_as_float(r.y); f[2] = __uint_as_float(r.z); f[3] = __uint_as_float(r.w);
  }
};
struct Bf16x8 {
  using T = __nv_bfloat16;
  using Raw = uint4;
  static constexpr int kPer = 8, kBytes = 16;
  static __device__ __forceinline__ Raw load(const char* p, uint64_t pol) { return F32x4::load(p, pol); }
  static __device__ __forceinline__ Raw zero() { return make_uint4(0u, 0u, 0u, 0u); }
  static __device__ __forceinline__ void unpack(const Raw& r, float (&f)[8]) {      // bf16 -> fp32 is a shift
    f[0] = __uint_as_float(r.x << 16); f[1] = __uint_as_float(r.x & 0xffff0000u);
    f[2] = __uint_as_float(r.y << 16); f[3] = __uint_as_float(r.y & 0xffff0000u);
    f[4] = __uint_as_float(r.z << 16); f[5] = __uint_as_float(r.z & 0xffff0000u);
    f[6] = __uint_as_float(r.w << 16); f[7] = __uint_as_float(r.w & 0xffff0000u);
  }
};
struct Bf16x4 {
  using T = __nv_bfloat16;
  using Raw = uint2;
  static constexpr int kPer = 4, kBytes = 8;
  static __device__ __forceinline__ Raw load(const char* p, uint64_t pol) {
    Raw v;
    asm volatile("ld.global.L1::no_allocate.L2::cache_hint.v2.u32 {%0,%1}, [%2], %3;"
                 : "=r"(v.x), "=r"(v.y) : "l"(p), "l"(pol));
    return v;
  }
  static __device__ __forceinline__ Raw zero() { return make_uint2(0u, 0u); }
  static __device__ __forceinline__ void unpack(const Raw& r, float (&f)[4]) {
    f[0] = __uint_as_float(r.x << 16); f[1] = __uint_as_float(r.x & 0xffff0000u);
    f[2] = __uint_as_float(r.y << 16); f[3] = __uint_as_float(r.y & 0xffff0000u);
  }
};
__device__ __forceinline__ const char* row_addr(const char* base, uint32_t id, uint32_t row_bytes) {
  return base + uint64_t(id) * row_bytes;
}
template <typename P>
__device__ __forceinline__ void fma_row(float (&acc)[P::kPer], float w, const typename P::Raw& raw) {
  float f[P::kPer];
  P::unpack(raw, f);
#pragma unroll
  for (int c = 0; c < P::kPer; ++c) acc[c] = fmaf(w, f[c], acc[c]);
}
// kPer consecutive fp32 of an output / partial row
template <int KP>
__device__ __forceinline__ void st_out(float* p, const float (&a)[KP], float scale, int epi) {
#pragma unroll
  for (int q = 0; q < KP / 4; ++q)
    st_stream_f32x4(p + 4 * q, make_float4(apply_epilogue(a[4 * q] * scale, epi), apply_epilogue(a[4 * q + 1] * scale, epi),
                                           apply_epilogue(a[4 * q + 2] * scale, epi), apply_epilogue(a[4 * q + 3] * scale, epi)));
}
template <int KP>
__device__ __forceinline__ void st_state(float* p, const float (&a)[KP]) {
#pragma unroll
  for (int q = 0; q < KP / 4; ++q)
    *reinterpret_cast<float4*>(p + 4 * q) = make_float4(a[4 * q], a[4 * q + 1], a[4 * q + 2], a[4 * q + 3]);
}
template <int KP>
__device__ __forceinline__ void ld_state(const float* p, float (&a)[KP]) {
#pragma unroll
  for (int q = 0; q < KP / 4; ++q) {
    const float4 t = ld_state_f32x4(p + 4 * q);
    a[4 * q] = t.x; a[4 * q + 1] = t.y; a[4 * q + 2] = t.z; a[4 * q + 3] = t.w;
  }
}

// ----------------------------------------------------------------------------------------
// weighted aggregate
//   WKIND 0: no weight, 1: scalar weight per edge (wh == 1), 2: per-head weight (wh > 1,
//   (f / wh) % kPer == 0 so a lane's features share a head; V == 1 only)
//   V: 16-byte pieces per lane and gathered row.  V = 2 lets one warp take a whole 1 KB row (256 fp32
//   features, the RMAT config) in one pass instead of walking the work list once per 128-feature window:
//   the indices, the item records and the page-table entries of a row are then touched once, not twice.
// ----------------------------------------------------------------------------------------
constexpr int kAggUnroll = GTA_AGG_UNROLL;
constexpr int kGatUnroll = GTA_GAT_UNROLL;
constexpr int kLlhUnroll = GTA_LLH_UNROLL;

template <typename P, int V, int LANES, int WKIND, bool DIV>
__global__ void __launch_bounds__(kAggThreads, (V == 1 && P::kPer == 4) ? GTA_AGG_MINBLOCKS : 8)
aggregate_kernel(const WorkList wl, const Exchange ex, const float* __restrict__ w, int wh,
                 const float* __restrict__ rowden, const typename P::T* __restrict__ x, const uint32_t row_bytes,
                 float* __restrict__ out, int64_t ldo, int f, int epilogue) {
  static_assert(V == 1 || (LANES == 32 && WKIND != 2), "two pieces per lane: full warps, no per-head weights");
  using Raw = typename P::Raw;
  constexpr int KP = P::kPer;
  constexpr int kWindow = LANES * KP * V;          // features one pass of a group covers
  constexpr int kEdges = kAggUnroll;      // edges whose loads (V each) are in flight together
  __shared__ __align__(16) uint2 s_a[kAggWarps][32];        // per warp: {source id, weight} of the staged batch
#if GTA_ITEM_PREFETCH
  __shared__ __align__(16) NextItem s_next[kAggWarps][32 / LANES];
#endif
  if (ex.world > 1 && blockIdx.y == 0 && blockIdx.x < ex.copy_ctas) {
    exchange_pull(ex, blockIdx.x);          // these CTAs move the peers' slots; everybody else reduces
    return;
  }
  const int lane = threadIdx.x & 31;
  const int l = lane & (LANES - 1);
  const int fo = blockIdx.y * kWindow + KP * l;          // piece v sits at fo + v * LANES * KP
  uint2* se = s_a[threadIdx.x >> 5];
  const uint2* mine = se + (lane & ~(LANES - 1));
  const uint4* mine2 = reinterpret_cast<const uint4*>(mine);      // two staged edges per LDS.128
  int32_t* counter = wl.work_counter + blockIdx.y;
  int32_t* flags = wl.chain_flags + int64_t(blockIdx.y) * wl.num_slots;
  const uint64_t pol_stream = wl.pol_stream, pol_keep = wl.pol_keep;
  bool on[V];
#pragma unroll
  for (int v = 0; v < V; ++v) on[v] = fo + v * LANES * KP < f;
  const char* xf = reinterpret_cast<const char*>(x + (on[0] ? fo : 0));

  ItemCursor cur = cursor_begin<LANES>(wl, ex, counter, lane);
  int32_t landed = 0;          // exchange_gate: highest peer slot this lane has seen complete
  const int head = (WKIND == 2 && on[0]) ? fo / (f / wh) : 0;
  // what an item needs before its first batch (see NextItem): record, first ids (and scalar weights), the row's
  // denominator, and -- in an exchange -- its last id, which names the highest slot it touches
  int4 it = make_int4(0, 0, 0, -1);
  bool have = false;
  int idx_nxt = 0, last_src = 0;
  float w_nxt = 0.f, den = 1.f;
  auto request_inputs = [&](const int4& t, bool hv) {
    const int cnt = hv ? t.z : 0;
    const int32_t* ib = wl.indices + t.y;
    idx_nxt = 0;
    w_nxt = 0.f;
    den = 1.f;
    if (l < cnt) {
      idx_nxt = ld_stream_i32(ib + l, pol_stream);
      if (WKIND == 1) w_nxt = ld_stream_f32(w + int64_t(t.y) * wh + l, pol_stream);
    }
    if (DIV && hv) den = rowden[int64_t(t.x) * wh + head];
    last_src = (ex.world > 1 && cnt > 0) ? __ldg(ib + cnt - 1) : 0;
  };
#if GTA_ITEM_PREFETCH
  NextItem* nx = &s_next[threadIdx.x >> 5][lane / LANES];
  {
    const int64_t group = int64_t(cur.first) + lane / LANES;
    have = group < wl.num_items;
    if (have) it = __ldg(wl.items + group);
    request_inputs(it, have);
  }
#endif
  while (cur.first < wl.num_items) {
#if GTA_ITEM_PREFETCH
    ItemCursor nxt = cur;
    cursor_next<LANES>(nxt, wl, counter, lane);
    const int64_t ngroup = int64_t(nxt.first) + lane / LANES;
    const bool nhave = nxt.first < wl.num_items && ngroup < wl.num_items;
    if (l == 0) {
      if (nhave) cp_async_16(&nx->item, wl.items + ngroup);
      if (have && it.w >= 0) {
        cp_async_4(&nx->s0, wl.row_slots + it.x);
        cp_async_4(&nx->s1, wl.row_slots + it.x + 1);
      }
    }
#else
    {
      const int64_t group = int64_t(cur.first) + lane / LANES;
      have = group < wl.num_items;
      it = have ? __ldg(wl.items + group) : make_int4(0, 0, 0, -1);
      request_inputs(it, have);
    }
#endif
    const bool active = have && on[0];
    const int count = have ? it.z : 0;
    const int max_count = (LANES == 32) ? count : warp_max_i32(count);
    const int32_t* idx_base = wl.indices + it.y;
    const float* w_base = (WKIND != 0) ? w + int64_t(it.y) * wh : nullptr;
    const float den_cur = den;          // the next item's denominator replaces `den` before this item is folded

    float acc[V][KP];
#pragma unroll
    for (int v = 0; v < V; ++v)
#pragma unroll
      for (int c = 0; c < KP; ++c) acc[v][c] = 0.f;
    // software pipeline: ids (and scalar weights) of batch b+1 are in flight under the gathers of batch b
    if (ex.world > 1) {          // the item's slots may still be on their way from the peers
      if (l == 0 && count > 0) exchange_gate(ex, last_src, landed);
      __syncwarp();
    }
    for (int base = 0; base < max_count; base += LANES) {
      int n = count - base;
      n = n < 0 ? 0 : (n > LANES ? LANES : n);
      const int my_idx = idx_nxt;
      float my_w = (l < n) ? w_nxt : 0.f;
      // divide only where an edge exists: a neighbouring group's empty row has den = 0 (0/0 = NaN)
      if (WKIND == 1 && DIV && l < n) my_w = my_w / den_cur;
      if (base + LANES + l < count) {
        idx_nxt = ld_stream_i32(idx_base + base + LANES + l, pol_stream);
        if (WKIND == 1) w_nxt = ld_stream_f32(w_base + base + LANES + l, pol_stream);
      }
      // stage {source id, weight} of the batch in shared memory: one LDS.128 per two edges in the gather loop
      se[lane] = make_uint2(uint32_t(my_idx), __float_as_uint(WKIND == 1 ? my_w : 1.f));
      __syncwarp();
      const bool full = (LANES == 32) ? (n == LANES) : __all_sync(0xffffffffu, n == LANES && active);
      if (full && (V == 1 || on[V - 1])) {
        // whole batch, no predicates: kEdges * V loads in flight, then their FMAs (the outer loop stays
        // rolled: unrolled, ptxas hoists every load of the batch and spills)
        if (LANES < 32 || active) {
#pragma unroll 1
          for (int j = 0; j < LANES; j += kEdges) {
            uint4 ed[kEdges / 2];
            Raw raw[kEdges][V];
#pragma unroll
            for (int u = 0; u < kEdges / 2; ++u)
              if (j + 2 * u < LANES) ed[u] = mine2[(j >> 1) + u];
#pragma unroll
            for (int u = 0; u < kEdges; ++u) {
              if (j + u < LANES) {
                const char* rp = row_addr(xf, (u & 1) ? ed[u / 2].z : ed[u / 2].x, row_bytes);
#pragma unroll
                for (int v = 0; v < V; ++v) raw[u][v] = P::load(rp + v * LANES * P::kBytes, pol_keep);
              }
            }
#pragma unroll
            for (int u = 0; u < kEdges; ++u) {
              if (j + u < LANES) {
                float ws = __uint_as_float((u & 1) ? ed[u / 2].w : ed[u / 2].y);
                if (WKIND == 2) {
                  ws = __ldg(w_base + int64_t(base + j + u) * wh + head);
                  if (DIV) ws = ws / den_cur;
                }
#pragma unroll
                for (int v = 0; v < V; ++v) fma_row<P>(acc[v], ws, raw[u][v]);
              }
            }
          }
        }
      } else {
        const int nmax = (LANES == 32) ? n : LANES;
        for (int j = 0; j < nmax; j += kEdges) {
          Raw raw[kEdges][V];
          float wv[kEdges];
#pragma unroll
          for (int u = 0; u < kEdges; ++u) {
            if (j + u < LANES) {
              const uint2 ed = mine[j + u];
              const bool ok = have && (j + u) < n;
              float ws = ok ? __uint_as_float(ed.y) : 0.f;
              const char* rp = row_addr(xf, ed.x, row_bytes);
#pragma unroll
              for (int v = 0; v < V; ++v) {
                raw[u][v] = P::zero();
                if (ok && on[v]) raw[u][v] = P::load(rp + v * LANES * P::kBytes, pol_keep);
              }
              if (WKIND == 2) {
                ws = 0.f;
                if (ok && active) {
                  ws = __ldg(w_base + int64_t(base + j + u) * wh + head);
                  if (DIV) ws = ws / den_cur;
                }
              }
              wv[u] = ws;
            }
          }
#pragma unroll
          for (int u = 0; u < kEdges; ++u)
            if (j + u < LANES)
#pragma unroll
              for (int v = 0; v < V; ++v) fma_row<P>(acc[v], wv[u], raw[u][v]);
        }
      }
      __syncwarp();
    }
    const bool chained = have && it.w >= 0;
#if GTA_ITEM_PREFETCH
    cp_async_wait_all();
    __syncwarp();
    const int4 itn = nhave ? nx->item : make_int4(0, 0, 0, -1);
    const int slot0 = chained ? nx->s0 : 0, slot1 = chained ? nx->s1 : 0;
    __syncwarp();          // everybody has read the staging entry before lane 0 of the group overwrites it
    request_inputs(itn, nhave);
#endif
    for_groups_in_order<LANES>(lane, chained, [&]() {
      bool last = true;
      if (chained) {
#if GTA_ITEM_PREFETCH
        const int s0 = slot0, s1 = slot1;
#else
        const int s0 = __ldg(wl.row_slots + it.x), s1 = __ldg(wl.row_slots + it.x + 1);
#endif
        last = it.w == s1 - 1;
        if (it.w != s0) {
          if (l == 0) chain_wait(flags + it.w - 1);
          __syncwarp(group_mask<LANES>(lane));
#pragma unroll
          for (int v = 0; v < V; ++v) {
            if (have && on[v]) {
              float p[KP];
              ld_state<KP>(wl.partials + int64_t(it.w - 1) * f + fo + v * LANES * KP, p);
#pragma unroll
              for (int c = 0; c < KP; ++c) acc[v][c] = p[c] + acc[v][c];
            }
          }
        }
        if (!last) {
#pragma unroll
          for (int v = 0; v < V; ++v)
            if (have && on[v]) st_state<KP>(wl.partials + int64_t(it.w) * f + fo + v * LANES * KP, acc[v]);
          __syncwarp(group_mask<LANES>(lane));          // the group's stores happen-before lane 0's release
          if (l == 0) chain_publish(flags + it.w);
        }
      }
      if (last) {
#pragma unroll
        for (int v = 0; v < V; ++v)
          if (have && on[v]) st_out<KP>(out + int64_t(it.x) * ldo + fo + v * LANES * KP, acc[v], 1.f, epilogue);
      }
    });
#if GTA_ITEM_PREFETCH
    cur = nxt;
    it = itn;
    have = nhave;
#else
    cursor_next<LANES>(cur, wl, counter, lane);
#endif
  }
}

// ----------------------------------------------------------------------------------------
// GAT edge phase, single pass
// ----------------------------------------------------------------------------------------
template <int H>
__device__ __forceinline__ void load_heads(const float* __restrict__ p, float (&v)[H]) {
  if (H % 4 == 0) {
#pragma unroll
    for (int q = 0; q < H / 4; ++q) {
      float4 t = __ldg(reinterpret_cast<const float4*>(p) + q);
      v[4 * q] = t.x; v[4 * q + 1] = t.y; v[4 * q + 2] = t.z; v[4 * q + 3] = t.w;
    }
  } else if (H % 2 == 0) {
#pragma unroll
    for (int q = 0; q < H / 2; ++q) {
      float2 t = __ldg(reinterpret_cast<const float2*>(p) + q);
      v[2 * q] = t.x; v[2 * q + 1] = t.y;
    }
  } else {
#pragma unroll
    for (int q = 0; q < H; ++q) v[q] = __ldg(p + q);
  }
}

template <int H>
__device__ __forceinline__ float pick(const float (&v)[H], int h) {
  float r = v[0];
#pragma unroll
  for (int q = 1; q < H; ++q) r = (h == q) ? v[q] : r;
  return r;
}

// ---- softmax shift from a BOUND instead of the running maximum -----------------------------------
// leaky_relu is monotonic, so for every edge of row i whose source lies in column block cb
//     e = leaky(el[i,h] + er[j,h])  <=  leaky(el[i,h] + max_{j in cb} er[j,h])  =: bound(i, cb, h).
// Softmax is invariant under the shift, so p = exp(e - bound) needs no running maximum: no warp
// reductions, no rescale of the accumulator, one exp per edge and head instead of two (round 2 ncu: the 20
// shuffles per 32-edge batch were 0.9 L1/TEX data-pipe wavefronts per edge, the exps 0.6 ms of 5.0).
// A loose bound only costs exponent range, never precision: as long as max er - min er of the block is
// below kBoundRange every p stays above exp(-kBoundRange) relative to the row's largest term.  er_stats
// (gta_er_stats: ordered-int coded max er and max -er per column block and head) says so; blocks that
// fail the test, heads counts that are no power of two and calls that want the true row maximum back
// take the online path below.
constexpr float kBoundRange = 60.f;
// er_stats[cb*pitch + h] = code(max er), er_stats[cb*pitch + heads + h] = code(max -er); 0 = "no source seen".
// pitch = 2*heads for a gta_er_stats buffer, 64 for the statistics of a signal block (exchange.cuh).
__device__ __forceinline__ bool block_bound(const uint32_t* er_stats, int64_t cb0, int64_t cb1, int pitch, int heads,
                                            int h, float* er_max) {
  // an item may span several statistics blocks (an exchange groups its peers' slots): the bound and the
  // range test are taken over their union
  uint32_t cmax = 0u, cneg = 0u;
  bool seen = true;
  for (int64_t cb = cb0; cb <= cb1; ++cb) {
    const uint32_t a = __ldcg(er_stats + cb * pitch + h), b = __ldcg(er_stats + cb * pitch + heads + h);
    seen = seen && a != 0u && b != 0u;
    cmax = a > cmax ? a : cmax;
    cneg = b > cneg ? b : cneg;
  }
  const float hi = ordered_decode(cmax), lo = -ordered_decode(cneg);
  *er_max = hi;
  return seen && (hi - lo) < kBoundRange;      // NaN compares false
}

// ---- the bound, looked up per item -----------------------------------------------------------------
// Round-2 ncu of the GAT kernel on a low-degree shape (items of 20 edges, profiles/r02_gat_lowdeg_*): computing the
// bound cost about 230 of an item's 1 180 warp instructions and 12 % of its stall samples -- two emulated 64-bit
// divisions for the statistics blocks of the first and last source, then, head by head, a loop over those blocks
// whose two L2 loads are consumed inside the loop: H serialised L2 round trips in front of every item's first batch.
// Consecutive items of a warp almost always touch the same statistics blocks, so the warp (each lane group of it)
// keeps the last answer in shared memory, keyed by (first block, last block): the bound stays a pure function of the
// item, hence bitwise the same whoever computes it, and a hit costs two multiply-high divisions and one LDS.
template <int MAXH>
struct __align__(16) BoundCache {
  int32_t cb0, cb1, ok, pad;
  float hi[MAXH];
};
struct BlockDivider {          // source id -> statistics block, without the 64-bit division
  uint32_t d, magic;
  __device__ __forceinline__ explicit BlockDivider(int64_t col_block)
      : d(col_block > 0 ? uint32_t(col_block) : 0u), magic(d > 1u ? uint32_t((uint64_t(1) << 32) / d) : 0u) {}
  __device__ __forceinline__ int operator()(int src) const {
    if (d <= 1u) return d == 0u ? 0 : src;
    uint32_t q = __umulhi(uint32_t(src), magic);          // floor(2^32 / d): never above the quotient, at most 2 below
    uint32_t r = uint32_t(src) - q * d;
    while (r >= d) { ++q; r -= d; }
    return int(q);
  }
};
template <int LANES, int MAXH>
__device__ __forceinline__ void bound_lookup(BoundCache<MAXH>* bc, const uint32_t* er_stats, int cb0, int cb1, int pitch,
                                             int heads, int l, uint32_t gmask) {
  if (bc->cb0 != cb0 || bc->cb1 != cb1) {          // the same answer in every lane of the group
    __syncwarp(gmask);          // everybody has read the old key
    bool ok = true;
    for (int h = l; h < heads; h += LANES) {
      float hi;
      ok = block_bound(er_stats, cb0, cb1, pitch, heads, h, &hi) && ok;
      bc->hi[h] = hi;
    }
    ok = __all_sync(gmask, ok);
    if (l == 0) { bc->cb0 = cb0; bc->cb1 = cb1; bc->ok = ok ? 1 : 0; }
    __syncwarp(gmask);
  }
}

template <typename P, int LANES, int H>
__global__ void __launch_bounds__(kAggThreads, (H <= 4) ? GTA_GAT_MINBLOCKS : (GTA_GAT_MINBLOCKS + 1) / 2)
gat_aggregate_kernel(const WorkList wl, const Exchange ex, const float* __restrict__ el, const float* __restrict__ er,
                     int64_t lder, float slope, const typename P::T* __restrict__ z, const uint32_t row_bytes,
                     float* __restrict__ out, int64_t ldo, int f, int epilogue, float* __restrict__ rowmax,
                     float* __restrict__ rowsum, const uint32_t* er_stats, int stats_pitch, int64_t col_block) {
  // per warp: H rows of 32 staged edges, entry = {source id, softmax numerator}.  Row pitch kS = 34
  // entries: a lane's STS.64 lands beside its neighbour's (2 wavefronts per head, no conflicts) and the
  // LDS.128 of the gather loop -- two consecutive edges of one head, the 4 heads of a warp at once --
  // hits 4 disjoint bank quads (68 words = 4 mod 32).  Round 1 staged [edge][head]: 4-way conflicts on
  // every store, 27 % of the L1/TEX data-pipe wavefronts of the kernel.
  constexpr int kS = 34;
  using Raw = typename P::Raw;
  constexpr int KP = P::kPer;
  constexpr int kWindow = LANES * KP;
  __shared__ __align__(16) uint2 s_e[kAggWarps][H * kS];
  __shared__ BoundCache<H> s_bound[kAggWarps][32 / LANES];
#if GTA_ITEM_PREFETCH
  __shared__ __align__(16) NextItem s_next[kAggWarps][32 / LANES];
#endif
  if (ex.world > 1 && blockIdx.y == 0 && blockIdx.x < ex.copy_ctas) {
    exchange_pull(ex, blockIdx.x);          // these CTAs move the peers' slots; everybody else reduces
    return;
  }
  const int lane = threadIdx.x & 31;
  const int l = lane & (LANES - 1);
  const int gbase = lane & ~(LANES - 1);          // first lane of my group inside the warp
  BoundCache<H>* bc = &s_bound[threadIdx.x >> 5][lane / LANES];
  if (l == 0) { bc->cb0 = -1; bc->cb1 = -1; bc->ok = 0; }
  __syncwarp();
  const BlockDivider block_of(col_block);
  const int fo = blockIdx.y * kWindow + KP * l;
  const int head = (fo < f) ? fo / (f / H) : 0;
  uint2* se = s_e[threadIdx.x >> 5];
  const uint2* mine = se + head * kS + gbase;
  const uint4* mine2 = reinterpret_cast<const uint4*>(mine);
  int32_t* counter = wl.work_counter + blockIdx.y;
  int32_t* flags = wl.chain_flags + int64_t(blockIdx.y) * wl.num_slots;
  const uint64_t pol_stream = wl.pol_stream, pol_keep = wl.pol_keep;
  const int pstride = gat_partial_stride(f, H);
  const int stats = f + int(blockIdx.y) * gat_stats_stride(H);
  const char* zf = reinterpret_cast<const char*>(z + (fo < f ? fo : 0));

  ItemCursor cur = cursor_begin<LANES>(wl, ex, counter, lane);
  int32_t landed = 0;          // exchange_gate: highest peer slot this lane has seen complete
  // what an item needs before its first batch: the record, the row's el, the ids of its first two batches and its
  // last id (sources ascend: the last id names the highest slot / statistics block the item touches)
  int4 it = make_int4(0, 0, 0, -1);
  bool have = false;
  float elr[H];
  int idx_cur = 0, idx_nxt = 0, last_src = 0;
  const bool want_last = ex.world > 1 || er_stats != nullptr;
  auto request_inputs = [&](const int4& t, bool hv) {
    const int cnt = hv ? t.z : 0;
    const int32_t* ib = wl.indices + t.y;
#pragma unroll
    for (int h = 0; h < H; ++h) elr[h] = 0.f;
    if (hv) load_heads<H>(el + int64_t(t.x) * H, elr);
    idx_cur = 0;
    idx_nxt = 0;
    if (l < cnt) idx_cur = ld_stream_i32(ib + l, pol_stream);
    if (LANES + l < cnt) idx_nxt = ld_stream_i32(ib + LANES + l, pol_stream);
    last_src = (cnt > 0 && want_last) ? __ldg(ib + cnt - 1) : 0;
  };
#if GTA_ITEM_PREFETCH
  NextItem* nx = &s_next[threadIdx.x >> 5][lane / LANES];
  {
    const int64_t group = int64_t(cur.first) + lane / LANES;
    have = group < wl.num_items;
    if (have) it = __ldg(wl.items + group);
    request_inputs(it, have);
  }
#endif
  while (cur.first < wl.num_items) {
#if GTA_ITEM_PREFETCH
    // claim the next item now and let its record (and this row's slot range) travel into shared memory under the gathers
    ItemCursor nxt = cur;
    cursor_next<LANES>(nxt, wl, counter, lane);
    const int64_t ngroup = int64_t(nxt.first) + lane / LANES;
    const bool nhave = nxt.first < wl.num_items && ngroup < wl.num_items;
    if (l == 0) {
      if (nhave) cp_async_16(&nx->item, wl.items + ngroup);
      if (have && it.w >= 0) {
        cp_async_4(&nx->s0, wl.row_slots + it.x);
        cp_async_4(&nx->s1, wl.row_slots + it.x + 1);
      }
    }
#else
    {
      const int64_t group = int64_t(cur.first) + lane / LANES;
      have = group < wl.num_items;
      it = have ? __ldg(wl.items + group) : make_int4(0, 0, 0, -1);
      request_inputs(it, have);
    }
#endif
    const bool active = have && fo < f;
    const int count = have ? it.z : 0;
    const int max_count = (LANES == 32) ? count : warp_max_i32(count);
    const int32_t* idx_base = wl.indices + it.y;

    float m[H], s[H];      // s: this lane's share of the running sum (reduced at the end)
#pragma unroll
    for (int h = 0; h < H; ++h) { m[h] = -INFINITY; s[h] = 0.f; }
    float acc[KP];
#pragma unroll
    for (int c = 0; c < KP; ++c) acc[c] = 0.f;

    // software pipeline: source ids are loaded two batches ahead and the er rows one batch ahead, so
    // the id -> er -> softmax dependency chain of batch b+1 hides under the row gathers of batch b
    float er_cur[H];
#pragma unroll
    for (int h = 0; h < H; ++h) er_cur[h] = 0.f;
    if (ex.world > 1) {          // the item's slots (z, er and their er range) may still be on their way from the peers
      if (l == 0 && count > 0) exchange_gate(ex, last_src, landed);
      __syncwarp();
    }
    if (l < count) load_heads<H>(er + int64_t(idx_cur) * lder, er_cur);

    // bound path: the whole warp or nobody (the online path reduces with full-warp shuffles)
    bool bounded = false;
    if (er_stats != nullptr) {
      bool ok = true;
      const int first_src = __shfl_sync(0xffffffffu, idx_cur, gbase);
      if (count > 0) {
        bound_lookup<LANES, H>(bc, er_stats, block_of(first_src), block_of(last_src), stats_pitch, H, l, group_mask<LANES>(lane));
        ok = bc->ok != 0;
#pragma unroll
        for (int h = 0; h < H; ++h) m[h] = leaky(elr[h] + bc->hi[h], slope);
      }
      bounded = __all_sync(0xffffffffu, ok);
      if (!bounded) {
#pragma unroll
        for (int h = 0; h < H; ++h) m[h] = -INFINITY;
      }
    }

    for (int base = 0; base < max_count; base += LANES) {
      int n = count - base;
      n = n < 0 ? 0 : (n > LANES ? LANES : n);
      float e[H];
      const int my_idx = idx_cur;
#pragma unroll
      for (int h = 0; h < H; ++h) e[h] = (l < n) ? leaky(elr[h] + er_cur[h], slope) : -INFINITY;
      // prefetch: er of the next batch (its ids arrived during the previous iteration), ids of the one after
      idx_cur = idx_nxt;
      if (base + LANES + l < count) load_heads<H>(er + int64_t(idx_cur) * lder, er_cur);
      if (base + 2 * LANES + l < count) idx_nxt = ld_stream_i32(idx_base + base + 2 * LANES + l, pol_stream);
      if (bounded) {
#pragma unroll
        for (int h = 0; h < H; ++h) {
          const float p = (l < n) ? softmax_exp(e[h] - m[h]) : 0.f;
          s[h] += p;
          se[h * kS + lane] = make_uint2(uint32_t(my_idx), __float_as_uint(p));
        }
      } else {
        float my_scale = 1.f;
#pragma unroll
        for (int h = 0; h < H; ++h) {
          const float mn = fmaxf(m[h], group_max<LANES>(e[h]));
          // mn stays -inf only while this group has seen no edge (another group in the warp is running)
          const float sc = (mn == -INFINITY) ? 1.f : softmax_exp(m[h] - mn);
          const float p = (l < n) ? softmax_exp(e[h] - mn) : 0.f;
          s[h] = fmaf(s[h], sc, p);
          m[h] = mn;
          my_scale = (h == head) ? sc : my_scale;
          se[h * kS + lane] = make_uint2(uint32_t(my_idx), __float_as_uint(p));
        }
#pragma unroll
        for (int c = 0; c < KP; ++c) acc[c] *= my_scale;
      }
      __syncwarp();
      const bool full = (LANES == 32) ? (n == LANES) : __all_sync(0xffffffffu, n == LANES && active);
      if (full) {
        if (LANES < 32 || active) {
#pragma unroll 1
          for (int j = 0; j < LANES; j += kGatUnroll) {
            uint4 ed[kGatUnroll / 2];
            Raw raw[kGatUnroll];
#pragma unroll
            for (int u = 0; u < kGatUnroll / 2; ++u)
              if (j + 2 * u < LANES) ed[u] = mine2[(j >> 1) + u];
#pragma unroll
            for (int u = 0; u < kGatUnroll / 2; ++u) {
              if (j + 2 * u < LANES) {
                raw[2 * u] = P::load(row_addr(zf, ed[u].x, row_bytes), pol_keep);
                raw[2 * u + 1] = P::load(row_addr(zf, ed[u].z, row_bytes), pol_keep);
              }
            }
#pragma unroll
            for (int u = 0; u < kGatUnroll / 2; ++u) {
              if (j + 2 * u < LANES) {
                fma_row<P>(acc, __uint_as_float(ed[u].y), raw[2 * u]);
                fma_row<P>(acc, __uint_as_float(ed[u].w), raw[2 * u + 1]);
              }
            }
          }
        }
      } else {
        const int nmax = (LANES == 32) ? n : LANES;
        for (int j = 0; j < nmax; j += kGatUnroll) {
          Raw raw[kGatUnroll];
          float pv[kGatUnroll];
#pragma unroll
          for (int u = 0; u < kGatUnroll; ++u) {
            if (j + u < LANES) {
              const uint2 ed = mine[j + u];
              pv[u] = __uint_as_float(ed.y);
              raw[u] = P::zero();
              if (active && (j + u) < n) raw[u] = P::load(row_addr(zf, ed.x, row_bytes), pol_keep);
            }
          }
#pragma unroll
          for (int u = 0; u < kGatUnroll; ++u)
            if (j + u < LANES) fma_row<P>(acc, pv[u], raw[u]);
        }
      }
      __syncwarp();
    }
#pragma unroll
    for (int h = 0; h < H; ++h) s[h] = group_sum<LANES>(s[h]);
    const bool chained = have && it.w >= 0;
#if GTA_ITEM_PREFETCH
    // the gather loop is over: elr / idx_* are free, the staged record has long arrived.  Request the next item's
    // inputs now; they travel while this item's chain fold waits for its predecessor and writes its state.
    cp_async_wait_all();
    __syncwarp();
    const int4 itn = nhave ? nx->item : make_int4(0, 0, 0, -1);
    const int slot0 = chained ? nx->s0 : 0, slot1 = chained ? nx->s1 : 0;
    __syncwarp();          // everybody has read the staging entry before lane 0 of the group overwrites it
    request_inputs(itn, nhave);
#endif
    for_groups_in_order<LANES>(lane, chained, [&]() {
      bool last = true;
      if (chained) {
#if GTA_ITEM_PREFETCH
        const int s0 = slot0, s1 = slot1;
#else
        const int s0 = __ldg(wl.row_slots + it.x), s1 = __ldg(wl.row_slots + it.x + 1);
#endif
        last = it.w == s1 - 1;
        if (it.w != s0) {
          // fold the state of slots [s0, it.w) in: (max, sum, acc) triples merge like the online softmax itself
          if (l == 0) chain_wait(flags + it.w - 1);
          __syncwarp(group_mask<LANES>(lane));
          const float* prev = wl.partials + int64_t(it.w - 1) * pstride;
          float a_mine = 1.f, b_mine = 1.f;
#pragma unroll
          for (int h = 0; h < H; ++h) {
            const float pm = ld_state_f32(prev + stats + h), ps = ld_state_f32(prev + stats + H + h);
            const float mn = fmaxf(pm, m[h]);
            // one of the two factors is exp(0) = 1: a single exp per head (bit-identical to computing both)
            const float t = (mn == -INFINITY) ? 0.f : expf(fminf(pm, m[h]) - mn);
            const float a = (pm == -INFINITY) ? 0.f : (pm == mn ? 1.f : t);
            const float b = (m[h] == -INFINITY) ? 0.f : (m[h] == mn ? 1.f : t);
            s[h] = fmaf(ps, a, s[h] * b);
            m[h] = mn;
            a_mine = (h == head) ? a : a_mine;
            b_mine = (h == head) ? b : b_mine;
          }
          if (active) {
            float p[KP];
            ld_state<KP>(prev + fo, p);
#pragma unroll
            for (int c = 0; c < KP; ++c) acc[c] = fmaf(p[c], a_mine, acc[c] * b_mine);
          }
        }
        if (!last) {
          float* part = wl.partials + int64_t(it.w) * pstride;
          if (active) st_state<KP>(part + fo, acc);
          if (l < H) {
            part[stats + l] = pick<H>(m, l);
            part[stats + H + l] = pick<H>(s, l);
          }
          __syncwarp(group_mask<LANES>(lane));          // the group's stores happen-before lane 0's release
          if (l == 0) chain_publish(flags + it.w);
        }
      }
      if (last && have) {
        if (active) {
          const float sh = pick<H>(s, head);
          st_out<KP>(out + int64_t(it.x) * ldo + fo, acc, sh > 0.f ? 1.f / sh : 0.f, epilogue);
        }
        if (blockIdx.y == 0 && l < H) {
          const float ml = pick<H>(m, l);
          if (rowmax) rowmax[int64_t(it.x) * H + l] = (count > 0 || it.w >= 0) && ml != -INFINITY ? ml : 0.f;
          if (rowsum) rowsum[int64_t(it.x) * H + l] = pick<H>(s, l);
        }
      }
    });
#if GTA_ITEM_PREFETCH
    cur = nxt;
    it = itn;
    have = nhave;
#else
    cursor_next<LANES>(cur, wl, counter, lane);
#endif
  }
}

// ----------------------------------------------------------------------------------------
// GAT edge phase, lane-local-head variant (any H whose per-head width F/H is a multiple of 4, or 2, or 1;
// used for H >= 8 and for the narrow heads of the reference's third GAT layer, F = H = 16)
//
// The staged kernel above keeps el/max/sum/er for ALL heads in every lane (5H registers: H = 16
// spills and runs at a quarter of the H = 4 speed).  Here a lane tracks only the heads its own 4
// features belong to -- HPL = 1 head when the per-head width is a multiple of 4, 2 heads of width 2, 4 heads
// of width 1: one er gather of HPL floats per edge (the lanes of a row read the H consecutive floats of
// er[j]: one wavefront), softmax over groups of a few edges, no arrays over all heads, no shuffles.  Lanes
// of one head see the same edges in the same order, so their (max, sum) are bit-identical.
// ----------------------------------------------------------------------------------------
template <int N>
__device__ __forceinline__ void ldg_vec(const float* p, float (&v)[N]) {
  if constexpr (N == 4) {
    const float4 t = __ldg(reinterpret_cast<const float4*>(p));
    v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
  } else if constexpr (N == 2) {
    const float2 t = __ldg(reinterpret_cast<const float2*>(p));
    v[0] = t.x; v[1] = t.y;
  } else {
    v[0] = __ldg(p);
  }
}
// which of a lane's HPL heads its feature c (0..3) belongs to
template <int HPL>
__device__ __forceinline__ constexpr int head_of(int c) { return HPL == 1 ? 0 : (HPL == 2 ? c / 2 : c); }

template <int LANES, int HPL>
__global__ void __launch_bounds__(kAggThreads, GTA_LLH_MINBLOCKS)
gat_aggregate_llh_kernel(const WorkList wl, const Exchange ex, const float* __restrict__ el,
                         const float* __restrict__ er, int64_t lder, int heads, float slope,
                         const float* __restrict__ z, const uint32_t row_bytes, float* __restrict__ out, int64_t ldo,
                         int f, int epilogue, float* __restrict__ rowmax, float* __restrict__ rowsum,
                         const uint32_t* er_stats, int stats_pitch, int64_t col_block) {
  constexpr int kU = kLlhUnroll / HPL > 2 ? kLlhUnroll / HPL : 2;      // edges per softmax group: e / p are HPL wide
  __shared__ uint32_t s_id[kAggWarps][32];
  __shared__ BoundCache<32> s_bound[kAggWarps][32 / LANES];          // er_stats are only passed for heads <= 32
  if (ex.world > 1 && blockIdx.y == 0 && blockIdx.x < ex.copy_ctas) {
    exchange_pull(ex, blockIdx.x);          // these CTAs move the peers' slots; everybody else reduces
    return;
  }
  const int lane = threadIdx.x & 31;
  const int l = lane & (LANES - 1);
  BoundCache<32>* bc = &s_bound[threadIdx.x >> 5][lane / LANES];
  if (l == 0) { bc->cb0 = -1; bc->cb1 = -1; bc->ok = 0; }
  __syncwarp();
  const BlockDivider block_of(col_block);
  const int fo = blockIdx.y * 128 + 4 * l;
  const int d = f / heads;          // HPL == 1: a multiple of 4;  HPL == 2: 2;  HPL == 4: 1
  const int head = (fo < f) ? fo / d : 0;          // the lane's first head (a multiple of HPL)
  const float* erh = er + head;
  const uint32_t er_bytes = uint32_t(lder) * 4u;
  uint32_t* sid = s_id[threadIdx.x >> 5];
  const uint32_t* mine = sid + (lane & ~(LANES - 1));
  int32_t* counter = wl.work_counter + blockIdx.y;
  int32_t* flags = wl.chain_flags + int64_t(blockIdx.y) * wl.num_slots;
  const uint64_t pol_stream = wl.pol_stream, pol_keep = wl.pol_keep;
  const int pstride = gat_partial_stride(f, heads);
  const int stats = f + int(blockIdx.y) * gat_stats_stride(heads);

  ItemCursor cur = cursor_begin<LANES>(wl, ex, counter, lane);
  int32_t landed = 0;          // exchange_gate: highest peer slot this lane has seen complete
  while (cur.first < wl.num_items) {
    const int64_t group = int64_t(cur.first) + lane / LANES;
    const bool have = group < wl.num_items;
    const bool active = have && fo < f;
    const int4 it = have ? __ldg(wl.items + group) : make_int4(0, 0, 0, -1);
    const int count = have ? it.z : 0;
    const int max_count = (LANES == 32) ? count : warp_max_i32(count);
    const int32_t* idx_base = wl.indices + it.y;
    const float* zf = z + (active ? fo : 0);
    float elh[HPL], m[HPL], s[HPL];
#pragma unroll
    for (int k = 0; k < HPL; ++k) { elh[k] = 0.f; m[k] = -INFINITY; s[k] = 0.f; }
    if (active) ldg_vec<HPL>(el + int64_t(it.x) * heads + head, elh);

    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    int idx_nxt = 0;
    if (l < count) idx_nxt = ld_stream_i32(idx_base + l, pol_stream);
    const int last_src = (count > 0 && (ex.world > 1 || er_stats != nullptr)) ? __ldg(idx_base + count - 1) : 0;
    if (ex.world > 1) {
      if (l == 0 && count > 0) exchange_gate(ex, last_src, landed);
      __syncwarp();
    }
    // bound path (see gat_aggregate_kernel): a lane only needs the bound of its own heads; the choice is
    // per lane group here, nothing below synchronises across groups on it
    bool bounded = false;
    int first_src = 0;
    if (er_stats != nullptr) first_src = __shfl_sync(0xffffffffu, idx_nxt, lane & ~(LANES - 1));
    if (er_stats != nullptr && count > 0) {          // every head of the block must pass: lanes of one item agree
      bound_lookup<LANES, 32>(bc, er_stats, block_of(first_src), block_of(last_src), stats_pitch, heads, l, group_mask<LANES>(lane));
      bounded = bc->ok != 0;
#pragma unroll
      for (int k = 0; k < HPL; ++k) m[k] = bounded ? leaky(elh[k] + bc->hi[head + k], slope) : -INFINITY;
    }
    for (int base = 0; base < max_count; base += LANES) {
      int n = count - base;
      n = n < 0 ? 0 : (n > LANES ? LANES : n);
      sid[lane] = uint32_t(idx_nxt);
      if (base + LANES + l < count) idx_nxt = ld_stream_i32(idx_base + base + LANES + l, pol_stream);
      __syncwarp();
      const int nmax = (LANES == 32) ? n : LANES;
#pragma unroll 1
      for (int j = 0; j < nmax; j += kU) {
        float e[kU][HPL];
        float4 v[kU];
        uint32_t id[kU];
#pragma unroll
        for (int u = 0; u < kU; ++u) id[u] = (j + u < LANES) ? mine[(j + u) & (LANES - 1)] : 0u;
#pragma unroll
        for (int u = 0; u < kU; ++u) {
          const bool ok = (j + u) < n;
          float erv[HPL];
#pragma unroll
          for (int k = 0; k < HPL; ++k) erv[k] = 0.f;
          if (ok) ldg_vec<HPL>(row_ptr(erh, id[u], er_bytes), erv);
#pragma unroll
          for (int k = 0; k < HPL; ++k) e[u][k] = ok ? leaky(elh[k] + erv[k], slope) : -INFINITY;
        }
#pragma unroll
        for (int u = 0; u < kU; ++u) {
          v[u] = make_float4(0.f, 0.f, 0.f, 0.f);
          if (active && (j + u) < n) v[u] = ld_row_f32x4(row_ptr(zf, id[u], row_bytes), pol_keep);
        }
        if (!bounded) {
          // online softmax: new running maximum per head, rescale what has been accumulated.  A head that has
          // seen no edge yet (mn = -inf) keeps its zeros.
          float sc[HPL];
#pragma unroll
          for (int k = 0; k < HPL; ++k) {
            float bm = e[0][k];
#pragma unroll
            for (int u = 1; u < kU; ++u) bm = fmaxf(bm, e[u][k]);
            const float mn = fmaxf(m[k], bm);
            sc[k] = (mn == -INFINITY) ? 1.f : expf(m[k] - mn);          // m = -inf on the first group: sc = 0, acc and s are 0 anyway
            s[k] *= sc[k];
            m[k] = mn;
          }
          acc.x *= sc[head_of<HPL>(0)]; acc.y *= sc[head_of<HPL>(1)];
          acc.z *= sc[head_of<HPL>(2)]; acc.w *= sc[head_of<HPL>(3)];
        }
#pragma unroll
        for (int u = 0; u < kU; ++u) {
          // e = -inf for the padding of the last group: p = 0.  ex2.approx path: the argument is <= 0 and terms
          // that matter have small |e - m|; relative error < 2e-6, inside the 1e-5 tolerance
          float p[HPL];
#pragma unroll
          for (int k = 0; k < HPL; ++k) {
            p[k] = (m[k] == -INFINITY) ? 0.f : __expf(e[u][k] - m[k]);
            s[k] += p[k];
          }
          acc.x = fmaf(p[head_of<HPL>(0)], v[u].x, acc.x); acc.y = fmaf(p[head_of<HPL>(1)], v[u].y, acc.y);
          acc.z = fmaf(p[head_of<HPL>(2)], v[u].z, acc.z); acc.w = fmaf(p[head_of<HPL>(3)], v[u].w, acc.w);
        }
      }
      __syncwarp();
    }
    // who publishes a head's statistics: the first lane of the head (width >= 4), or the one lane that owns it
    const bool head_leader = active && (HPL > 1 || (fo % d) == 0);
    const bool chained = have && it.w >= 0;
    for_groups_in_order<LANES>(lane, chained, [&]() {
      bool last = true;
      if (chained) {
        const int s0 = __ldg(wl.row_slots + it.x), s1 = __ldg(wl.row_slots + it.x + 1);
        last = it.w == s1 - 1;
        if (it.w != s0) {
          if (l == 0) chain_wait(flags + it.w - 1);
          __syncwarp(group_mask<LANES>(lane));
          if (active) {
            const float* prev = wl.partials + int64_t(it.w - 1) * pstride;
            float a[HPL], b[HPL];
#pragma unroll
            for (int k = 0; k < HPL; ++k) {
              const float pm = ld_state_f32(prev + stats + head + k), ps = ld_state_f32(prev + stats + heads + head + k);
              const float mn = fmaxf(pm, m[k]);
              const float t = (mn == -INFINITY) ? 0.f : expf(fminf(pm, m[k]) - mn);      // the other factor is exp(0) = 1
              a[k] = (pm == -INFINITY) ? 0.f : (pm == mn ? 1.f : t);
              b[k] = (m[k] == -INFINITY) ? 0.f : (m[k] == mn ? 1.f : t);
              s[k] = fmaf(ps, a[k], s[k] * b[k]);
              m[k] = mn;
            }
            const float4 p = ld_state_f32x4(prev + fo);
            acc.x = fmaf(p.x, a[head_of<HPL>(0)], acc.x * b[head_of<HPL>(0)]);
            acc.y = fmaf(p.y, a[head_of<HPL>(1)], acc.y * b[head_of<HPL>(1)]);
            acc.z = fmaf(p.z, a[head_of<HPL>(2)], acc.z * b[head_of<HPL>(2)]);
            acc.w = fmaf(p.w, a[head_of<HPL>(3)], acc.w * b[head_of<HPL>(3)]);
          }
        }
        if (!last) {
          float* part = wl.partials + int64_t(it.w) * pstride;
          if (active) *reinterpret_cast<float4*>(part + fo) = acc;
          if (head_leader) {
#pragma unroll
            for (int k = 0; k < HPL; ++k) {
              part[stats + head + k] = m[k];
              part[stats + heads + head + k] = s[k];
            }
          }
          __syncwarp(group_mask<LANES>(lane));          // the group's stores happen-before lane 0's release
          if (l == 0) chain_publish(flags + it.w);
        }
      }
      if (last && active) {
        float inv[HPL];
#pragma unroll
        for (int k = 0; k < HPL; ++k) inv[k] = s[k] > 0.f ? 1.f / s[k] : 0.f;
        st_stream_f32x4(out + int64_t(it.x) * ldo + fo,
                        make_float4(apply_epilogue(acc.x * inv[head_of<HPL>(0)], epilogue), apply_epilogue(acc.y * inv[head_of<HPL>(1)], epilogue),
                                    apply_epilogue(acc.z * inv[head_of<HPL>(2)], epilogue), apply_epilogue(acc.w * inv[head_of<HPL>(3)], epilogue)));
        if (head_leader) {
#pragma unroll
          for (int k = 0; k < HPL; ++k) {
            if (rowmax) rowmax[int64_t(it.x) * heads + head + k] = (count > 0 || it.w >= 0) && m[k] != -INFINITY ? m[k] : 0.f;
            if (rowsum) rowsum[int64_t(it.x) * heads + head + k] = s[k];
          }
        }
      }
    });
    cursor_next<LANES>(cur, wl, counter, lane);
  }
}

// ----------------------------------------------------------------------------------------
// Edge phase "sum of three, then a unary, then the row sum" in one pass (PNA ops 5-8,
// genGraphOP.py:110-147:  gather_R( SF( edge + scatterC(a) + scatterR(b) ) )):
//     out[i, :] = epilogue( sum_{k in row i} unary( edge[k, :] + x[src_k, :] + rowterm[i, :] ) )
// Any of the three terms may be absent.  The generic path materialises three E x F tensors (two adds and the
// unary) before the segment sum; here the E x F operand is read once, streaming, and nothing E x F is written.
// Same work list, dynamic item fetch and slot chain as aggregate_kernel; the reduction is the plain ascending
// edge order per lane.  A lane moves one 16-byte piece of every row it touches.
// ----------------------------------------------------------------------------------------
__device__ __forceinline__ float edge_unary_value(float v, int unary, float slope) {
  if (unary == GTA_UN_RELU) return fmaxf(v, 0.f);
  if (unary == GTA_UN_ELU) return elu1(v);
  if (unary == GTA_UN_EXP_LEAKY_RELU) return expf(leaky(v, slope));
  return v;
}
template <int LANES, bool HAS_X, bool HAS_E>
__global__ void __launch_bounds__(kAggThreads, 8)
edge_sum_kernel(const WorkList wl, const Exchange ex, const float* __restrict__ edge, int64_t lde,
                const float* __restrict__ x, const uint32_t row_bytes, const float* __restrict__ rowterm, int64_t ldr,
                int unary, float slope, float* __restrict__ out, int64_t ldo, int f, int epilogue) {
  constexpr int kU = 4;
  __shared__ uint32_t s_id[kAggWarps][32];
  const int lane = threadIdx.x & 31;
  const int l = lane & (LANES - 1);
  const int fo = blockIdx.y * (LANES * 4) + 4 * l;
  uint32_t* sid = s_id[threadIdx.x >> 5];
  const uint32_t* mine = sid + (lane & ~(LANES - 1));
  int32_t* counter = wl.work_counter + blockIdx.y;
  int32_t* flags = wl.chain_flags + int64_t(blockIdx.y) * wl.num_slots;
  const uint64_t pol_stream = wl.pol_stream, pol_keep = wl.pol_keep;

  ItemCursor cur = cursor_begin<LANES>(wl, ex, counter, lane);
  while (cur.first < wl.num_items) {
    const int64_t group = int64_t(cur.first) + lane / LANES;
    const bool have = group < wl.num_items;
    const bool active = have && fo < f;
    const int4 it = have ? __ldg(wl.items + group) : make_int4(0, 0, 0, -1);
    const int count = have ? it.z : 0;
    const int max_count = (LANES == 32) ? count : warp_max_i32(count);
    const int32_t* idx_base = wl.indices + it.y;
    const float* xf = HAS_X ? x + (active ? fo : 0) : nullptr;
    const float* ef = HAS_E ? edge + int64_t(it.y) * lde + (active ? fo : 0) : nullptr;
    float4 r4 = make_float4(0.f, 0.f, 0.f, 0.f);
    if (active && rowterm != nullptr) r4 = __ldg(reinterpret_cast<const float4*>(rowterm + int64_t(it.x) * ldr + fo));
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    int idx_nxt = 0;
    if (HAS_X && l < count) idx_nxt = ld_stream_i32(idx_base + l, pol_stream);
    for (int base = 0; base < max_count; base += LANES) {
      int n = count - base;
      n = n < 0 ? 0 : (n > LANES ? LANES : n);
      if (HAS_X) {
        sid[lane] = uint32_t(idx_nxt);
        if (base + LANES + l < count) idx_nxt = ld_stream_i32(idx_base + base + LANES + l, pol_stream);
        __syncwarp();
      }
      const int nmax = (LANES == 32) ? n : LANES;
#pragma unroll 1
      for (int j = 0; j < nmax; j += kU) {
        float4 xv[kU], ev[kU];
#pragma unroll
        for (int u = 0; u < kU; ++u) {
          const bool ok = active && (j + u) < n;
          xv[u] = make_float4(0.f, 0.f, 0.f, 0.f);
          ev[u] = make_float4(0.f, 0.f, 0.f, 0.f);
          if (HAS_X && ok) xv[u] = ld_row_f32x4(row_ptr(xf, mine[(j + u) & (LANES - 1)], row_bytes), pol_keep);
          if (HAS_E && ok) ev[u] = ld_gather_f32x4(ef + int64_t(base + j + u) * lde, pol_stream);          // streamed: read once, evict first
        }
#pragma unroll
        for (int u = 0; u < kU; ++u) {
          if ((j + u) < n) {          // a padded slot would contribute unary(rowterm), not 0
            acc.x += edge_unary_value(ev[u].x + xv[u].x + r4.x, unary, slope);
            acc.y += edge_unary_value(ev[u].y + xv[u].y + r4.y, unary, slope);
            acc.z += edge_unary_value(ev[u].z + xv[u].z + r4.z, unary, slope);
            acc.w += edge_unary_value(ev[u].w + xv[u].w + r4.w, unary, slope);
          }
        }
      }
      if (HAS_X) __syncwarp();
    }
    const bool chained = have && it.w >= 0;
    for_groups_in_order<LANES>(lane, chained, [&]() {
      bool last = true;
      if (chained) {
        const int s0 = __ldg(wl.row_slots + it.x), s1 = __ldg(wl.row_slots + it.x + 1);
        last = it.w == s1 - 1;
        if (it.w != s0) {
          if (l == 0) chain_wait(flags + it.w - 1);
          __syncwarp(group_mask<LANES>(lane));
          if (active) {
            const float4 p = ld_state_f32x4(wl.partials + int64_t(it.w - 1) * f + fo);
            acc.x = p.x + acc.x; acc.y = p.y + acc.y; acc.z = p.z + acc.z; acc.w = p.w + acc.w;
          }
        }
        if (!last) {
          if (active) *reinterpret_cast<float4*>(wl.partials + int64_t(it.w) * f + fo) = acc;
          __syncwarp(group_mask<LANES>(lane));          // the group's stores happen-before lane 0's release
          if (l == 0) chain_publish(flags + it.w);
        }
      }
      if (last && active) st_stream_f32x4(out + int64_t(it.x) * ldo + fo, epilogue4(acc, 1.f, epilogue));
    });
    cursor_next<LANES>(cur, wl, counter, lane);
  }
}

// ----------------------------------------------------------------------------------------
// Roofline denominator of the gather kernels: random whole-row gathers from a table that fits L2, with the
// kernels' own load instruction (one 128-bit load per lane, L1 no-allocate, L2 evict_last), 8 in flight per
// lane, ids from a hash so nothing else touches memory.  bench.py reports gather bytes / time of the real
// kernel against this measured peak (roofline.l2_frac): the HBM roofline says little about a kernel whose
// 96 % of the traffic is L2 hits.
// ----------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kAggThreads, 8)
gather_peak_kernel(const float* __restrict__ table, uint32_t rows, uint32_t row_bytes, int lanes_per_row,
                   int64_t gathers_per_group, float4* __restrict__ sink, const uint64_t pol_keep) {
  const int lane = threadIdx.x & 31;
  const int l = lane % lanes_per_row;
  const uint64_t group = (uint64_t(blockIdx.x) * blockDim.x + threadIdx.x) / lanes_per_row;
  const float* base = table + 4 * l;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  uint32_t state = uint32_t(group * 2654435761u) | 1u;
  for (int64_t i = 0; i < gathers_per_group; i += 8) {
    float4 v[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      state = state * 1664525u + 1013904223u;            // LCG: the lanes of a group draw the same ids
      const uint32_t id = uint32_t((uint64_t(state) * rows) >> 32);
      v[u] = ld_row_f32x4(row_ptr(base, id, row_bytes), pol_keep);
    }
#pragma unroll
    for (int u = 0; u < 8; ++u) { acc.x += v[u].x; acc.y += v[u].y; acc.z += v[u].z; acc.w += v[u].w; }
  }
  if (acc.x == 1.2345e38f) sink[group] = acc;          // keeps the loads alive, never true in practice
}

// ----------------------------------------------------------------------------------------
// er_stats: per column block and head, max er and max -er as ordered-int codes (atomicMax on zeroed words)
// ----------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
er_stats_kernel(const float* __restrict__ er, int64_t lder, int64_t num_sources, int64_t col_block, int heads,
                uint32_t* __restrict__ stats) {
  // lane -> head (heads is a power of two <= 32), 32/heads rows per warp step
  const int lane = threadIdx.x & 31;
  const int h = lane & (heads - 1);
  const int rows_per_step = 32 / heads;
  const int64_t cb = blockIdx.y;
  const int64_t lo = col_block > 0 ? cb * col_block : 0;
  const int64_t hi = col_block > 0 ? (lo + col_block < num_sources ? lo + col_block : num_sources) : num_sources;
  const int64_t warp = (blockIdx.x * int64_t(blockDim.x) + threadIdx.x) >> 5;
  const int64_t warps = (int64_t(gridDim.x) * blockDim.x) >> 5;
  float mx = -INFINITY, mn = INFINITY;
  bool seen = false;
  for (int64_t r = lo + warp * rows_per_step + lane / heads; r < hi; r += warps * rows_per_step) {
    const float v = __ldg(er + r * lder + h);
    mx = fmaxf(mx, v);
    mn = fminf(mn, v);
    seen = true;
  }
  // lanes with the same head: xor offsets heads, 2*heads, ...
  for (int o = heads; o < 32; o <<= 1) {
    mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, o));
    seen = __shfl_xor_sync(0xffffffffu, int(seen), o) || seen;
  }
  if (lane < heads && seen) {
    atomicMax(stats + (cb * 2) * heads + h, ordered_code(mx));
    atomicMax(stats + (cb * 2 + 1) * heads + h, ordered_code(-mn));
  }
}

// ----------------------------------------------------------------------------------------
// GAT block [4,5,6,7,8]: numerators p[E,H] (STORE_E) and row sums S[N,H]; warp per row
// ----------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
gat_logits_kernel(const int64_t* __restrict__ indptr, const int32_t* __restrict__ indices, int64_t row_begin,
                  int64_t row_end, const float* __restrict__ el, const float* __restrict__ er, int heads,
                  float slope, int stabilize, float* __restrict__ p, float* __restrict__ rowmax,
                  float* __restrict__ rowsum) {
  const int lane = threadIdx.x & 31;
  const int64_t r = row_begin + ((blockIdx.x * int64_t(blockDim.x) + threadIdx.x) >> 5);
  if (r >= row_end) return;
  const int64_t b = indptr[r], e = indptr[r + 1];
  const int64_t lr = r - row_begin;
  for (int h = 0; h < heads; ++h) {
    const float elv = el[lr * heads + h];
    float mx = -INFINITY;
    if (stabilize) {
      for (int64_t k = b + lane; k < e; k += 32)
        mx = fmaxf(mx, leaky(elv + er[int64_t(indices[k]) * heads + h], slope));
      mx = group_max<32>(mx);
    }
    if (!stabilize || mx == -INFINITY) mx = 0.f;
    // deterministic sum: fixed lane-strided partial sums, then a fixed butterfly
    float sum = 0.f;
    for (int64_t k = b + lane; k < e; k += 32) {
      float v = expf(leaky(elv + er[int64_t(indices[k]) * heads + h], slope) - mx);
      p[k * heads + h] = v;
      sum += v;
    }
    sum = group_sum<32>(sum);
    if (lane == 0) {
      if (rowmax) rowmax[lr * heads + h] = mx;
      rowsum[lr * heads + h] = sum;
    }
  }
}

// lanes per item for rows of f features, kper features per 16-byte piece, v pieces per lane
static int lanes_for(int f, int kper = 4, int v = 1) {
  const int window = 32 * kper * v;
  int need = ((f < window ? f : window) + kper * v - 1) / (kper * v);
  int l = 1;
  while (l < need) l <<= 1;
  return l < 4 ? 4 : l;
}

template <typename K>
static int resident_ctas(K kernel) {
  static int cached = 0;          // one static per kernel instantiation
  if (cached > 0) return cached;
  int per_sm = 0, dev = 0, sms = kNumSMs;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, kAggThreads, 0) != cudaSuccess || per_sm < 1) per_sm = 1;
  if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  cached = per_sm * sms;
  return cached;
}

// dynamic item fetch (take = 1) unless the caller asked for static striding (GTA_PHASE_STATIC: long lists of
// tiny items) and the list is long enough for every resident warp to get a few dozen group-steps
template <typename K>
static int32_t take_for(K kernel, const WorkList& wl, int lanes) {
  const int groups = 32 / lanes;
  const int64_t warps = int64_t(resident_ctas(kernel)) * (kAggThreads / 32);
  return (wl.take == 0 && wl.num_items / (warps * groups) >= 32) ? 0 : 1;      // wl.take == 0: the caller's hint
}

template <typename K>
static dim3 persistent_grid(K kernel, const WorkList& wl, int lanes, int f, const Exchange& ex) {
  const int64_t num_items = wl.num_items;
  const int64_t need = (num_items * lanes + kAggThreads - 1) / kAggThreads;
  int64_t cap = resident_ctas(kernel);
  // the copy CTAs of an exchange come first in the grid, so they are resident before any CTA can wait on them;
  // with static striding every work CTA must be resident too (a chain may wait on any of them)
  const int64_t copy = ex.world > 1 ? ex.copy_ctas : 0;
  if (take_for(kernel, wl, lanes) == 0 && cap > copy + 1) cap -= copy;
  return dim3((unsigned)((need < cap ? need : cap) + copy), (unsigned)((f + 127) / 128));
}

static WorkList with_take(WorkList wl, int32_t take) {
  wl.take = take;
  return wl;
}

// gta_exchange_t (host) -> Exchange (kernel parameter); arrived[] lives behind the item counters
static int make_exchange(const char* who, const gta_exchange_t* h, int32_t* arrived, int64_t pitch_bytes, Exchange* ex) {
  memset(ex, 0, sizeof(*ex));
  if (h == nullptr || h->world <= 1) return GTA_OK;
  GTA_REQUIRE(h->world <= GTA_MAX_RANKS && h->rank >= 0 && h->rank < h->world && h->step >= 1,
              "%s: exchange world %d rank %d step %d", who, h->world, h->rank, h->step);
  GTA_REQUIRE(h->table && h->signals && h->slot_rows > 0, "%s: exchange table / signals / slot_rows missing", who);
  GTA_REQUIRE(h->row_bytes == pitch_bytes && h->row_bytes % 16 == 0,
              "%s: exchange row_bytes %lld does not match the table's row pitch %lld", who, (long long)h->row_bytes,
              (long long)pitch_bytes);
  GTA_REQUIRE((h->slot_rows * h->row_bytes) % 128 == 0,
              "%s: a slot (%lld rows of %lld bytes) must be a whole number of 128-byte lines", who,
              (long long)h->slot_rows, (long long)h->row_bytes);
  ex->world = h->world;
  ex->copy_ctas = h->copy_ctas > 0 ? h->copy_ctas : 96;
  ex->step = h->step;
  ex->row_bytes = uint32_t(h->row_bytes);
  ex->slot_rows = h->slot_rows;
  ex->table = static_cast<char*>(h->table);
  ex->signals = static_cast<const ExchangeSignals*>(h->signals);
  ex->arrived = arrived;
  for (int k = 0; k < h->world; ++k) {
    GTA_REQUIRE(k == 0 || h->peer_table[k], "%s: table of slot %d's owner is not mapped", who, k);
    GTA_REQUIRE(h->slot_valid_rows[k] >= 0 && h->slot_valid_rows[k] <= h->slot_rows, "%s: slot %d has %lld rows", who, k,
                (long long)h->slot_valid_rows[k]);
    ex->peer[k] = static_cast<const char*>(h->peer_table[k]);
    ex->valid_rows[k] = int32_t(h->slot_valid_rows[k]);
  }
  return GTA_OK;
}

template <typename P, int V, int LANES>
static void dispatch_aggregate(int wkind, bool div, cudaStream_t st, const WorkList& wl, const Exchange& ex,
                               const float* w, int wh, const float* rowden, const typename P::T* x, int64_t ldx,
                               float* out, int64_t ldo, int f, int epi) {
  using T = typename P::T;
  constexpr int kWin = LANES * P::kPer * V;
#define GTA_AGG(K, D)                                                                                               \
  do {                                                                                                              \
    auto kern = aggregate_kernel<P, V, LANES, K, D>;                                                                \
    dim3 grid = persistent_grid(kern, wl, LANES, 1, ex);                                                  \
    grid.y = (unsigned)((f + kWin - 1) / kWin);                                                                     \
    kern<<<grid, kAggThreads, 0, st>>>(with_take(wl, take_for(kern, wl, LANES)), ex, w, wh, rowden, x,    \
                                       uint32_t(ldx * sizeof(T)), out, ldo, f, epi);                                \
  } while (0)
  if (wkind == 0) GTA_AGG(0, false);
  else if (wkind == 1 && !div) GTA_AGG(1, false);
  else if (wkind == 1 && div) GTA_AGG(1, true);
  else if constexpr (V == 1) {
    if (!div) GTA_AGG(2, false);
    else GTA_AGG(2, true);
  }
#undef GTA_AGG
}

template <typename P, int H>
static int dispatch_gat(int lanes, cudaStream_t st, const WorkList& wl, const Exchange& ex, const float* el,
                        const float* er, int64_t lder, float slope, const typename P::T* z, int64_t ldz, float* out, int64_t ldo,
                        int f, int epi, float* rowmax, float* rowsum, const uint32_t* er_stats, int stats_pitch,
                        int64_t col_block) {
#define GTA_GAT(L)                                                                                                  \
  do {                                                                                                              \
    auto kern = gat_aggregate_kernel<P, L, H>;                                                                      \
    dim3 grid = persistent_grid(kern, wl, L, 1, ex);                                                      \
    grid.y = (unsigned)((f + L * P::kPer - 1) / (L * P::kPer));                                                     \
    kern<<<grid, kAggThreads, 0, st>>>(with_take(wl, take_for(kern, wl, L)), ex, el, er, lder, slope, z,  \
                                       uint32_t(ldz * sizeof(typename P::T)), out, ldo, f, epi, rowmax, rowsum,     \
                                       er_stats,                                                                    \
                                       stats_pitch, col_block);                                                     \
  } while (0)
  switch (lanes) {
    case 4: if (H <= 4) { GTA_GAT(4); return GTA_OK; } break;
    case 8: if (H <= 8) { GTA_GAT(8); return GTA_OK; } break;
    case 16: GTA_GAT(16); return GTA_OK;
    case 32: GTA_GAT(32); return GTA_OK;
  }
#undef GTA_GAT
  return GTA_ERR_UNSUPPORTED;
}

// common argument checks, the RESET phase (clear the chain flags of every feature window) and the item
// counters (cleared before every launch).  chain_state = [windows][num_slots] flags, then [windows] item
// counters, then GTA_MAX_RANKS slot-arrival counters of an exchange.
static int prepare_worklist(const char* who, WorkList& wl, int32_t* chain_state, int32_t f, int32_t phases,
                            cudaStream_t st) {
  GTA_REQUIRE(chain_state, "%s: chain_state is required (chain flags and the item counters live there)", who);
  GTA_REQUIRE(wl.num_slots == 0 || (wl.partials && wl.row_slots),
              "%s: partials and row_slots are required for %lld slots", who, (long long)wl.num_slots);
  const size_t windows = size_t((f + 127) / 128);
  wl.chain_flags = chain_state;
  wl.work_counter = chain_state + windows * size_t(wl.num_slots);
  if ((phases & GTA_PHASE_RESET) && wl.num_slots > 0) {
    GTA_CUDA(cudaMemsetAsync(wl.chain_flags, 0, windows * size_t(wl.num_slots) * sizeof(int32_t), st));
    count_launch();
  }
  wl.take = (phases & GTA_PHASE_STATIC) ? 0 : 1;
  if ((phases & GTA_PHASE_MAIN) && wl.num_items > 0) {
    GTA_CUDA(cudaMemsetAsync(wl.work_counter, 0, (windows + GTA_MAX_RANKS) * sizeof(int32_t), st));
    count_launch();
    CachePolicies pol;
    int rc = cache_policies(&pol);
    if (rc != GTA_OK) return rc;
    wl.pol_stream = pol.stream;
    wl.pol_keep = pol.keep;
  }
  return GTA_OK;
}

// ---- the two aggregation entry points, for either storage type of the gathered table ---------------
template <typename P>
static int aggregate_run(const char* who, int wkind, bool div, cudaStream_t st, const WorkList& wl, const Exchange& ex,
                         const float* w, int wh, const float* rowden, const typename P::T* x, int64_t ldx, float* out,
                         int64_t ldo, int f, int epilogue) {
  constexpr int KP = P::kPer;
  GTA_REQUIRE(f % KP == 0, "%s: f=%d must be a multiple of %d (pad the table)", who, f, KP);
  if (wkind == 2 && (f / wh) % KP != 0) {
    set_error("%s: per-head width f/wh=%d is not a multiple of %d", who, f / wh, KP);
    return GTA_ERR_UNSUPPORTED;
  }
  // rows wider than one 32-lane pass of single pieces: two pieces per lane (one walk of the work list, not two)
  if (f > 32 * KP && wkind != 2) {
    dispatch_aggregate<P, 2, 32>(wkind, div, st, wl, ex, w, wh, rowden, x, ldx, out, ldo, f, epilogue);
  } else {
    switch (lanes_for(f, KP)) {
      case 4: dispatch_aggregate<P, 1, 4>(wkind, div, st, wl, ex, w, wh, rowden, x, ldx, out, ldo, f, epilogue); break;
      case 8: dispatch_aggregate<P, 1, 8>(wkind, div, st, wl, ex, w, wh, rowden, x, ldx, out, ldo, f, epilogue); break;
      case 16: dispatch_aggregate<P, 1, 16>(wkind, div, st, wl, ex, w, wh, rowden, x, ldx, out, ldo, f, epilogue); break;
      default: dispatch_aggregate<P, 1, 32>(wkind, div, st, wl, ex, w, wh, rowden, x, ldx, out, ldo, f, epilogue); break;
    }
  }
  return GTA_OK;
}

template <typename T>
static int aggregate_impl(const char* who, const int32_t* items_, int64_t num_items, const int32_t* row_slots,
                          int64_t num_slots, const int32_t* indices, int32_t wmode, const float* w, int32_t wh,
                          const float* rowden, const T* x, int64_t ldx, float* out, int64_t ldo, int32_t f,
                          int32_t epilogue, float* partials, int32_t* chain_state, const gta_exchange_t* exchange,
                          int32_t phases, void* stream_) {
  constexpr int kRow = 16 / int(sizeof(T));          // elements per 16 bytes: the row pitch granule
  cudaStream_t st = static_cast<cudaStream_t>(stream_);
  GTA_REQUIRE(f > 0 && f % 4 == 0, "%s: f=%d must be a positive multiple of 4 (pad the table)", who, f);
  GTA_REQUIRE(num_items >= 0 && num_items < (int64_t(1) << 31) - 64, "%s: bad item count", who);
  WorkList wl{reinterpret_cast<const int4*>(items_), num_items, row_slots, num_slots, indices, partials, nullptr, nullptr, 1, 0, 0};
  int rc = prepare_worklist(who, wl, chain_state, f, phases, st);
  if (rc != GTA_OK) return rc;
  if (num_items == 0 || !(phases & GTA_PHASE_MAIN)) return GTA_OK;
  GTA_REQUIRE(items_ && indices && x && out, "%s: null pointer", who);
  GTA_REQUIRE(ldx % kRow == 0 && ldo % 4 == 0 && ldx >= f && ldo >= f && ldx * int64_t(sizeof(T)) < (int64_t(1) << 32),
              "%s: leading dimensions must be whole 16-byte pieces, >= f, and a row below 4 GiB", who);
  GTA_REQUIRE((reinterpret_cast<uintptr_t>(x) & 15) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0, "%s: tables must be 16-byte aligned", who);
  GTA_REQUIRE(wmode >= GTA_W_NONE && wmode <= GTA_W_EDGE_DIV, "%s: bad wmode %d", who, wmode);
  int wkind = 0;
  bool div = wmode == GTA_W_EDGE_DIV;
  if (wmode != GTA_W_NONE) {
    GTA_REQUIRE(w && wh >= 1 && f % wh == 0, "%s: weight width %d must divide f=%d", who, wh, f);
    GTA_REQUIRE(!div || rowden, "%s: rowden required for GTA_W_EDGE_DIV", who);
    wkind = wh == 1 ? 1 : 2;
  }
  Exchange ex;
  rc = make_exchange(who, exchange, wl.work_counter + (f + 127) / 128, ldx * int64_t(sizeof(T)), &ex);
  if (rc != GTA_OK) return rc;
  GTA_REQUIRE(ex.world <= 1 || ex.table == reinterpret_cast<const char*>(x), "%s: x is not the exchange table", who);
  if constexpr (sizeof(T) == 4) {
    rc = aggregate_run<F32x4>(who, wkind, div, st, wl, ex, w, wh, rowden, x, ldx, out, ldo, f, epilogue);
  } else if (f <= 128) {
    rc = aggregate_run<Bf16x4>(who, wkind, div, st, wl, ex, w, wh, rowden, x, ldx, out, ldo, f, epilogue);
  } else {
    rc = aggregate_run<Bf16x8>(who, wkind, div, st, wl, ex, w, wh, rowden, x, ldx, out, ldo, f, epilogue);
  }
  if (rc != GTA_OK) return rc;
  GTA_CHECK_LAUNCH("aggregate_kernel");
  return GTA_OK;
}

template <typename P>
static int gat_run(const char* who, int heads, cudaStream_t st, const WorkList& wl, const Exchange& ex, const float* el,
                   const float* er, int64_t lder, float slope, const typename P::T* z, int64_t ldz, float* out, int64_t ldo,
                   int f, int epilogue, float* rowmax, float* rowsum, const uint32_t* er_stats, int stats_pitch,
                   int64_t col_block) {
  constexpr int KP = P::kPer;
  if (f % KP != 0 || (f / heads) % KP != 0) {
    set_error("%s: per-head width f/heads=%d is not a multiple of %d", who, f / heads, KP);
    return GTA_ERR_UNSUPPORTED;
  }
  const int lanes = lanes_for(f, KP);
  int rc = GTA_ERR_UNSUPPORTED;
#define GTA_GAT_H(HH) rc = dispatch_gat<P, HH>(lanes, st, wl, ex, el, er, lder, slope, z, ldz, out, ldo, f, epilogue, rowmax, rowsum, er_stats, stats_pitch, col_block)
  switch (heads) {
    case 1: GTA_GAT_H(1); break;
    case 2: GTA_GAT_H(2); break;
    default: GTA_GAT_H(4); break;
  }
#undef GTA_GAT_H
  if (rc != GTA_OK) set_error("%s: no kernel for heads=%d, f=%d", who, heads, f);
  return rc;
}

template <typename T>
static int gat_aggregate_impl(const char* who, const int32_t* items_, int64_t num_items, const int32_t* row_slots,
                              int64_t num_slots, const int32_t* indices, const float* el, const float* er, int64_t lder,
                              int32_t heads, float slope, const T* z, int64_t ldz, float* out, int64_t ldo, int32_t f,
                              int32_t epilogue, float* rowmax, float* rowsum, float* partials, int32_t* chain_state,
                              const uint32_t* er_stats, int64_t col_block, const gta_exchange_t* exchange,
                              int32_t phases, void* stream_) {
  constexpr int kRow = 16 / int(sizeof(T));
  cudaStream_t st = static_cast<cudaStream_t>(stream_);
  GTA_REQUIRE(f > 0 && f % 4 == 0, "%s: f=%d must be a positive multiple of 4", who, f);
  GTA_REQUIRE(num_items >= 0 && num_items < (int64_t(1) << 31) - 64, "%s: bad item count", who);
  WorkList wl{reinterpret_cast<const int4*>(items_), num_items, row_slots, num_slots, indices, partials, nullptr, nullptr, 1, 0, 0};
  int rc = prepare_worklist(who, wl, chain_state, f, phases, st);
  if (rc != GTA_OK) return rc;
  if (num_items == 0 || !(phases & GTA_PHASE_MAIN)) return GTA_OK;
  GTA_REQUIRE(items_ && indices && el && er && z && out, "%s: null pointer", who);
  GTA_REQUIRE(ldz % kRow == 0 && ldo % 4 == 0 && ldz >= f && ldo >= f && ldz * int64_t(sizeof(T)) < (int64_t(1) << 32),
              "%s: leading dimensions must be whole 16-byte pieces, >= f, and a row below 4 GiB", who);
  GTA_REQUIRE((reinterpret_cast<uintptr_t>(z) & 15) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0 &&
              (reinterpret_cast<uintptr_t>(er) & 15) == 0 && (reinterpret_cast<uintptr_t>(el) & 15) == 0,
              "%s: tables must be 16-byte aligned", who);
  GTA_REQUIRE(heads >= 1 && f % heads == 0, "%s: heads=%d must divide f=%d", who, heads, f);
  GTA_REQUIRE(lder >= heads && (heads % 4 != 0 || lder % 4 == 0) && (heads % 2 != 0 || lder % 2 == 0),
              "%s: er row stride %lld breaks the vector alignment of %d heads", who, (long long)lder, heads);
  Exchange ex;
  rc = make_exchange(who, exchange, wl.work_counter + (f + 127) / 128, ldz * int64_t(sizeof(T)), &ex);
  if (rc != GTA_OK) return rc;
  int stats_pitch = 2 * heads;
  if (ex.world > 1) {
    GTA_REQUIRE(ex.table == reinterpret_cast<const char*>(z), "%s: z is not the exchange table", who);
    // the slot owners published their er range with the step; a slot's statistics are valid once it has landed
    er_stats = &ex.signals->stats[ex.step & 1][0][0];
    stats_pitch = 64;
    col_block = ex.slot_rows;
    if ((heads & (heads - 1)) != 0 || heads > 32) er_stats = nullptr;
  }
  // the bound path does not track the true row maximum: callers that want it back run the online softmax
  if (rowmax != nullptr) er_stats = nullptr;
  // H <= 4 with whole pieces per head: staged kernel (all heads per lane, softmax once per 32-edge batch);  H >= 8, an
  // unusual H or heads narrower than a piece: lane-local-head kernel (per-head width a multiple of 4, or 2, or 1;
  // constant register footprint; fp32 tables only)
  bool staged = !GTA_GAT_FORCE_LLH && (heads == 1 || heads == 2 || heads == 4);
  if constexpr (sizeof(T) == 4) staged = staged && (f / heads) % 4 == 0;
  if (staged) {
    if constexpr (sizeof(T) == 4) {
      rc = gat_run<F32x4>(who, heads, st, wl, ex, el, er, lder, slope, z, ldz, out, ldo, f, epilogue, rowmax, rowsum,
                          er_stats, stats_pitch, col_block);
    } else if (f <= 128) {
      rc = gat_run<Bf16x4>(who, heads, st, wl, ex, el, er, lder, slope, z, ldz, out, ldo, f, epilogue, rowmax, rowsum,
                           er_stats, stats_pitch, col_block);
    } else {
      rc = gat_run<Bf16x8>(who, heads, st, wl, ex, el, er, lder, slope, z, ldz, out, ldo, f, epilogue, rowmax, rowsum,
                           er_stats, stats_pitch, col_block);
    }
    if (rc != GTA_OK) return rc;
  } else if constexpr (sizeof(T) == 4) {
    const int width = f / heads;
    const int hpl = width % 4 == 0 ? 1 : (width == 2 ? 2 : (width == 1 ? 4 : 0));      // heads per 4-feature lane
    if (hpl == 0) {
      set_error("%s: per-head width f/heads=%d is neither a multiple of 4 nor 2 nor 1", who, width);
      return GTA_ERR_UNSUPPORTED;
    }
    const int lanes = lanes_for(f, 4);
#define GTA_LLH2(L, HP)                                                                                                 \
  gat_aggregate_llh_kernel<L, HP><<<persistent_grid(gat_aggregate_llh_kernel<L, HP>, wl, L, f, ex), kAggThreads, 0,      \
                                    st>>>(with_take(wl, take_for(gat_aggregate_llh_kernel<L, HP>, wl, L)), ex, el, er,   \
                                          lder, heads, slope, z, uint32_t(ldz) * 4u, out, ldo, f, epilogue, rowmax,      \
                                          rowsum, er_stats, stats_pitch, col_block)
#define GTA_LLH(L)                                                                                                      \
  do {                                                                                                                  \
    if (hpl == 1) GTA_LLH2(L, 1);                                                                                       \
    else if (hpl == 2) GTA_LLH2(L, 2);                                                                                  \
    else GTA_LLH2(L, 4);                                                                                                \
  } while (0)
    switch (lanes) {
      case 4: GTA_LLH(4); break;
      case 8: GTA_LLH(8); break;
      case 16: GTA_LLH(16); break;
      default: GTA_LLH(32); break;
    }
#undef GTA_LLH
#undef GTA_LLH2
  } else {
    set_error("%s: %d heads on a bf16 table has no kernel yet (fp32 tables: any head count)", who, heads);
    return GTA_ERR_UNSUPPORTED;
  }
  GTA_CHECK_LAUNCH("gat_aggregate_kernel");
  return GTA_OK;
}

}  // namespace gta

using namespace gta;

extern "C" {

int32_t gta_gat_partial_stride(int32_t f, int32_t heads) { return gat_partial_stride(f, heads); }

int gta_aggregate_f32(const int32_t* items, int64_t num_items, const int32_t* row_slots, int64_t num_slots,
                      const int32_t* indices, int32_t wmode, const float* w, int32_t wh, const float* rowden,
                      const float* x, int64_t ldx, float* out, int64_t ldo, int32_t f, int32_t epilogue,
                      float* partials, int32_t* chain_state, const gta_exchange_t* exchange, int32_t phases,
                      void* stream) {
  return aggregate_impl<float>("gta_aggregate_f32", items, num_items, row_slots, num_slots, indices, wmode, w, wh, rowden,
                               x, ldx, out, ldo, f, epilogue, partials, chain_state, exchange, phases, stream);
}

int gta_aggregate_bf16(const int32_t* items, int64_t num_items, const int32_t* row_slots, int64_t num_slots,
                       const int32_t* indices, int32_t wmode, const float* w, int32_t wh, const float* rowden,
                       const void* x, int64_t ldx, float* out, int64_t ldo, int32_t f, int32_t epilogue,
                       float* partials, int32_t* chain_state, const gta_exchange_t* exchange, int32_t phases,
                       void* stream) {
  return aggregate_impl<__nv_bfloat16>("gta_aggregate_bf16", items, num_items, row_slots, num_slots, indices, wmode, w, wh,
                                       rowden, static_cast<const __nv_bfloat16*>(x), ldx, out, ldo, f, epilogue, partials,
                                       chain_state, exchange, phases, stream);
}

int gta_aggregate_edge_sum_f32(const int32_t* items_, int64_t num_items, const int32_t* row_slots, int64_t num_slots,
                               const int32_t* indices, const float* edge, int64_t lde, const float* x, int64_t ldx,
                               const float* rowterm, int64_t ldr, int32_t unary, float slope, float* out, int64_t ldo,
                               int32_t f, int32_t epilogue, float* partials, int32_t* chain_state, int32_t phases,
                               void* stream_) {
  const char* who = "gta_aggregate_edge_sum_f32";
  cudaStream_t st = static_cast<cudaStream_t>(stream_);
  GTA_REQUIRE(f > 0 && f % 4 == 0, "%s: f=%d must be a positive multiple of 4 (pad the table)", who, f);
  GTA_REQUIRE(num_items >= 0 && num_items < (int64_t(1) << 31) - 64, "%s: bad item count", who);
  WorkList wl{reinterpret_cast<const int4*>(items_), num_items, row_slots, num_slots, indices, partials, nullptr, nullptr, 1, 0, 0};
  int rc = prepare_worklist(who, wl, chain_state, f, phases, st);
  if (rc != GTA_OK) return rc;
  if (num_items == 0 || !(phases & GTA_PHASE_MAIN)) return GTA_OK;
  GTA_REQUIRE(items_ && out && (edge || x || rowterm), "%s: null pointer (at least one of edge / x / rowterm is needed)", who);
  GTA_REQUIRE(!x || indices, "%s: indices are required to gather x", who);
  GTA_REQUIRE(unary >= GTA_UN_EXP_LEAKY_RELU && unary <= GTA_UN_COPY, "%s: bad unary %d", who, unary);
  GTA_REQUIRE(ldo % 4 == 0 && ldo >= f && (reinterpret_cast<uintptr_t>(out) & 15) == 0, "%s: out must be 16-byte aligned rows >= f", who);
  GTA_REQUIRE(!edge || (lde % 4 == 0 && lde >= f && (reinterpret_cast<uintptr_t>(edge) & 15) == 0),
              "%s: edge must be 16-byte aligned rows >= f", who);
  GTA_REQUIRE(!x || (ldx % 4 == 0 && ldx >= f && ldx * 4 < (int64_t(1) << 32) && (reinterpret_cast<uintptr_t>(x) & 15) == 0),
              "%s: x must be 16-byte aligned rows >= f, a row below 4 GiB", who);
  GTA_REQUIRE(!rowterm || (ldr % 4 == 0 && ldr >= f && (reinterpret_cast<uintptr_t>(rowterm) & 15) == 0),
              "%s: rowterm must be 16-byte aligned rows >= f", who);
  Exchange ex;
  memset(&ex, 0, sizeof(ex));
#define GTA_ES2(L, HX, HE)                                                                                          \
  do {                                                                                                              \
    auto kern = edge_sum_kernel<L, HX, HE>;                                                                         \
    dim3 grid = persistent_grid(kern, wl, L, 1, ex);                                                                \
    grid.y = (unsigned)((f + 127) / 128);                                                                           \
    kern<<<grid, kAggThreads, 0, st>>>(with_take(wl, take_for(kern, wl, L)), ex, edge, lde, x, uint32_t(ldx * 4),   \
                                       rowterm, ldr, unary, slope, out, ldo, f, epilogue);                          \
  } while (0)
#define GTA_ES(L)                                                                                                   \
  do {                                                                                                              \
    if (x && edge) GTA_ES2(L, true, true);                                                                          \
    else if (x) GTA_ES2(L, true, false);                                                                            \
    else if (edge) GTA_ES2(L, false, true);                                                                         \
    else GTA_ES2(L, false, false);                                                                                  \
  } while (0)
  switch (lanes_for(f, 4)) {
    case 4: GTA_ES(4); break;
    case 8: GTA_ES(8); break;
    case 16: GTA_ES(16); break;
    default: GTA_ES(32); break;
  }
#undef GTA_ES
#undef GTA_ES2
  GTA_CHECK_LAUNCH("edge_sum_kernel");
  return GTA_OK;
}

int gta_gather_peak_probe(const float* table, int64_t rows, int64_t ld, int32_t f, int64_t gathers_per_group,
                          float* sink, void* stream_) {
  cudaStream_t st = static_cast<cudaStream_t>(stream_);
  if (!(table && sink && rows > 0 && rows < (int64_t(1) << 32) && f >= 4 && f <= 128 && f % 4 == 0 && ld >= f &&
        ld % 4 == 0 && gathers_per_group > 0)) {
    set_error("gta_gather_peak_probe: bad arguments");
    return -GTA_ERR_INVALID;
  }
  CachePolicies pol;
  if (cache_policies(&pol) != GTA_OK) return -GTA_ERR_CUDA;
  const int lanes = lanes_for(f);
  const int ctas = resident_ctas(gather_peak_kernel);
  gather_peak_kernel<<<ctas, kAggThreads, 0, st>>>(table, uint32_t(rows), uint32_t(ld) * 4u, lanes, gathers_per_group,
                                                  reinterpret_cast<float4*>(sink), pol.keep);
  count_launch();
  if (check_cuda(cudaGetLastError(), "gather_peak_kernel") != GTA_OK) return -GTA_ERR_CUDA;
  return ctas * kAggThreads / lanes;          // > 0: the number of groups that ran (each did gathers_per_group gathers)
}

int gta_er_stats(const float* er, int64_t lder, int64_t num_sources, int64_t col_block, int32_t heads,
                 uint32_t* stats, void* stream_) {
  cudaStream_t st = static_cast<cudaStream_t>(stream_);
  GTA_REQUIRE(er && stats && num_sources >= 0 && lder >= heads, "gta_er_stats: bad arguments");
  if (heads < 1 || heads > 32 || (heads & (heads - 1)) != 0) {
    set_error("gta_er_stats: heads=%d is not a power of two <= 32 (run the aggregation without er_stats)", heads);
    return GTA_ERR_UNSUPPORTED;
  }
  const int64_t n_cb = (col_block > 0 && col_block < num_sources) ? (num_sources + col_block - 1) / col_block : 1;
  GTA_REQUIRE(n_cb <= 65535, "gta_er_stats: %lld column blocks", (long long)n_cb);
  GTA_CUDA(cudaMemsetAsync(stats, 0, size_t(n_cb) * 2 * heads * sizeof(uint32_t), st));
  count_launch();
  if (num_sources == 0) return GTA_OK;
  const int64_t rows_per_block = n_cb > 1 ? col_block : num_sources;
  int64_t ctas = (rows_per_block * heads + 256 * 8 - 1) / (256 * 8);        // about 8 rows per thread
  if (ctas < 1) ctas = 1;
  if (ctas > 4 * kNumSMs) ctas = 4 * kNumSMs;
  er_stats_kernel<<<dim3((unsigned)ctas, (unsigned)n_cb), 256, 0, st>>>(er, lder, num_sources, n_cb > 1 ? col_block : 0,
                                                                         heads, stats);
  GTA_CHECK_LAUNCH("er_stats_kernel");
  return GTA_OK;
}

int gta_gat_aggregate_f32(const int32_t* items, int64_t num_items, const int32_t* row_slots, int64_t num_slots,
                          const int32_t* indices, const float* el, const float* er, int64_t lder, int32_t heads,
                          float slope, const float* z, int64_t ldz, float* out, int64_t ldo, int32_t f,
                          int32_t epilogue, float* rowmax, float* rowsum, float* partials, int32_t* chain_state,
                          const uint32_t* er_stats, int64_t col_block, const gta_exchange_t* exchange,
                          int32_t phases, void* stream) {
  return gat_aggregate_impl<float>("gta_gat_aggregate_f32", items, num_items, row_slots, num_slots, indices, el, er, lder,
                                   heads, slope, z, ldz, out, ldo, f, epilogue, rowmax, rowsum, partials, chain_state,
                                   er_stats, col_block, exchange, phases, stream);
}

int gta_gat_aggregate_bf16(const int32_t* items, int64_t num_items, const int32_t* row_slots, int64_t num_slots,
                           const int32_t* indices, const float* el, const float* er, int64_t lder, int32_t heads,
                           float slope, const void* z, int64_t ldz, float* out, int64_t ldo, int32_t f,
                           int32_t epilogue, float* rowmax, float* rowsum, float* partials, int32_t* chain_state,
                           const uint32_t* er_stats, int64_t col_block, const gta_exchange_t* exchange,
                           int32_t phases, void* stream) {
  return gat_aggregate_impl<__nv_bfloat16>("gta_gat_aggregate_bf16", items, num_items, row_slots, num_slots, indices, el,
                                           er, lder, heads, slope, static_cast<const __nv_bfloat16*>(z), ldz, out, ldo, f,
                                           epilogue, rowmax, rowsum, partials, chain_state, er_stats, col_block, exchange,
                                           phases, stream);
}

int gta_gat_logits_f32(const int64_t* indptr, const int32_t* indices, int64_t row_begin, int64_t row_end,
                       const float* el, const float* er, int32_t heads, float slope, int32_t stabilize, float* p,
                       float* rowmax, float* rowsum, void* stream_) {
  cudaStream_t st = static_cast<cudaStream_t>(stream_);
  GTA_REQUIRE(indptr && indices && el && er && p && rowsum, "gta_gat_logits_f32: null pointer");
  GTA_REQUIRE(heads >= 1, "gta_gat_logits_f32: heads must be >= 1");
  int64_t rows = row_end - row_begin;
  if (rows <= 0) return GTA_OK;
  gat_logits_kernel<<<(unsigned)((rows * 32 + 255) / 256), 256, 0, st>>>(indptr, indices, row_begin, row_end, el, er,
                                                                         heads, slope, stabilize, p, rowmax, rowsum);
  GTA_CHECK_LAUNCH("gat_logits_kernel");
  return GTA_OK;
}

}  // extern "C"
