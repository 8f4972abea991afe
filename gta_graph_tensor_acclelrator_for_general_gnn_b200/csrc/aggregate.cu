// Segmented scatter-reduce kernels: the HBM/L2-bound heart of the hot path.
//
//   aggregate_kernel      COMP_MUL_COMP_ADD / COMP_ADD gather  (interpreter.py:575-638, 85-106)
//   gat_aggregate_kernel  GAT ops 3-13 in one pass, online softmax (genGraphOP.py:52-62)
//   gat_logits_kernel     GAT block [4,5,6,7,8] alone (STORE_E p, STORE_N S)
//
// Mapping.  A work item (<= chunk edges of one destination row, see gta_schedule_build) is
// owned by a GROUP of LANES = min(F,128)/4 lanes; each lane owns 4 consecutive features, so
// one gathered source row is ONE 128-bit load per lane and a full 512 B row per 32 lanes.
// Wider rows (F > 128) are covered by blockIdx.y feature windows of 128.  Source ids (and
// scalar edge weights) are read once per group, coalesced and streaming (L2 evict_first),
// and handed round the group by shuffle / shared memory; gathered rows use the read-only
// path with L2 evict_last so the feature table stays resident in the 126 MB L2.
// UNROLL independent row loads are in flight per lane.
//
// Determinism.  Every destination row is reduced by exactly one group in ascending source
// order inside an item, and items of a long row are combined in item order by
// *_combine_kernel: a fixed-shape reduction, bitwise reproducible run to run, no atomics.
#include "common.cuh"

namespace gta {

constexpr int kAggThreads = 256;
constexpr int kUnroll = 8;

// floats per partial slot of the GAT kernel: acc[f] | max[H] | sum[H], padded to 16 bytes
__host__ __device__ inline int gat_partial_stride(int f, int heads) { return f + ((2 * heads + 3) & ~3); }

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

// ----------------------------------------------------------------------------------------
// weighted aggregate
//   WKIND 0: no weight, 1: scalar weight per edge (wh == 1), 2: per-head weight (wh > 1,
//   (f / wh) % 4 == 0 so a lane's 4 features share a head)
// ----------------------------------------------------------------------------------------
template <int LANES, int WKIND, bool DIV>
__global__ void __launch_bounds__(kAggThreads)
aggregate_kernel(const int4* __restrict__ items, int64_t num_items, const int32_t* __restrict__ indices,
                 const float* __restrict__ w, int wh, const float* __restrict__ rowden,
                 const float* __restrict__ x, int64_t ldx, float* __restrict__ out, int64_t ldo,
                 int f, int epilogue, float* __restrict__ partials) {
  const uint64_t pol_stream = policy_evict_first();
  const uint64_t pol_keep = policy_evict_last();
  const int lane = threadIdx.x & 31;
  const int l = lane & (LANES - 1);
  const int64_t group = (blockIdx.x * int64_t(kAggThreads) + threadIdx.x) / LANES;
  const int fo = blockIdx.y * 128 + 4 * l;
  const bool have = group < num_items;
  const bool active = have && fo < f;
  int4 it = have ? items[group] : make_int4(0, 0, 0, -1);
  const int count = have ? it.z : 0;
  const int max_count = (LANES == 32) ? count : warp_max_i32(count);
  const int32_t* idx_base = indices + it.y;
  const float* w_base = (WKIND != 0) ? w + int64_t(it.y) * wh : nullptr;
  int head = 0;
  float den = 1.f;
  if (WKIND == 2) head = active ? fo / (f / wh) : 0;
  if (DIV && have) den = rowden[int64_t(it.x) * wh + head];

  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int base = 0; base < max_count; base += LANES) {
    int n = count - base;
    n = n < 0 ? 0 : (n > LANES ? LANES : n);
    int my_idx = 0;
    float my_w = 0.f;
    if (l < n) {
      my_idx = ld_stream_i32(idx_base + base + l, pol_stream);
      if (WKIND == 1) {
        my_w = ld_stream_f32(w_base + base + l, pol_stream);
        if (DIV) my_w = my_w / den;
      }
    }
    for (int j = 0; j < LANES; j += kUnroll) {
      if (LANES == 32 && j >= n) break;   // warp-uniform when a group is a whole warp
      float4 v[kUnroll];
      float wv[kUnroll];
#pragma unroll
      for (int u = 0; u < kUnroll; ++u) {
        if (j + u < LANES) {
          int src = __shfl_sync(0xffffffffu, my_idx, j + u, LANES);
          float ws = 1.f;
          if (WKIND == 1) ws = __shfl_sync(0xffffffffu, my_w, j + u, LANES);
          const bool ok = active && (j + u) < n;
          v[u] = ok ? ld_gather_f32x4(x + int64_t(src) * ldx + fo, pol_keep) : make_float4(0.f, 0.f, 0.f, 0.f);
          if (WKIND == 2) {
            ws = ok ? __ldg(w_base + int64_t(base + j + u) * wh + head) : 0.f;
            if (DIV) ws = ws / den;
          }
          wv[u] = ws;
        }
      }
#pragma unroll
      for (int u = 0; u < kUnroll; ++u) {
        if (j + u < LANES) {
          acc.x = fmaf(wv[u], v[u].x, acc.x);
          acc.y = fmaf(wv[u], v[u].y, acc.y);
          acc.z = fmaf(wv[u], v[u].z, acc.z);
          acc.w = fmaf(wv[u], v[u].w, acc.w);
        }
      }
    }
  }
  if (!active) return;
  if (it.w < 0) {
    acc.x = apply_epilogue(acc.x, epilogue);
    acc.y = apply_epilogue(acc.y, epilogue);
    acc.z = apply_epilogue(acc.z, epilogue);
    acc.w = apply_epilogue(acc.w, epilogue);
    st_stream_f32x4(out + int64_t(it.x) * ldo + fo, acc);
  } else {
    *reinterpret_cast<float4*>(partials + int64_t(it.w) * f + fo) = acc;
  }
}

// rows cut into several items: sum the partial rows in item order
__global__ void __launch_bounds__(kAggThreads)
aggregate_combine_kernel(const int4* __restrict__ items, int64_t num_items, const float* __restrict__ partials,
                         float* __restrict__ out, int64_t ldo, int f, int epilogue) {
  const int lane = threadIdx.x & 31;
  const int64_t idx = (blockIdx.x * int64_t(kAggThreads) + threadIdx.x) >> 5;
  if (idx >= num_items) return;
  int4 it = items[idx];
  if (it.w < 0) return;
  if (idx > 0 && items[idx - 1].x == it.x) return;   // not the first item of its row
  for (int fo = 4 * lane; fo < f; fo += 128) {
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int64_t c = idx; c < num_items; ++c) {
      int4 ic = items[c];
      if (ic.x != it.x) break;
      float4 p = *reinterpret_cast<const float4*>(partials + int64_t(ic.w) * f + fo);
      acc.x += p.x; acc.y += p.y; acc.z += p.z; acc.w += p.w;
    }
    acc.x = apply_epilogue(acc.x, epilogue);
    acc.y = apply_epilogue(acc.y, epilogue);
    acc.z = apply_epilogue(acc.z, epilogue);
    acc.w = apply_epilogue(acc.w, epilogue);
    *reinterpret_cast<float4*>(out + int64_t(it.x) * ldo + fo) = acc;
  }
}

// ----------------------------------------------------------------------------------------
// GAT edge phase, single pass (online softmax over batches of LANES edges)
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

template <int LANES, int H>
__global__ void __launch_bounds__(kAggThreads)
gat_aggregate_kernel(const int4* __restrict__ items, int64_t num_items, const int32_t* __restrict__ indices,
                     const float* __restrict__ el, const float* __restrict__ er, float slope,
                     const float* __restrict__ z, int64_t ldz, float* __restrict__ out, int64_t ldo,
                     int f, int epilogue, float* __restrict__ rowmax, float* __restrict__ rowsum,
                     float* __restrict__ partials) {
  // per warp: 32 staged edges = source id + H softmax numerators
  __shared__ int s_idx[kAggThreads / 32][32];
  __shared__ float s_p[kAggThreads / 32][32][H];
  const uint64_t pol_stream = policy_evict_first();
  const uint64_t pol_keep = policy_evict_last();
  const int warp_in_cta = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int l = lane & (LANES - 1);
  const int gbase = lane & ~(LANES - 1);          // first lane of my group inside the warp
  const int64_t group = (blockIdx.x * int64_t(kAggThreads) + threadIdx.x) / LANES;
  const int fo = blockIdx.y * 128 + 4 * l;
  const bool have = group < num_items;
  const bool active = have && fo < f;
  int4 it = have ? items[group] : make_int4(0, 0, 0, -1);
  const int count = have ? it.z : 0;
  const int max_count = (LANES == 32) ? count : warp_max_i32(count);
  const int32_t* idx_base = indices + it.y;
  const int head = active ? fo / (f / H) : 0;

  float elr[H], m[H], s[H];
  if (have) load_heads<H>(el + int64_t(it.x) * H, elr);
#pragma unroll
  for (int h = 0; h < H; ++h) { m[h] = -INFINITY; s[h] = 0.f; if (!have) elr[h] = 0.f; }
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);

  for (int base = 0; base < max_count; base += LANES) {
    int n = count - base;
    n = n < 0 ? 0 : (n > LANES ? LANES : n);
    float e[H];
    int my_idx = 0;
    if (l < n) {
      my_idx = ld_stream_i32(idx_base + base + l, pol_stream);
      float erv[H];
      load_heads<H>(er + int64_t(my_idx) * H, erv);
#pragma unroll
      for (int h = 0; h < H; ++h) e[h] = leaky(elr[h] + erv[h], slope);
    } else {
#pragma unroll
      for (int h = 0; h < H; ++h) e[h] = -INFINITY;
    }
    float my_scale = 1.f;
#pragma unroll
    for (int h = 0; h < H; ++h) {
      float bm = group_max<LANES>(e[h]);
      float mn = fmaxf(m[h], bm);
      // n == 0 for this group (another group in the warp is still running): mn may be -inf
      float sc = (mn == -INFINITY) ? 1.f : expf(m[h] - mn);
      float p = (l < n) ? expf(e[h] - mn) : 0.f;
      float bs = group_sum<LANES>(p);
      s[h] = s[h] * sc + bs;
      m[h] = mn;
      my_scale = (h == head) ? sc : my_scale;
      s_p[warp_in_cta][lane][h] = p;
    }
    s_idx[warp_in_cta][lane] = my_idx;
    acc.x *= my_scale; acc.y *= my_scale; acc.z *= my_scale; acc.w *= my_scale;
    __syncwarp();
    for (int j = 0; j < LANES; j += kUnroll) {
      if (LANES == 32 && j >= n) break;
      float4 v[kUnroll];
      float pv[kUnroll];
#pragma unroll
      for (int u = 0; u < kUnroll; ++u) {
        if (j + u < LANES) {
          const bool ok = active && (j + u) < n;
          int src = s_idx[warp_in_cta][gbase + j + u];
          pv[u] = s_p[warp_in_cta][gbase + j + u][head];
          v[u] = ok ? ld_gather_f32x4(z + int64_t(src) * ldz + fo, pol_keep) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
      }
#pragma unroll
      for (int u = 0; u < kUnroll; ++u) {
        if (j + u < LANES) {
          acc.x = fmaf(pv[u], v[u].x, acc.x);
          acc.y = fmaf(pv[u], v[u].y, acc.y);
          acc.z = fmaf(pv[u], v[u].z, acc.z);
          acc.w = fmaf(pv[u], v[u].w, acc.w);
        }
      }
    }
    __syncwarp();
  }
  if (!have) return;
  if (it.w < 0) {
    if (active) {
      float sh = pick<H>(s, head);
      float inv = sh > 0.f ? 1.f / sh : 0.f;
      acc.x = apply_epilogue(acc.x * inv, epilogue);
      acc.y = apply_epilogue(acc.y * inv, epilogue);
      acc.z = apply_epilogue(acc.z * inv, epilogue);
      acc.w = apply_epilogue(acc.w * inv, epilogue);
      st_stream_f32x4(out + int64_t(it.x) * ldo + fo, acc);
    }
    if (blockIdx.y == 0 && l < H) {
      if (rowmax) rowmax[int64_t(it.x) * H + l] = (count > 0) ? pick<H>(m, l) : 0.f;
      if (rowsum) rowsum[int64_t(it.x) * H + l] = pick<H>(s, l);
    }
  } else {
    float* part = partials + int64_t(it.w) * gat_partial_stride(f, H);
    if (active) *reinterpret_cast<float4*>(part + fo) = acc;
    if (blockIdx.y == 0 && l < H) {
      part[f + l] = pick<H>(m, l);
      part[f + H + l] = pick<H>(s, l);
    }
  }
}

// merge the (max, sum, acc) triples of a long row in item order
template <int H>
__global__ void __launch_bounds__(kAggThreads)
gat_combine_kernel(const int4* __restrict__ items, int64_t num_items, const float* __restrict__ partials,
                   float* __restrict__ out, int64_t ldo, int f, int epilogue,
                   float* __restrict__ rowmax, float* __restrict__ rowsum) {
  const int lane = threadIdx.x & 31;
  const int64_t idx = (blockIdx.x * int64_t(kAggThreads) + threadIdx.x) >> 5;
  if (idx >= num_items) return;
  int4 it = items[idx];
  if (it.w < 0) return;
  if (idx > 0 && items[idx - 1].x == it.x) return;
  const int stride = gat_partial_stride(f, H);
  const int d = f / H;
  // pass 1: global max and rescaled sum per head (every lane redundantly, H is small)
  float gm[H], gs[H];
#pragma unroll
  for (int h = 0; h < H; ++h) { gm[h] = -INFINITY; gs[h] = 0.f; }
  for (int64_t c = idx; c < num_items; ++c) {
    int4 ic = items[c];
    if (ic.x != it.x) break;
    const float* part = partials + int64_t(ic.w) * stride;
#pragma unroll
    for (int h = 0; h < H; ++h) gm[h] = fmaxf(gm[h], part[f + h]);
  }
  for (int64_t c = idx; c < num_items; ++c) {
    int4 ic = items[c];
    if (ic.x != it.x) break;
    const float* part = partials + int64_t(ic.w) * stride;
#pragma unroll
    for (int h = 0; h < H; ++h) gs[h] += part[f + H + h] * expf(part[f + h] - gm[h]);
  }
  for (int fo = 4 * lane; fo < f; fo += 128) {
    const int head = fo / d;
    const float mh = pick<H>(gm, head);
    const float sh = pick<H>(gs, head);
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int64_t c = idx; c < num_items; ++c) {
      int4 ic = items[c];
      if (ic.x != it.x) break;
      const float* part = partials + int64_t(ic.w) * stride;
      float sc = expf(part[f + head] - mh);
      float4 p = *reinterpret_cast<const float4*>(part + fo);
      acc.x = fmaf(sc, p.x, acc.x); acc.y = fmaf(sc, p.y, acc.y);
      acc.z = fmaf(sc, p.z, acc.z); acc.w = fmaf(sc, p.w, acc.w);
    }
    float inv = sh > 0.f ? 1.f / sh : 0.f;
    acc.x = apply_epilogue(acc.x * inv, epilogue);
    acc.y = apply_epilogue(acc.y * inv, epilogue);
    acc.z = apply_epilogue(acc.z * inv, epilogue);
    acc.w = apply_epilogue(acc.w * inv, epilogue);
    *reinterpret_cast<float4*>(out + int64_t(it.x) * ldo + fo) = acc;
  }
  if (lane < H) {
    if (rowmax) rowmax[int64_t(it.x) * H + lane] = pick<H>(gm, lane);
    if (rowsum) rowsum[int64_t(it.x) * H + lane] = pick<H>(gs, lane);
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

static int lanes_for(int f) {
  int v = (f < 128 ? f : 128) / 4;
  int l = 1;
  while (l < v) l <<= 1;
  return l < 4 ? 4 : l;
}

template <int LANES, int WKIND, bool DIV>
static void launch_aggregate(dim3 grid, cudaStream_t st, const int4* items, int64_t num_items, const int32_t* indices,
                             const float* w, int wh, const float* rowden, const float* x, int64_t ldx, float* out,
                             int64_t ldo, int f, int epi, float* partials) {
  aggregate_kernel<LANES, WKIND, DIV><<<grid, kAggThreads, 0, st>>>(items, num_items, indices, w, wh, rowden, x, ldx,
                                                                    out, ldo, f, epi, partials);
}

template <int LANES>
static int dispatch_aggregate(int wkind, bool div, dim3 grid, cudaStream_t st, const int4* items, int64_t num_items,
                              const int32_t* indices, const float* w, int wh, const float* rowden, const float* x,
                              int64_t ldx, float* out, int64_t ldo, int f, int epi, float* partials) {
#define GTA_AGG(K, D) launch_aggregate<LANES, K, D>(grid, st, items, num_items, indices, w, wh, rowden, x, ldx, out, ldo, f, epi, partials)
  if (wkind == 0) GTA_AGG(0, false);
  else if (wkind == 1 && !div) GTA_AGG(1, false);
  else if (wkind == 1 && div) GTA_AGG(1, true);
  else if (wkind == 2 && !div) GTA_AGG(2, false);
  else GTA_AGG(2, true);
#undef GTA_AGG
  return GTA_OK;
}

template <int LANES, int H>
static void launch_gat(dim3 grid, cudaStream_t st, const int4* items, int64_t num_items, const int32_t* indices,
                       const float* el, const float* er, float slope, const float* z, int64_t ldz, float* out,
                       int64_t ldo, int f, int epi, float* rowmax, float* rowsum, float* partials) {
  gat_aggregate_kernel<LANES, H><<<grid, kAggThreads, 0, st>>>(items, num_items, indices, el, er, slope, z, ldz, out,
                                                               ldo, f, epi, rowmax, rowsum, partials);
}

template <int H>
static int dispatch_gat(int lanes, dim3 grid, cudaStream_t st, const int4* items, int64_t num_items,
                        const int32_t* indices, const float* el, const float* er, float slope, const float* z,
                        int64_t ldz, float* out, int64_t ldo, int f, int epi, float* rowmax, float* rowsum,
                        float* partials) {
#define GTA_GAT(L) launch_gat<L, H>(grid, st, items, num_items, indices, el, er, slope, z, ldz, out, ldo, f, epi, rowmax, rowsum, partials)
  switch (lanes) {
    case 4: if (H <= 4) { GTA_GAT(4); return GTA_OK; } break;
    case 8: if (H <= 8) { GTA_GAT(8); return GTA_OK; } break;
    case 16: GTA_GAT(16); return GTA_OK;
    case 32: GTA_GAT(32); return GTA_OK;
  }
#undef GTA_GAT
  return GTA_ERR_UNSUPPORTED;
}

}  // namespace gta

using namespace gta;

extern "C" {

int32_t gta_gat_partial_stride(int32_t f, int32_t heads) { return gat_partial_stride(f, heads); }

int gta_aggregate_f32(const int32_t* items_, int64_t num_items, int64_t num_slots, const int64_t* indptr,
                      const int32_t* indices, int32_t wmode, const float* w, int32_t wh, const float* rowden,
                      const float* x, int64_t ldx, float* out, int64_t ldo, int32_t f, int32_t epilogue,
                      float* partials, void* stream_) {
  (void)indptr;
  cudaStream_t st = static_cast<cudaStream_t>(stream_);
  if (num_items == 0) return GTA_OK;
  GTA_REQUIRE(items_ && indices && x && out, "gta_aggregate_f32: null pointer");
  GTA_REQUIRE(f > 0 && f % 4 == 0, "gta_aggregate_f32: f=%d must be a positive multiple of 4 (pad the table)", f);
  GTA_REQUIRE(ldx % 4 == 0 && ldo % 4 == 0 && ldx >= f && ldo >= f, "gta_aggregate_f32: leading dimensions must be multiples of 4 and >= f");
  GTA_REQUIRE((reinterpret_cast<uintptr_t>(x) & 15) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0, "gta_aggregate_f32: tables must be 16-byte aligned");
  GTA_REQUIRE(wmode >= GTA_W_NONE && wmode <= GTA_W_EDGE_DIV, "gta_aggregate_f32: bad wmode %d", wmode);
  GTA_REQUIRE(num_slots == 0 || partials, "gta_aggregate_f32: partials required for %lld slots", (long long)num_slots);
  int wkind = 0;
  bool div = wmode == GTA_W_EDGE_DIV;
  if (wmode != GTA_W_NONE) {
    GTA_REQUIRE(w && wh >= 1 && f % wh == 0, "gta_aggregate_f32: weight width %d must divide f=%d", wh, f);
    GTA_REQUIRE(!div || rowden, "gta_aggregate_f32: rowden required for GTA_W_EDGE_DIV");
    wkind = wh == 1 ? 1 : 2;
    if (wkind == 2 && (f / wh) % 4 != 0) {
      set_error("gta_aggregate_f32: per-head width f/wh=%d is not a multiple of 4", f / wh);
      return GTA_ERR_UNSUPPORTED;
    }
  }
  const int4* items = reinterpret_cast<const int4*>(items_);
  int lanes = lanes_for(f);
  int64_t threads = num_items * lanes;
  dim3 grid((unsigned)((threads + kAggThreads - 1) / kAggThreads), (unsigned)((f + 127) / 128));
  switch (lanes) {
    case 4: dispatch_aggregate<4>(wkind, div, grid, st, items, num_items, indices, w, wh, rowden, x, ldx, out, ldo, f, epilogue, partials); break;
    case 8: dispatch_aggregate<8>(wkind, div, grid, st, items, num_items, indices, w, wh, rowden, x, ldx, out, ldo, f, epilogue, partials); break;
    case 16: dispatch_aggregate<16>(wkind, div, grid, st, items, num_items, indices, w, wh, rowden, x, ldx, out, ldo, f, epilogue, partials); break;
    default: dispatch_aggregate<32>(wkind, div, grid, st, items, num_items, indices, w, wh, rowden, x, ldx, out, ldo, f, epilogue, partials); break;
  }
  GTA_CHECK_LAUNCH("aggregate_kernel");
  if (num_slots > 0) {
    int64_t cthreads = num_items * 32;
    aggregate_combine_kernel<<<(unsigned)((cthreads + kAggThreads - 1) / kAggThreads), kAggThreads, 0, st>>>(
        items, num_items, partials, out, ldo, f, epilogue);
    GTA_CHECK_LAUNCH("aggregate_combine_kernel");
  }
  return GTA_OK;
}

int gta_gat_aggregate_f32(const int32_t* items_, int64_t num_items, int64_t num_slots, const int64_t* indptr,
                          const int32_t* indices, const float* el, const float* er, int32_t heads, float slope,
                          const float* z, int64_t ldz, float* out, int64_t ldo, int32_t f, int32_t epilogue,
                          float* rowmax, float* rowsum, float* partials, void* stream_) {
  (void)indptr;
  cudaStream_t st = static_cast<cudaStream_t>(stream_);
  if (num_items == 0) return GTA_OK;
  GTA_REQUIRE(items_ && indices && el && er && z && out, "gta_gat_aggregate_f32: null pointer");
  GTA_REQUIRE(f > 0 && f % 4 == 0, "gta_gat_aggregate_f32: f=%d must be a positive multiple of 4", f);
  GTA_REQUIRE(ldz % 4 == 0 && ldo % 4 == 0 && ldz >= f && ldo >= f, "gta_gat_aggregate_f32: leading dimensions must be multiples of 4 and >= f");
  GTA_REQUIRE((reinterpret_cast<uintptr_t>(z) & 15) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0 &&
              (reinterpret_cast<uintptr_t>(er) & 15) == 0 && (reinterpret_cast<uintptr_t>(el) & 15) == 0,
              "gta_gat_aggregate_f32: tables must be 16-byte aligned");
  GTA_REQUIRE(heads >= 1 && f % heads == 0, "gta_gat_aggregate_f32: heads=%d must divide f=%d", heads, f);
  GTA_REQUIRE(num_slots == 0 || partials, "gta_gat_aggregate_f32: partials required for %lld slots", (long long)num_slots);
  if ((f / heads) % 4 != 0) {
    set_error("gta_gat_aggregate_f32: per-head width f/heads=%d is not a multiple of 4", f / heads);
    return GTA_ERR_UNSUPPORTED;
  }
  const int4* items = reinterpret_cast<const int4*>(items_);
  int lanes = lanes_for(f);
  int64_t threads = num_items * lanes;
  dim3 grid((unsigned)((threads + kAggThreads - 1) / kAggThreads), (unsigned)((f + 127) / 128));
  int rc = GTA_ERR_UNSUPPORTED;
#define GTA_GAT_H(HH) rc = dispatch_gat<HH>(lanes, grid, st, items, num_items, indices, el, er, slope, z, ldz, out, ldo, f, epilogue, rowmax, rowsum, partials)
  switch (heads) {
    case 1: GTA_GAT_H(1); break;
    case 2: GTA_GAT_H(2); break;
    case 4: GTA_GAT_H(4); break;
    case 8: GTA_GAT_H(8); break;
    case 16: GTA_GAT_H(16); break;
    default: break;
  }
#undef GTA_GAT_H
  if (rc != GTA_OK) {
    set_error("gta_gat_aggregate_f32: no kernel for heads=%d, f=%d", heads, f);
    return rc;
  }
  GTA_CHECK_LAUNCH("gat_aggregate_kernel");
  if (num_slots > 0) {
    int64_t cthreads = num_items * 32;
    unsigned cgrid = (unsigned)((cthreads + kAggThreads - 1) / kAggThreads);
    switch (heads) {
      case 1: gat_combine_kernel<1><<<cgrid, kAggThreads, 0, st>>>(items, num_items, partials, out, ldo, f, epilogue, rowmax, rowsum); break;
      case 2: gat_combine_kernel<2><<<cgrid, kAggThreads, 0, st>>>(items, num_items, partials, out, ldo, f, epilogue, rowmax, rowsum); break;
      case 4: gat_combine_kernel<4><<<cgrid, kAggThreads, 0, st>>>(items, num_items, partials, out, ldo, f, epilogue, rowmax, rowsum); break;
      case 8: gat_combine_kernel<8><<<cgrid, kAggThreads, 0, st>>>(items, num_items, partials, out, ldo, f, epilogue, rowmax, rowsum); break;
      default: gat_combine_kernel<16><<<cgrid, kAggThreads, 0, st>>>(items, num_items, partials, out, ldo, f, epilogue, rowmax, rowsum); break;
    }
    GTA_CHECK_LAUNCH("gat_combine_kernel");
  }
  return GTA_OK;
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
