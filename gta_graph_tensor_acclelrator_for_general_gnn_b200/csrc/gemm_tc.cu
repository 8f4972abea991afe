// COMP_MM applynode on the 5th-generation tensor cores: Z[N,F] = X[N,K] . W[K,F] in fp32
// accuracy by 3xTF32 split accumulation, with GAT ops 1/2 (el = Z.Al, er = Z.Ar) fused into
// the epilogue so Z is never re-read (interpreter.py:145-161 COMP_MM; genGraphOP.py:49-51).
//
// Why 3xTF32: tcgen05 has no IEEE-fp32 MMA; one TF32 pass gives ~1e-3.  Every operand is split
// x = hi + lo with hi = rna_tf32(x), lo = rna_tf32(x - hi), and D += hi.hi + lo.hi + hi.lo
// (lo.lo ~ 2^-22 is dropped).  W is split (and transposed to K-major) once by a pre-pass; X is
// split per tile inside the kernel, from shared memory, so HBM sees X exactly once.
//
// Structure (persistent, one CTA per SM, 320 threads, warp-specialised):
//   warp 0      TMA producer: X tile [128 x 32] + Wt_hi / Wt_lo tiles [F x 32] per k-block,
//               SWIZZLE_128B, out-of-bounds rows/columns zero-filled by the tensor map
//   warps 6-9   splitter: hi/lo of the X tile in shared memory, fence.proxy.async, signal
//   warp 1      MMA issuer (one elected lane): 4 k-steps x 3 tcgen05.mma.kind::tf32 (M=128, N=F,
//               K=8) per k-block into a TMEM accumulator; tcgen05.commit frees the stage
//   warps 2-5   epilogue: tcgen05.ld of the accumulator (double-buffered in TMEM, so tile t+1's
//               MMAs overlap tile t's epilogue), fp32 row stores, el/er dot products
// Roofline: HBM-bound (X read once: 53 FLOP/B at Reddit shape against a ~170 FLOP/B TF32 ridge).
#include <cuda.h>
#include <cuda_bf16.h>

#include "common.cuh"

namespace gta {

constexpr int kTcBM = 128;          // rows per tile (UMMA M)
constexpr int kTcBK = 32;           // fp32 per k-block = one 128-byte swizzle row
constexpr int kTcThreads = 320;
constexpr int kTcMaxStages = 4;
constexpr uint32_t kTcATileBytes = kTcBM * kTcBK * 4;   // 16 KB

struct TcParams {
  float* z;          // fp32 rows, or bf16 rows when z_bf16 (ldz then counts bf16 elements)
  int z_bf16;
  int64_t ldz;
  int64_t num_rows;
  int k_blocks;
  int f;
  int heads;
  int64_t lder;      // row stride of er (it may live beside z in a gathered [F | H] table)
  const float* al;
  const float* ar;
  float* el;
  float* er;
  int stages;
  int tiles;
  int tmem_cols;
};

// ---- PTX wrappers --------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
// bounded wait: a protocol bug traps (context error) instead of hanging the GPU
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done = 0;
  for (uint32_t spin = 0; spin < (1u << 26); ++spin) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    if (done) return;
  }
  __trap();
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, int c0, int c1, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
      ::"r"(dst), "l"(map), "r"(c0), "r"(c1), "r"(bar) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tc_mma_tf32(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                            uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ float rna_tf32(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return __uint_as_float(r);
}

// K-major, SWIZZLE_128B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor layout):
// start>>4 [0,14) | LBO>>4 [16,30) = 1 | SBO>>4 [32,46) = 1024/16 | version [46,48) = 1 | layout [61,64) = 2
__device__ __forceinline__ uint64_t make_desc_sw128(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= uint64_t((smem_addr >> 4) & 0x3FFF);
  d |= uint64_t(1) << 16;
  d |= uint64_t(1024 >> 4) << 32;
  d |= uint64_t(1) << 46;
  d |= uint64_t(2) << 61;
  return d;
}

// instruction descriptor (cute::UMMA::InstrDescriptor): D = F32, A = B = TF32, both K-major
__host__ __device__ inline uint32_t make_idesc_tf32(int m, int n) {
  return (1u << 4) | (2u << 7) | (2u << 10) | (uint32_t(n >> 3) << 17) | (uint32_t(m >> 4) << 24);
}

template <int HMAX>
__global__ void __launch_bounds__(kTcThreads, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmap_x, const __grid_constant__ CUtensorMap tmap_whi,
               const __grid_constant__ CUtensorMap tmap_wlo, const TcParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t b_tile_bytes = uint32_t(p.f) * kTcBK * 4;
  const uint32_t stage_bytes = 2 * kTcATileBytes + 2 * b_tile_bytes;
  const uint32_t attn_bytes = uint32_t(p.heads > 0 ? 2 * p.f * p.heads * 4 : 0);
  const uint32_t attn_off = p.stages * stage_bytes;
  const uint32_t bar_off = (attn_off + attn_bytes + 15u) & ~15u;
  // barriers: full_tma[4] full_split[4] empty[4] tmem_full[2] tmem_empty[2], then the TMEM base word
  const uint32_t bar_base = smem_base + bar_off;
  auto full_tma = [&](int s) { return bar_base + 8u * s; };
  auto full_split = [&](int s) { return bar_base + 8u * (kTcMaxStages + s); };
  auto empty = [&](int s) { return bar_base + 8u * (2 * kTcMaxStages + s); };
  auto tmem_full = [&](int b) { return bar_base + 8u * (3 * kTcMaxStages + b); };
  auto tmem_empty = [&](int b) { return bar_base + 8u * (3 * kTcMaxStages + 2 + b); };
  volatile uint32_t* tmem_ptr_smem = reinterpret_cast<volatile uint32_t*>(smem_gen + bar_off + 8u * (3 * kTcMaxStages + 4));
  float* attn_smem = reinterpret_cast<float*>(smem_gen + attn_off);

  if (threadIdx.x == 0) {
    for (int s = 0; s < kTcMaxStages; ++s) {
      mbar_init(full_tma(s), 1);
      mbar_init(full_split(s), 128);
      mbar_init(empty(s), 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(tmem_full(b), 1);
      mbar_init(tmem_empty(b), 128);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_x) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_whi) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_wlo) : "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                 ::"r"(smem_u32(const_cast<uint32_t*>(tmem_ptr_smem))), "r"(uint32_t(p.tmem_cols)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (p.heads > 0) {   // Al | Ar, [F, H] row-major each, broadcast-read by the epilogue
    const int n_attn = p.f * p.heads;
    for (int i = threadIdx.x; i < n_attn; i += kTcThreads) {
      attn_smem[i] = p.al ? p.al[i] : 0.f;
      attn_smem[n_attn + i] = p.ar ? p.ar[i] : 0.f;
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  if (warp == 0) {
    // ===== TMA producer =====
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < p.tiles; tile += gridDim.x) {
        for (int kb = 0; kb < p.k_blocks; ++kb) {
          mbar_wait(empty(stage), phase ^ 1);
          const uint32_t a_hi = smem_base + stage * stage_bytes;
          const uint32_t b_hi = a_hi + 2 * kTcATileBytes;
          mbar_expect_tx(full_tma(stage), kTcATileBytes + 2 * b_tile_bytes);
          tma_load_2d(a_hi, &tmap_x, kb * kTcBK, tile * kTcBM, full_tma(stage));
          tma_load_2d(b_hi, &tmap_whi, kb * kTcBK, 0, full_tma(stage));
          tma_load_2d(b_hi + b_tile_bytes, &tmap_wlo, kb * kTcBK, 0, full_tma(stage));
          if (++stage == p.stages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer =====
    const uint32_t idesc = make_idesc_tf32(kTcBM, p.f);
    int stage = 0;
    uint32_t phase = 0;
    int it = 0;
    for (int tile = blockIdx.x; tile < p.tiles; tile += gridDim.x, ++it) {
      const int buf = it & 1;
      mbar_wait(tmem_empty(buf), ((it >> 1) & 1) ^ 1);
      tc_fence_after();
      const uint32_t acc = tmem_base + uint32_t(buf * p.f);
      for (int kb = 0; kb < p.k_blocks; ++kb) {
        mbar_wait(full_tma(stage), phase);
        mbar_wait(full_split(stage), phase);
        tc_fence_after();
        if (lane == 0) {
          const uint32_t a_hi = smem_base + stage * stage_bytes;
          const uint32_t a_lo = a_hi + kTcATileBytes;
          const uint32_t b_hi = a_hi + 2 * kTcATileBytes;
          const uint32_t b_lo = b_hi + b_tile_bytes;
#pragma unroll
          for (int ks = 0; ks < kTcBK / 8; ++ks) {
            const uint32_t koff = ks * 32;      // 8 tf32 = 32 bytes inside the 128-byte swizzle row
            const uint64_t da_hi = make_desc_sw128(a_hi + koff);
            const uint64_t da_lo = make_desc_sw128(a_lo + koff);
            const uint64_t db_hi = make_desc_sw128(b_hi + koff);
            const uint64_t db_lo = make_desc_sw128(b_lo + koff);
            tc_mma_tf32(acc, da_hi, db_hi, idesc, (kb | ks) != 0);
            tc_mma_tf32(acc, da_lo, db_hi, idesc, 1);
            tc_mma_tf32(acc, da_hi, db_lo, idesc, 1);
          }
          tc_commit(empty(stage));
          if (kb == p.k_blocks - 1) tc_commit(tmem_full(buf));
        }
        __syncwarp();
        if (++stage == p.stages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp < 6) {
    // ===== epilogue (warps 2..5): TMEM lane quarter = warp % 4 =====
    const int q = warp & 3;
    const int n_attn = p.f * p.heads;
    int it = 0;
    for (int tile = blockIdx.x; tile < p.tiles; tile += gridDim.x, ++it) {
      const int buf = it & 1;
      mbar_wait(tmem_full(buf), (it >> 1) & 1);
      tc_fence_after();
      const int64_t row = int64_t(tile) * kTcBM + q * 32 + lane;
      const bool live = row < p.num_rows;
      const uint32_t taddr = tmem_base + (uint32_t(q * 32) << 16) + uint32_t(buf * p.f);
      constexpr int HR = HMAX > 0 ? HMAX : 1;
      float sl[HR], sr[HR];
#pragma unroll
      for (int h = 0; h < HR; ++h) { sl[h] = 0.f; sr[h] = 0.f; }
      for (int c0 = 0; c0 < p.f; c0 += 16) {
        uint32_t r[16];
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
            "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
            : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
              "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
            : "r"(taddr + uint32_t(c0)));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        if (live && !p.z_bf16) {
          float* zr = p.z + row * p.ldz + c0;
#pragma unroll
          for (int v = 0; v < 4; ++v)
            *reinterpret_cast<float4*>(zr + 4 * v) = make_float4(__uint_as_float(r[4 * v]), __uint_as_float(r[4 * v + 1]),
                                                                __uint_as_float(r[4 * v + 2]), __uint_as_float(r[4 * v + 3]));
        } else if (live) {
          // bf16 storage mode: the fp32 accumulator is rounded once, here (el/er below still use the fp32 values)
          __nv_bfloat16* zr = reinterpret_cast<__nv_bfloat16*>(p.z) + row * p.ldz + c0;
          uint32_t packed[8];
#pragma unroll
          for (int v = 0; v < 8; ++v) {
            const __nv_bfloat162 b = __floats2bfloat162_rn(__uint_as_float(r[2 * v]), __uint_as_float(r[2 * v + 1]));
            packed[v] = *reinterpret_cast<const uint32_t*>(&b);
          }
          *reinterpret_cast<uint4*>(zr) = make_uint4(packed[0], packed[1], packed[2], packed[3]);
          *reinterpret_cast<uint4*>(zr + 8) = make_uint4(packed[4], packed[5], packed[6], packed[7]);
        }
        if (HMAX > 0) {
#pragma unroll
          for (int c = 0; c < 16; ++c) {
            const float zv = __uint_as_float(r[c]);
            const float* al = attn_smem + (c0 + c) * p.heads;
#pragma unroll
            for (int h = 0; h < HR; ++h) {
              if (h < p.heads) {
                sl[h] = fmaf(zv, al[h], sl[h]);
                sr[h] = fmaf(zv, al[n_attn + h], sr[h]);
              }
            }
          }
        }
      }
      tc_fence_before();
      mbar_arrive(tmem_empty(buf));
      if (live && HMAX > 0) {
#pragma unroll
        for (int h = 0; h < HR; ++h) {
          if (h < p.heads) {
            if (p.el) p.el[row * p.heads + h] = sl[h];
            if (p.er) p.er[row * p.lder + h] = sr[h];
          }
        }
      }
    }
  } else {
    // ===== splitter (warps 6..9): X tile -> hi (in place) and lo =====
    const int t = threadIdx.x - 6 * 32;       // 0..127
    int stage = 0;
    uint32_t phase = 0;
    for (int tile = blockIdx.x; tile < p.tiles; tile += gridDim.x) {
      for (int kb = 0; kb < p.k_blocks; ++kb) {
        mbar_wait(full_tma(stage), phase);
        float4* hi = reinterpret_cast<float4*>(smem_gen + stage * stage_bytes);
        float4* lo = reinterpret_cast<float4*>(smem_gen + stage * stage_bytes + kTcATileBytes);
#pragma unroll
        for (int j = 0; j < int(kTcATileBytes / 16 / 128); ++j) {
          const int i = t + 128 * j;
          float4 x = hi[i];
          float4 h, l;
          h.x = rna_tf32(x.x); h.y = rna_tf32(x.y); h.z = rna_tf32(x.z); h.w = rna_tf32(x.w);
          l.x = rna_tf32(x.x - h.x); l.y = rna_tf32(x.y - h.y); l.z = rna_tf32(x.z - h.z); l.w = rna_tf32(x.w - h.w);
          hi[i] = h;
          lo[i] = l;
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        mbar_arrive(full_split(stage));
        if (++stage == p.stages) { stage = 0; phase ^= 1; }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(uint32_t(p.tmem_cols)) : "memory");
  }
}

// W [K,F] row-major -> Wt_hi, Wt_lo [F, Kp] (K-major, zero padded to Kp = ceil32(K))
__global__ void split_w_kernel(const float* __restrict__ w, int64_t ldw, int k, int f, int kp, float* __restrict__ whi,
                               float* __restrict__ wlo) {
  int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
  if (i >= int64_t(f) * kp) return;
  int n = int(i / kp), kk = int(i % kp);
  float v = kk < k ? w[int64_t(kk) * ldw + n] : 0.f;
  float h = rna_tf32(v);
  whi[i] = h;
  wlo[i] = rna_tf32(v - h);
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  if (fn) return fn;
  void* sym = nullptr;
  cudaDriverEntryPointQueryResult qres;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &qres) != cudaSuccess || !sym) return nullptr;
  fn = reinterpret_cast<EncodeTiledFn>(sym);
  return fn;
}

static int make_map_2d(CUtensorMap* map, const void* base, uint64_t inner, uint64_t outer, uint64_t pitch_bytes,
                       uint32_t box_inner, uint32_t box_outer) {
  EncodeTiledFn enc = get_encode();
  if (!enc) {
    set_error("cuTensorMapEncodeTiled entry point not available");
    return GTA_ERR_CUDA;
  }
  cuuint64_t dims[2] = {inner, outer};
  cuuint64_t strides[1] = {pitch_bytes};
  cuuint32_t box[2] = {box_inner, box_outer};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed with CUresult %d", int(r));
    return GTA_ERR_CUDA;
  }
  return GTA_OK;
}

size_t gemm_tc_workspace(int k, int f) {
  int kp = (k + kTcBK - 1) / kTcBK * kTcBK;
  return 2 * align_up(size_t(f) * kp * 4, 1024);
}

int gemm_tc_launch(const float* x, int64_t ldx, const float* w, int64_t ldw, float* z, int64_t ldz, int64_t num_rows,
                   int k, int f, const float* al, const float* ar, int heads, float* el, float* er, int64_t lder,
                   void* workspace, size_t workspace_bytes, cudaStream_t st, int z_bf16) {
  const bool want_attn = (el && al) || (er && ar);
  if (f % 16 != 0 || f < 16 || f > 256) return GTA_ERR_UNSUPPORTED;
  if (ldx % 4 != 0 || ldz % (z_bf16 ? 8 : 4) != 0 || (reinterpret_cast<uintptr_t>(x) & 15) || (reinterpret_cast<uintptr_t>(z) & 15))
    return GTA_ERR_UNSUPPORTED;
  if (want_attn && (heads < 1 || heads > 16)) return GTA_ERR_UNSUPPORTED;
  if (num_rows >= (int64_t(1) << 31) - kTcBM) return GTA_ERR_UNSUPPORTED;
  const size_t need = gemm_tc_workspace(k, f);
  if (!workspace || workspace_bytes < need || (reinterpret_cast<uintptr_t>(workspace) & 127)) return GTA_ERR_UNSUPPORTED;

  const int kp = (k + kTcBK - 1) / kTcBK * kTcBK;
  float* whi = static_cast<float*>(workspace);
  float* wlo = reinterpret_cast<float*>(static_cast<char*>(workspace) + need / 2);
  const int64_t nw = int64_t(f) * kp;
  split_w_kernel<<<(unsigned)((nw + 255) / 256), 256, 0, st>>>(w, ldw, k, f, kp, whi, wlo);
  GTA_CHECK_LAUNCH("split_w_kernel");

  CUtensorMap mx, mhi, mlo;
  int rc = make_map_2d(&mx, x, uint64_t(k), uint64_t(num_rows), uint64_t(ldx) * 4, kTcBK, kTcBM);
  if (rc != GTA_OK) return rc;
  rc = make_map_2d(&mhi, whi, uint64_t(kp), uint64_t(f), uint64_t(kp) * 4, kTcBK, uint32_t(f));
  if (rc != GTA_OK) return rc;
  rc = make_map_2d(&mlo, wlo, uint64_t(kp), uint64_t(f), uint64_t(kp) * 4, kTcBK, uint32_t(f));
  if (rc != GTA_OK) return rc;

  TcParams p{};
  p.z = z; p.z_bf16 = z_bf16; p.ldz = ldz; p.num_rows = num_rows; p.k_blocks = kp / kTcBK; p.f = f;
  p.heads = want_attn ? heads : 0;
  p.lder = lder;
  p.al = want_attn ? al : nullptr; p.ar = want_attn ? ar : nullptr;
  p.el = want_attn ? el : nullptr; p.er = want_attn ? er : nullptr;
  p.tiles = int((num_rows + kTcBM - 1) / kTcBM);
  int cols = 32;
  while (cols < 2 * f) cols <<= 1;
  p.tmem_cols = cols;
  const size_t stage_bytes = 2 * size_t(kTcATileBytes) + 2 * size_t(f) * kTcBK * 4;
  const size_t fixed = 1024 /*alignment slack*/ + size_t(p.heads > 0 ? 2 * f * p.heads * 4 : 0) + 256 /*barriers*/;
  int dev = 0, max_smem = 0, sms = 0;
  GTA_CUDA(cudaGetDevice(&dev));
  GTA_CUDA(cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
  GTA_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  int stages = int((size_t(max_smem) - fixed) / stage_bytes);
  if (stages > kTcMaxStages) stages = kTcMaxStages;
  if (stages < 2) return GTA_ERR_UNSUPPORTED;
  p.stages = stages;
  const size_t smem = fixed + stages * stage_bytes;
  const int grid = p.tiles < sms ? p.tiles : sms;
#define GTA_TC_LAUNCH(HM)                                                                                         \
  do {                                                                                                            \
    GTA_CUDA(cudaFuncSetAttribute(gemm_tc_kernel<HM>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));   \
    gemm_tc_kernel<HM><<<grid, kTcThreads, smem, st>>>(mx, mhi, mlo, p);                                          \
  } while (0)
  if (p.heads == 0) GTA_TC_LAUNCH(0);
  else if (p.heads <= 4) GTA_TC_LAUNCH(4);
  else if (p.heads <= 8) GTA_TC_LAUNCH(8);
  else GTA_TC_LAUNCH(16);
#undef GTA_TC_LAUNCH
  GTA_CHECK_LAUNCH("gemm_tc_kernel");
  return GTA_OK;
}

}  // namespace gta
