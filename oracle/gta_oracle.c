/*
 * gta_oracle.c -- plain-C restatement of the hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
 * may link or call this; the product (libgta_b200.so) never does.
 *
 * PARITY UNPINNED for floating point: the reference never executes a tensor op
 * (vTCAD/code/interpreter.py:1-3); the semantics below are the ones pinned in
 * oracle/gta_oracle.py (which cites the reference's only statements: template/
 * ISA_defination.yaml:28-61, template/GAT_op.png, vTCAD/GraphOP/genGraphOP.py:34-77,
 * tile walk vTCAD/code/simulator.py:262-263,292) and this file is checked against that
 * module in tests/test_cpu_oracle_c.py.  It exists because the numpy oracle materialises
 * E x F and cannot time Reddit-shape samples; this one streams, and is the "port" CPU
 * baseline (OpenMP over destination rows, all host cores).
 *
 * Per-destination reductions run in ascending source order (CSR order), serially, in the
 * accumulator type T_ACC (double for the checker, float for the timed fp32 baseline).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* torchrun exports OMP_NUM_THREADS=1; the CPU baseline is meant to use every host core */
void gta_oracle_set_threads(int n) {
#ifdef _OPENMP
  if (n > 0) omp_set_num_threads(n);
#else
  (void)n;
#endif
}

int gta_oracle_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

/* COMP_MM applynode: Z[n,f] = X[n,k] . W[k,f]   (ISA_defination.yaml:28-31) */
#define DEFINE_GEMM(NAME, T_IN, T_ACC)                                                        \
  void NAME(const T_IN* x, int64_t ldx, const T_IN* w, T_IN* z, int64_t ldz, int64_t n, int k, \
            int f) {                                                                          \
    _Pragma("omp parallel for schedule(static)") for (int64_t i = 0; i < n; ++i) {            \
      T_ACC acc[1024];                                                                        \
      for (int c = 0; c < f; ++c) acc[c] = 0;                                                 \
      for (int kk = 0; kk < k; ++kk) {                                                        \
        T_ACC a = (T_ACC)x[i * ldx + kk];                                                     \
        const T_IN* wr = w + (int64_t)kk * f;                                                 \
        for (int c = 0; c < f; ++c) acc[c] += a * (T_ACC)wr[c];                               \
      }                                                                                       \
      for (int c = 0; c < f; ++c) z[i * ldz + c] = (T_IN)acc[c];                              \
    }                                                                                         \
  }
DEFINE_GEMM(gta_oracle_gemm_f32, float, float)
DEFINE_GEMM(gta_oracle_gemm_f64, double, double)

/* COMP_MUL_COMP_ADD: out[i] = sum_k w[k] * x[src k]   (w may be NULL; scalar per edge) */
#define DEFINE_SPMM(NAME, T_IN, T_ACC)                                                          \
  void NAME(const int64_t* indptr, const int32_t* indices, const T_IN* w, const T_IN* x,        \
            int64_t ldx, T_IN* out, int64_t ldo, int64_t row_begin, int64_t row_end, int f) {   \
    _Pragma("omp parallel for schedule(dynamic, 64)") for (int64_t i = row_begin; i < row_end; ++i) { \
      T_ACC acc[1024];                                                                          \
      for (int c = 0; c < f; ++c) acc[c] = 0;                                                   \
      for (int64_t e = indptr[i]; e < indptr[i + 1]; ++e) {                                     \
        const T_IN* xr = x + (int64_t)indices[e] * ldx;                                         \
        T_ACC wv = w ? (T_ACC)w[e] : (T_ACC)1;                                                  \
        for (int c = 0; c < f; ++c) acc[c] += wv * (T_ACC)xr[c];                                \
      }                                                                                         \
      for (int c = 0; c < f; ++c) out[(i - row_begin) * ldo + c] = (T_IN)acc[c];                \
    }                                                                                           \
  }
DEFINE_SPMM(gta_oracle_spmm_f32, float, float)
DEFINE_SPMM(gta_oracle_spmm_f64, double, double)

/* GAT edge phase, ops 3-13 (genGraphOP.py:52-62; GAT_op.png):
 *   e = leaky_relu(el[i,h] + er[j,h]); p = exp(e - rowmax); alpha = p / sum p;
 *   out[i] = ELU(sum_k alpha[k,h(c)] * z[j,c]).  Two passes per row (max, then sums),
 *   both in ascending source order.  el is indexed by (i - row_begin). */
#define DEFINE_GAT(NAME, T_IN, T_ACC, EXP, EXPM1)                                               \
  void NAME(const int64_t* indptr, const int32_t* indices, const T_IN* el, const T_IN* er,      \
            int heads, T_IN slope, const T_IN* z, int64_t ldz, T_IN* out, int64_t ldo,          \
            int64_t row_begin, int64_t row_end, int f, int activation) {                        \
    const int d = f / heads;                                                                    \
    _Pragma("omp parallel for schedule(dynamic, 64)") for (int64_t i = row_begin; i < row_end; ++i) { \
      T_ACC acc[1024], mx[64], sm[64];                                                          \
      const T_IN* eli = el + (i - row_begin) * heads;                                           \
      for (int h = 0; h < heads; ++h) { mx[h] = -INFINITY; sm[h] = 0; }                         \
      for (int c = 0; c < f; ++c) acc[c] = 0;                                                   \
      for (int64_t e = indptr[i]; e < indptr[i + 1]; ++e) {                                     \
        const T_IN* erj = er + (int64_t)indices[e] * heads;                                     \
        for (int h = 0; h < heads; ++h) {                                                       \
          T_ACC s = (T_ACC)eli[h] + (T_ACC)erj[h];                                              \
          s = s > 0 ? s : s * (T_ACC)slope;                                                     \
          if (s > mx[h]) mx[h] = s;                                                             \
        }                                                                                       \
      }                                                                                         \
      for (int64_t e = indptr[i]; e < indptr[i + 1]; ++e) {                                     \
        const int32_t j = indices[e];                                                           \
        const T_IN* erj = er + (int64_t)j * heads;                                              \
        const T_IN* zj = z + (int64_t)j * ldz;                                                  \
        for (int h = 0; h < heads; ++h) {                                                       \
          T_ACC s = (T_ACC)eli[h] + (T_ACC)erj[h];                                              \
          s = s > 0 ? s : s * (T_ACC)slope;                                                     \
          T_ACC p = EXP(s - mx[h]);                                                             \
          sm[h] += p;                                                                           \
          for (int c = h * d; c < (h + 1) * d; ++c) acc[c] += p * (T_ACC)zj[c];                 \
        }                                                                                       \
      }                                                                                         \
      for (int c = 0; c < f; ++c) {                                                             \
        T_ACC s = sm[c / d];                                                                    \
        T_ACC o = s > 0 ? acc[c] / s : 0;                                                       \
        if (activation) o = o > 0 ? o : EXPM1(o);                                               \
        out[(i - row_begin) * ldo + c] = (T_IN)o;                                               \
      }                                                                                         \
    }                                                                                           \
  }
DEFINE_GAT(gta_oracle_gat_f32, float, float, expf, expm1f)
DEFINE_GAT(gta_oracle_gat_f64, double, double, exp, expm1)

/* The same edge phase in double precision, returning beside out[i,c] the error scale of that element:
 *   scale[i,c] = sum_k alpha[k,h(c)] * zabs[j,c],   zabs = |X|.|W|  (what bounds the rounding error of Z[j,c]).
 * This is the `rowscale` of the stated fp32 tolerance |y - y64| <= 1e-5 |y64| + 1e-5 rowscale
 *  for a whole GAT layer; the at-scale parity checks (bench.py, tests/) use it. */
void gta_oracle_gat_scaled_f64(const int64_t* indptr, const int32_t* indices, const double* el, const double* er,
                               int heads, double slope, const double* z, const double* zabs, int64_t ldz,
                               double* out, double* scale, int64_t ldo, int64_t row_begin, int64_t row_end,
                               int f, int activation) {
  const int d = f / heads;
#pragma omp parallel for schedule(dynamic, 64)
  for (int64_t i = row_begin; i < row_end; ++i) {
    double acc[1024], sc[1024], mx[64], sm[64];
    const double* eli = el + (i - row_begin) * heads;
    for (int h = 0; h < heads; ++h) { mx[h] = -INFINITY; sm[h] = 0; }
    for (int c = 0; c < f; ++c) { acc[c] = 0; sc[c] = 0; }
    for (int64_t e = indptr[i]; e < indptr[i + 1]; ++e) {
      const double* erj = er + (int64_t)indices[e] * heads;
      for (int h = 0; h < heads; ++h) {
        double s = eli[h] + erj[h];
        s = s > 0 ? s : s * slope;
        if (s > mx[h]) mx[h] = s;
      }
    }
    for (int64_t e = indptr[i]; e < indptr[i + 1]; ++e) {
      const int32_t j = indices[e];
      const double* erj = er + (int64_t)j * heads;
      const double* zj = z + (int64_t)j * ldz;
      const double* aj = zabs + (int64_t)j * ldz;
      for (int h = 0; h < heads; ++h) {
        double s = eli[h] + erj[h];
        s = s > 0 ? s : s * slope;
        const double p = exp(s - mx[h]);
        sm[h] += p;
        for (int c = h * d; c < (h + 1) * d; ++c) { acc[c] += p * zj[c]; sc[c] += p * aj[c]; }
      }
    }
    for (int c = 0; c < f; ++c) {
      const double s = sm[c / d];
      double o = s > 0 ? acc[c] / s : 0;
      if (activation) o = o > 0 ? o : expm1(o);
      out[(i - row_begin) * ldo + c] = o;
      scale[(i - row_begin) * ldo + c] = s > 0 ? sc[c] / s : 0;
    }
  }
}

/* GAT ops 1,2: el = Z.Al, er = Z.Ar with [f,heads] weights */
#define DEFINE_PROJ(NAME, T_IN, T_ACC)                                                        \
  void NAME(const T_IN* z, int64_t ldz, const T_IN* a, T_IN* out, int64_t n, int f, int heads) { \
    _Pragma("omp parallel for schedule(static)") for (int64_t i = 0; i < n; ++i) {            \
      for (int h = 0; h < heads; ++h) {                                                       \
        T_ACC s = 0;                                                                          \
        for (int c = 0; c < f; ++c) s += (T_ACC)z[i * ldz + c] * (T_ACC)a[(int64_t)c * heads + h]; \
        out[i * heads + h] = (T_IN)s;                                                         \
      }                                                                                       \
    }                                                                                         \
  }
DEFINE_PROJ(gta_oracle_proj_f32, float, float)
DEFINE_PROJ(gta_oracle_proj_f64, double, double)

/* calculate_sparsity(row = tile_rows, col = 1) (code/preprocessing.py:12-40), streamed */
void gta_oracle_tile_nnz(const int64_t* indptr, const int32_t* indices, int64_t n, int64_t tile_rows,
                         int64_t* counts) {
  int64_t tiles = (n + tile_rows - 1) / tile_rows;
  memset(counts, 0, (size_t)(tiles * n) * sizeof(int64_t));
  for (int64_t i = 0; i < n; ++i)
    for (int64_t e = indptr[i]; e < indptr[i + 1]; ++e)
      if (indices[e] != i) counts[(i / tile_rows) * n + indices[e]] += 1;
}
