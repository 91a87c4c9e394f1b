// scan_rows.cuh -- the exact row scorer shared by the streaming scan (scan.cu) and the batched-query
// candidate re-score (batch.cu): reference arithmetic of Distances.l2 / Distances.cosine
// (J/util/Distances.java:31-153) with TPR = L/4 threads per row owning the modelled SIMD lanes,
// an fp32 pre-filter against the collector's current k-th key, and the fp64 score for survivors.
#pragma once

#include "topk.cuh"

namespace vs {

constexpr int SCAN_IV = 8;  // 128-bit loads per row issued back to back

// float threshold for the fp32 pre-filter: a row whose fp32 figure is > fthr cannot beat thr.
template <bool COSINE>
static __device__ __noinline__ float scan_filter_threshold(const Key& thr) {
  if (thr.hi == KEY_EMPTY64) return __int_as_float(0x7f800000);   // no bound yet: everything passes
  if (thr.hi == 0ull) return __int_as_float(0xff800000);          // k-th is NaN: only NaN rows pass
  const double val = f64_from_ordered(thr.hi);                    // -score of the k-th
  if (COSINE) {
    // fp32 estimate of -sim is within 2^-20 relative of the exact fp64 value
    double m = val + fabs(val) * (1.0 / 1048576.0) + 1e-37;
    return f32_next_up(__double2float_ru(m));
  }
  // val = l2 distance of the k-th; rows are filtered on the fp32 lane sum s, sum >= s exactly
  double m = val * val * (1.0 + 1.0 / 1125899906842624.0);
  float f = __double2float_ru(m);
  return (f == __int_as_float(0x7f800000)) ? f : f32_next_up(f);
}

// exact fp64 score of one surviving row -> ranking key (cold path: a few rows per warp)
template <bool COSINE>
static __device__ __noinline__ Key scan_exact_key(const float* __restrict__ q, const float* __restrict__ xr, int ub, int d,
                                           float s, float sn, double qq, int64_t row) {
  double score;
  if (COSINE) {
    const double dot = ref_add_tail<REF_DOT>((double)s, q, xr, ub, d);
    const double xx = ref_add_tail<REF_DOT>((double)sn, xr, xr, ub, d);
    score = ref_cosine_from_sums(dot, qq, xx);
  } else {
    const double sum = ref_add_tail<REF_L2SQ>((double)s, q, xr, ub, d);
    score = -__dsqrt_rn(sum);
  }
  return Key{rank_hi_from_score(score), (uint64_t)row};
}


// L2 with d % lanes != 0: the pre-filter compares the fp32 lane part over [0, ub) only.  A row whose ONLY NaN sits in the
// fp64 tail [ub, d) has a finite lane sum and would be dropped once a threshold exists, while the reference scores it
// NaN -- which Double.compare sorts FIRST in descending score order (J/fdb/FdbVectorIndex.java:708).  Rows the filter
// rejects therefore get their (at most lanes - 1) tail elements checked for NaN before they are discarded.
static __device__ __forceinline__ bool scan_tail_has_nan(const float* __restrict__ xr, int ub, int d) {
  bool nan = false;
  for (int i = ub; i < d; i++) {
    const float v = xr[i];
    nan |= v != v;
  }
  return nan;
}

// One batch = (32/TPR) * U consecutive rows starting at row_base, register-staged 128-bit streaming
// loads.  All 32 lanes must call.  qs = the query in shared memory, q = the same query in global memory.
template <int TPR, int U, bool COSINE, class TK>
__device__ __forceinline__ void scan_batch_ldg(const float* __restrict__ X, int64_t n, int d,
                                               const float* __restrict__ q, const float* __restrict__ qs,
                                               const uint8_t* __restrict__ skip, int64_t row_base, int nv, int ub,
                                               double qq, float qn, float fthr, TK& tk, int lane) {
  constexpr int G = 32 / TPR;
  const int t = lane & (TPR - 1);
  const int g = lane / TPR;
    const int64_t row0 = row_base + g;
    float acc[U][4];
    float accn[COSINE ? U : 1][4];
    const float4* rp[U];
#pragma unroll
    for (int u = 0; u < U; u++) {
      int64_t r = row0 + (int64_t)u * G;
      r = r < n ? r : n - 1;
      rp[u] = reinterpret_cast<const float4*>(X + (size_t)r * d) + t;
#pragma unroll
      for (int c = 0; c < 4; c++) acc[u][c] = 0.0f;
      if (COSINE) {
#pragma unroll
        for (int c = 0; c < 4; c++) accn[u][c] = 0.0f;
      }
    }
    for (int i0 = 0; i0 < nv; i0 += SCAN_IV) {
      float4 x[U][SCAN_IV];
#pragma unroll
      for (int u = 0; u < U; u++) {
#pragma unroll
        for (int j = 0; j < SCAN_IV; j++) {
          if (i0 + j < nv) x[u][j] = ld_stream_f4(rp[u] + (size_t)(i0 + j) * TPR);
        }
      }
#pragma unroll
      for (int j = 0; j < SCAN_IV; j++) {
        if (i0 + j < nv) {
          const float4 q4 = reinterpret_cast<const float4*>(qs)[(i0 + j) * TPR + t];
#pragma unroll
          for (int u = 0; u < U; u++) {
            if (COSINE) {
              acc[u][0] = __fmaf_rn(x[u][j].x, q4.x, acc[u][0]);
              acc[u][1] = __fmaf_rn(x[u][j].y, q4.y, acc[u][1]);
              acc[u][2] = __fmaf_rn(x[u][j].z, q4.z, acc[u][2]);
              acc[u][3] = __fmaf_rn(x[u][j].w, q4.w, acc[u][3]);
              accn[u][0] = __fmaf_rn(x[u][j].x, x[u][j].x, accn[u][0]);
              accn[u][1] = __fmaf_rn(x[u][j].y, x[u][j].y, accn[u][1]);
              accn[u][2] = __fmaf_rn(x[u][j].z, x[u][j].z, accn[u][2]);
              accn[u][3] = __fmaf_rn(x[u][j].w, x[u][j].w, accn[u][3]);
            } else {
              // Distances.l2Squared: diff = q - emb (the query is argument a), diff.fma(diff, acc)
              const float dx = __fsub_rn(q4.x, x[u][j].x), dy = __fsub_rn(q4.y, x[u][j].y);
              const float dz = __fsub_rn(q4.z, x[u][j].z), dw = __fsub_rn(q4.w, x[u][j].w);
              acc[u][0] = __fmaf_rn(dx, dx, acc[u][0]);
              acc[u][1] = __fmaf_rn(dy, dy, acc[u][1]);
              acc[u][2] = __fmaf_rn(dz, dz, acc[u][2]);
              acc[u][3] = __fmaf_rn(dw, dw, acc[u][3]);
            }
          }
        }
      }
    }
    // reduceLanes(ADD): ordered ascending-lane fp32 sum, chained through the TPR threads of a row
    float s[U], sn[COSINE ? U : 1];
#pragma unroll
    for (int u = 0; u < U; u++) {
      s[u] = 0.0f;
      if (COSINE) sn[u] = 0.0f;
#pragma unroll
      for (int j = 0; j < TPR; j++) {
        float sin = s[u], snin = COSINE ? sn[u] : 0.0f;
        if (j > 0) {
          sin = __shfl_sync(FULL_MASK, s[u], (lane & ~(TPR - 1)) + j - 1);
          if (COSINE) snin = __shfl_sync(FULL_MASK, sn[u], (lane & ~(TPR - 1)) + j - 1);
        }
        if (t == j) {
          sin = __fadd_rn(sin, acc[u][0]);
          sin = __fadd_rn(sin, acc[u][1]);
          sin = __fadd_rn(sin, acc[u][2]);
          sin = __fadd_rn(sin, acc[u][3]);
          s[u] = sin;
          if (COSINE) {
            snin = __fadd_rn(snin, accn[u][0]);
            snin = __fadd_rn(snin, accn[u][1]);
            snin = __fadd_rn(snin, accn[u][2]);
            snin = __fadd_rn(snin, accn[u][3]);
            sn[u] = snin;
          }
        }
      }
    }
    // pre-filter + exact fp64 score for the survivors
#pragma unroll
    for (int u = 0; u < U; u++) {
      const int64_t row = row0 + (int64_t)u * G;
      float est;
      if (COSINE) {
        est = -(s[u] / (qn * sqrtf(sn[u])));  // NaN when a norm is 0: passes the filter
      } else {
        est = s[u];
      }
      bool cand = (t == TPR - 1) && row < n && !(est > fthr);
      if (!COSINE && ub < d && (t == TPR - 1) && row < n && !cand) cand = scan_tail_has_nan(X + (size_t)row * d, ub, d);
      Key key = key_empty();
      if (cand) {
        if (skip != nullptr && skip[row]) {
          cand = false;
        } else {
          key = scan_exact_key<COSINE>(q, X + (size_t)row * d, ub, d, s[u], COSINE ? sn[u] : 0.0f, qq, row);
        }
      }
      tk.push(key, cand, lane);
    }
}

}  // namespace vs
