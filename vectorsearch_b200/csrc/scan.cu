// scan.cu -- K1: exact brute-force scan (L2 / cosine) of one query over a resident segment.
//
// Replaces the scoring loop + sort of searchBruteForceSegment
// (J/fdb/FdbVectorIndex.java:676-721): for every live row, Distances.l2(q, emb) or
// Distances.cosine(q, emb), stable sort by score descending, subList(0, k).
//
// HBM-bound streaming kernel (algorithmic traffic: d*4 bytes per distance evaluation).
//  * The arithmetic is the reference's own: Distances.l2Squared / dot / norm keep one fp32
//    accumulator per SIMD lane (L = 16 on an AVX-512 JVM), updated with a fused multiply-add, and
//    reduce the lanes in ascending order (J/util/Distances.java:48-64,103-140).  TPR = L/4
//    threads share a row; thread t owns SIMD lanes 4t..4t+3 and walks the row with 128-bit
//    streaming loads at stride L floats, so every fp32 operation happens in the reference's
//    order and the per-row sum is bit-identical to the JVM's.
//  * A warp has 32/TPR rows x U in flight; one load instruction covers 32/TPR rows x 16*TPR
//    contiguous bytes (full 32-byte sectors), the IV loads of a chunk are issued back to back.
//  * Top-k: one fp32 compare per row against a warp-uniform threshold; only rows that can
//    still enter the top-k get their fp64 score (sqrt / divide) and go through topk.cuh.
//  * grid = SMs x resident CTAs; warps stride over row batches.  The last CTA to finish merges
//    the per-CTA lists and writes ids/scores (no second launch).
#include "kernels.h"
#include "topk.cuh"

namespace vs {

constexpr int SCAN_IV = 8;  // 128-bit loads per row issued back to back

// float threshold for the fp32 pre-filter: a row whose fp32 figure is > fthr cannot beat thr.
template <bool COSINE>
__device__ __forceinline__ float scan_filter_threshold(const Key& thr) {
  if (key_is_empty(thr)) return __int_as_float(0x7f800000);       // +inf: everything passes
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

template <int TPR, int U, bool COSINE>
__global__ void __launch_bounds__(SCAN_THREADS)
scan_kernel(const float* __restrict__ X, int64_t n, int d, const float* __restrict__ Q,
            const uint8_t* __restrict__ skip, int k, int kp, TopkOut out) {
  extern __shared__ __align__(16) ulonglong2 smem[];
  constexpr int L = TPR * 4;
  constexpr int G = 32 / TPR;  // rows per warp per unroll slot
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int nw = blockDim.x >> 5;
  const int stride_keys = kp + TOPK_BUF;
  const int t = lane & (TPR - 1);
  const int g = lane / TPR;
  const float* __restrict__ q = Q + (size_t)blockIdx.y * d;

  WarpTopK tk;
  tk.init(smem + (size_t)warp * stride_keys, kp, k, lane);
  float* qs = reinterpret_cast<float*>(smem + (size_t)nw * stride_keys);
  for (int i = threadIdx.x; i < d; i += blockDim.x) qs[i] = q[i];
  __shared__ double s_qq;
  if (COSINE && threadIdx.x == 0) s_qq = ref_sum_thread_L<L, REF_DOT>(q, q, d);
  __syncthreads();
  const double qq = COSINE ? s_qq : 0.0;
  const float qn = COSINE ? (float)sqrt(qq) : 0.0f;

  const int ub = d - (d % L);
  const int nv = ub / L;  // vector-loop iterations of the modelled JVM
  const int64_t rows_per_batch = (int64_t)G * U;
  const int64_t nbatches = (n + rows_per_batch - 1) / rows_per_batch;
  const int64_t total_warps = (int64_t)gridDim.x * nw;
  float fthr = __int_as_float(0x7f800000);
  uint64_t seen_hi = KEY_EMPTY64, seen_lo = KEY_EMPTY64;

  for (int64_t b = (int64_t)blockIdx.x * nw + warp; b < nbatches; b += total_warps) {
    const int64_t row0 = b * rows_per_batch + g;
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
      Key key = key_empty();
      if (cand) {
        if (skip != nullptr && skip[row]) {
          cand = false;
        } else {
          const float* xr = X + (size_t)row * d;
          double score;
          if (COSINE) {
            const double dot = ref_add_tail<REF_DOT>((double)s[u], q, xr, ub, d);
            const double xx = ref_add_tail<REF_DOT>((double)sn[u], xr, xr, ub, d);
            score = ref_cosine_from_sums(dot, qq, xx);
          } else {
            const double sum = ref_add_tail<REF_L2SQ>((double)s[u], q, xr, ub, d);
            score = -__dsqrt_rn(sum);
          }
          key = Key{rank_hi_from_score(score), (uint64_t)row};
        }
      }
      tk.push(key, cand, lane);
    }
    if (tk.thr.hi != seen_hi || tk.thr.lo != seen_lo) {
      seen_hi = tk.thr.hi;
      seen_lo = tk.thr.lo;
      fthr = scan_filter_threshold<COSINE>(tk.thr);
    }
  }
  topk_epilogue(tk, smem, kp, k, out);
}

// Any d, any lane count: one thread per row, reference arithmetic straight from common.cuh.
// Uncoalesced; used only for shapes the streaming kernel does not take (d % 4 != 0 or d < L).
template <bool COSINE>
__global__ void __launch_bounds__(SCAN_THREADS)
scan_rowthread_kernel(const float* __restrict__ X, int64_t n, int d, const float* __restrict__ Q,
                      const uint8_t* __restrict__ skip, int lanes, int k, int kp, TopkOut out) {
  extern __shared__ __align__(16) ulonglong2 smem[];
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int nw = blockDim.x >> 5;
  const int stride_keys = kp + TOPK_BUF;
  const float* __restrict__ q = Q + (size_t)blockIdx.y * d;
  WarpTopK tk;
  tk.init(smem + (size_t)warp * stride_keys, kp, k, lane);
  __shared__ double s_qq;
  if (COSINE && threadIdx.x == 0) s_qq = ref_sum_thread<REF_DOT>(q, q, d, lanes);
  __syncthreads();
  const double qq = COSINE ? s_qq : 0.0;
  const int64_t nb = (n + 31) / 32;
  for (int64_t b = (int64_t)blockIdx.x * nw + warp; b < nb; b += (int64_t)gridDim.x * nw) {
    const int64_t row = b * 32 + lane;
    bool cand = row < n && !(skip != nullptr && skip[row]);
    Key key = key_empty();
    if (cand) {
      const float* xr = X + (size_t)row * d;
      double score;
      if (COSINE) {
        score = ref_cosine_from_sums(ref_sum_thread<REF_DOT>(q, xr, d, lanes), qq,
                                     ref_sum_thread<REF_DOT>(xr, xr, d, lanes));
      } else {
        score = -__dsqrt_rn(ref_sum_thread<REF_L2SQ>(q, xr, d, lanes));
      }
      key = Key{rank_hi_from_score(score), (uint64_t)row};
    }
    tk.push(key, cand, lane);
  }
  topk_epilogue(tk, smem, kp, k, out);
}

// ------------------------------------------------------------------------------------------------
// host launchers
// ------------------------------------------------------------------------------------------------
template <typename K>
static cudaError_t set_smem(K kern, size_t smem) {
  if (smem > 48 * 1024)
    return cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  return cudaSuccess;
}

template <int TPR, int U, bool COSINE>
static cudaError_t launch_scan_t(const ScanLaunch& L, const TopkOut& o, cudaStream_t st) {
  auto kern = scan_kernel<TPR, U, COSINE>;
  cudaError_t e = set_smem(kern, L.smem_bytes);
  if (e != cudaSuccess) return e;
  kern<<<dim3(L.grid, L.nq), L.threads, L.smem_bytes, st>>>(L.X, L.n, L.d, L.q, L.skip, L.k, L.kp, o);
  count_launch();
  return cudaGetLastError();
}

template <int TPR, int U, bool COSINE>
static int occupancy_t(int threads, size_t smem) {
  auto kern = scan_kernel<TPR, U, COSINE>;
  if (set_smem(kern, smem) != cudaSuccess) return 0;
  int nb = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kern, threads, smem) != cudaSuccess) nb = 0;
  return nb;
}

// the cosine pre-filter needs the whole dot product in the fp32 lanes (no fp64 tail)
bool scan_is_streaming(int d, int lanes, bool cosine) {
  if ((d % 4) != 0 || d < lanes || lanes < 4) return false;
  return !cosine || (d % lanes) == 0;
}

size_t scan_smem_bytes(int d, int kp, int threads) {
  return (size_t)(threads / 32) * topk_warp_smem(kp) + (((size_t)d * 4 + 15) / 16) * 16;
}

int scan_occupancy(int d, int lanes, bool cosine, int threads, size_t smem) {
  if (!scan_is_streaming(d, lanes, cosine)) {
    int nb = 0;
    if (cosine) {
      set_smem(scan_rowthread_kernel<true>, smem);
      cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, scan_rowthread_kernel<true>, threads, smem);
    } else {
      set_smem(scan_rowthread_kernel<false>, smem);
      cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, scan_rowthread_kernel<false>, threads, smem);
    }
    return nb;
  }
  if (cosine) {
    if (lanes == 16) return occupancy_t<4, 1, true>(threads, smem);
    if (lanes == 8) return occupancy_t<2, 1, true>(threads, smem);
    return occupancy_t<1, 1, true>(threads, smem);
  }
  if (lanes == 16) return occupancy_t<4, 2, false>(threads, smem);
  if (lanes == 8) return occupancy_t<2, 2, false>(threads, smem);
  return occupancy_t<1, 2, false>(threads, smem);
}

cudaError_t launch_scan(const ScanLaunch& L, cudaStream_t st) {
  TopkOut o{L.partial, L.ticket, L.ids_out, L.scores_out, L.counts_out, L.id_base, 0};
  if (!scan_is_streaming(L.d, L.lanes, L.cosine)) {
    cudaError_t e;
    if (L.cosine) {
      e = set_smem(scan_rowthread_kernel<true>, L.smem_bytes);
      if (e != cudaSuccess) return e;
      scan_rowthread_kernel<true><<<dim3(L.grid, L.nq), L.threads, L.smem_bytes, st>>>(
          L.X, L.n, L.d, L.q, L.skip, L.lanes, L.k, L.kp, o);
    } else {
      e = set_smem(scan_rowthread_kernel<false>, L.smem_bytes);
      if (e != cudaSuccess) return e;
      scan_rowthread_kernel<false><<<dim3(L.grid, L.nq), L.threads, L.smem_bytes, st>>>(
          L.X, L.n, L.d, L.q, L.skip, L.lanes, L.k, L.kp, o);
    }
    count_launch();
    return cudaGetLastError();
  }
  if (L.cosine) {
    if (L.lanes == 16) return launch_scan_t<4, 1, true>(L, o, st);
    if (L.lanes == 8) return launch_scan_t<2, 1, true>(L, o, st);
    return launch_scan_t<1, 1, true>(L, o, st);
  }
  if (L.lanes == 16) return launch_scan_t<4, 2, false>(L, o, st);
  if (L.lanes == 8) return launch_scan_t<2, 2, false>(L, o, st);
  return launch_scan_t<1, 2, false>(L, o, st);
}

}  // namespace vs
