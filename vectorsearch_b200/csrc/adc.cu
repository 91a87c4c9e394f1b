// adc.cu -- K5/K6: ADC lookup-table build and the full-segment ADC scan over uint8 PQ codes.
//
// Replaces buildLut (J/fdb/FdbVectorIndex.java:1067-1079), pqApproxDistance (:1057-1065) and the
// sealed-segment scan + sort (:754-769, first n_cand of the ascending stable sort :820-822).
//
// HBM-bound on paper (M bytes per distance evaluation), shared-memory-gather-bound in practice.
//  * LUT entries are the reference's own doubles (Distances.l2Squared on sub-vectors; with
//    subDim < SIMD lanes that is pure fp64).  The scan keeps an fp32 image of the LUT in shared
//    memory and sums M fp32 lookups per row: an ESTIMATE whose relative error is bounded
//    (entries are non-negative), used only to discard rows that cannot enter the top n_cand.
//    Rows that survive the fp32 compare are re-summed in fp64 from the real LUT in subspace
//    order, exactly like pqApproxDistance, and ranked by (approx, row) -- so ids and distances
//    are the reference's, ties go to the lowest row.
//  * one thread per code row, 128-bit streaming loads, U rows in flight per thread;
//    grid = SMs x resident CTAs, last CTA merges (same epilogue as scan.cu).
#include "kernels.h"
#include "topk.cuh"

namespace vs {

// ---- LUT build: one thread per (query, subspace, centroid) ----------------------------------------
__global__ void build_lut_kernel(const float* __restrict__ centroids, int M, int K, int subDim,
                                 const float* __restrict__ Q, int nq, int lanes,
                                 double* __restrict__ lut64) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t per_q = (int64_t)M * K;
  if (t >= per_q * nq) return;
  const int qi = (int)(t / per_q);
  const int e = (int)(t % per_q);
  const int s = e / K;
  const float* q = Q + (size_t)qi * M * subDim + (size_t)s * subDim;
  const float* c = centroids + (size_t)e * subDim;
  lut64[t] = ref_sum_thread<REF_L2SQ>(q, c, subDim, lanes);
}

cudaError_t launch_build_lut(const float* centroids, int M, int K, int subDim, const float* q, int nq,
                             int lanes, double* lut64, cudaStream_t st) {
  const int64_t total = (int64_t)M * K * nq;
  if (total <= 0) return cudaSuccess;
  build_lut_kernel<<<(unsigned)((total + 127) / 128), 128, 0, st>>>(centroids, M, K, subDim, q, nq, lanes, lut64);
  count_launch();
  return cudaGetLastError();
}

// pqApproxDistance: fp64 adds in subspace order, codes >= K skipped (:1061)
__device__ __forceinline__ double adc_exact(const double* __restrict__ lut, const uint8_t* __restrict__ cr,
                                            int M, int K) {
  double ad = 0.0;
  for (int s = 0; s < M; s++) {
    const int ci = cr[s];
    if (ci >= K) continue;
    ad = __dadd_rn(ad, lut[(size_t)s * K + ci]);
  }
  return ad;
}

__global__ void approx_distance_kernel(const double* __restrict__ lut, int M, int K,
                                       const uint8_t* __restrict__ codes, int64_t n,
                                       double* __restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = adc_exact(lut, codes + (size_t)i * M, M, K);
}

cudaError_t launch_approx_distance(const double* lut, int M, int K, const uint8_t* codes, int64_t n,
                                   double* out, cudaStream_t st) {
  if (n <= 0) return cudaSuccess;
  approx_distance_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(lut, M, K, codes, n, out);
  count_launch();
  return cudaGetLastError();
}

// BEST_FIRST expansion scoring (J/fdb/FdbVectorIndex.java:950-963): pqApproxDistance of the listed ids against the
// resident codes.  ids outside [id_base, id_base + n) have no code (codeMap.get(nb) == null): valid 0, distance NaN.
__global__ void adc_gather_kernel(const double* __restrict__ lut, int M, int K, const uint8_t* __restrict__ codes, int64_t n,
                                  int64_t id_base, const int64_t* __restrict__ ids, int64_t n_ids,
                                  double* __restrict__ out, uint8_t* __restrict__ valid) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_ids) return;
  const int64_t row = ids[i] - id_base;
  const bool ok = row >= 0 && row < n;
  out[i] = ok ? adc_exact(lut, codes + (size_t)row * M, M, K) : __longlong_as_double(0x7ff8000000000000ll);
  valid[i] = ok ? 1 : 0;
}

cudaError_t launch_adc_gather(const double* lut, int M, int K, const uint8_t* codes, int64_t n, int64_t id_base,
                              const int64_t* ids, int64_t n_ids, double* out, uint8_t* valid, cudaStream_t st) {
  if (n_ids <= 0) return cudaSuccess;
  adc_gather_kernel<<<(unsigned)((n_ids + 127) / 128), 128, 0, st>>>(lut, M, K, codes, n, id_base, ids, n_ids, out, valid);
  count_launch();
  return cudaGetLastError();
}

// fp32 estimate vs exact fp64 sum: each fp32 LUT entry is within 2^-24 relative of the double,
// M-1 fp32 adds of non-negative terms add 2^-24 each -> (M+1) * 2^-24; doubled for slack.
__device__ __forceinline__ float adc_filter_threshold(const Key& thr, int M) {
  if (thr.hi == KEY_EMPTY64) return __int_as_float(0x7f800000);
  const double val = dist_from_rank_hi(thr.hi);
  if (val != val) return __int_as_float(0x7f800000);  // k-th is NaN: every non-NaN row beats it
  const double m = val * (1.0 + (double)(M + 2) * (1.0 / 8388608.0)) + 1e-37;
  const float f = __double2float_ru(m);
  return (f == __int_as_float(0x7f800000)) ? f : f32_next_up(f);
}

constexpr int ADC_KS = 256;  // shared LUT row stride: codes are bytes, entries >= K read 0

// MW = M/4 words per code row (M % 4 == 0), U rows in flight per thread.
template <int MW, int U, class TK>
__global__ void __launch_bounds__(SCAN_THREADS)
adc_scan_kernel(const uint8_t* __restrict__ codes, int64_t n, int K, const double* __restrict__ LUT64,
                int k, int kp, TopkOut out) {
  extern __shared__ __align__(128) ulonglong2 smem[];
  constexpr int M = MW * 4;
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int nw = blockDim.x >> 5;
  const int stride_keys = kp + TOPK_BUF;
  const double* __restrict__ lut64 = LUT64 + (size_t)blockIdx.y * M * K;

  TK tk;
  tk.init(smem + (size_t)warp * stride_keys, kp, k, lane);
  float* lut = reinterpret_cast<float*>(reinterpret_cast<unsigned char*>(smem) + ((topk_block_smem(k, kp, nw) + 15) / 16) * 16);
  for (int i = threadIdx.x; i < M * ADC_KS; i += blockDim.x) {
    const int s = i / ADC_KS, c = i % ADC_KS;
    lut[i] = c < K ? (float)lut64[(size_t)s * K + c] : 0.0f;
  }
  __syncthreads();

  const int64_t rows_per_batch = 32 * U;
  const int64_t nbatches = (n + rows_per_batch - 1) / rows_per_batch;
  const int64_t total_warps = (int64_t)gridDim.x * nw;
  float fthr = __int_as_float(0x7f800000);
  uint64_t seen_hi = KEY_EMPTY64, seen_lo = KEY_EMPTY64;

  for (int64_t b = (int64_t)blockIdx.x * nw + warp; b < nbatches; b += total_warps) {
    const int64_t row0 = b * rows_per_batch + lane;
    uint32_t w[U][MW];
#pragma unroll
    for (int u = 0; u < U; u++) {
      int64_t r = row0 + (int64_t)u * 32;
      r = r < n ? r : n - 1;
      const uint8_t* p = codes + (size_t)r * M;
      if (MW % 4 == 0) {
#pragma unroll
        for (int j = 0; j < MW / 4; j++) {
          uint4 v = ld_stream_u4(reinterpret_cast<const uint4*>(p) + j);
          w[u][4 * j + 0] = v.x; w[u][4 * j + 1] = v.y; w[u][4 * j + 2] = v.z; w[u][4 * j + 3] = v.w;
        }
      } else {
#pragma unroll
        for (int j = 0; j < MW; j++) w[u][j] = __ldg(reinterpret_cast<const uint32_t*>(p) + j);
      }
    }
    float est[U];
#pragma unroll
    for (int u = 0; u < U; u++) {
      float a = 0.0f;
#pragma unroll
      for (int j = 0; j < MW; j++) {
        const uint32_t x = w[u][j];
        a += lut[(4 * j + 0) * ADC_KS + (x & 0xffu)];
        a += lut[(4 * j + 1) * ADC_KS + ((x >> 8) & 0xffu)];
        a += lut[(4 * j + 2) * ADC_KS + ((x >> 16) & 0xffu)];
        a += lut[(4 * j + 3) * ADC_KS + (x >> 24)];
      }
      est[u] = a;
    }
#pragma unroll
    for (int u = 0; u < U; u++) {
      const int64_t row = row0 + (int64_t)u * 32;
      const bool cand = row < n && !(est[u] > fthr);
      Key key = key_empty();
      if (cand) key = Key{rank_hi_from_dist(adc_exact(lut64, codes + (size_t)row * M, M, K)), (uint64_t)row};
      tk.push(key, cand, lane);
    }
    if (tk.thr.hi != seen_hi || tk.thr.lo != seen_lo) {
      seen_hi = tk.thr.hi;
      seen_lo = tk.thr.lo;
      fthr = adc_filter_threshold(tk.thr, M);
    }
  }
  topk_epilogue(tk, smem, kp, k, out);
}

// any M: byte loads, exact fp64 for every row (no estimate).  Correctness path for odd shapes.
template <class TK>
__global__ void __launch_bounds__(SCAN_THREADS)
adc_scan_generic_kernel(const uint8_t* __restrict__ codes, int64_t n, int M, int K,
                        const double* __restrict__ LUT64, int k, int kp, TopkOut out) {
  extern __shared__ __align__(128) ulonglong2 smem[];
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int nw = blockDim.x >> 5;
  const int stride_keys = kp + TOPK_BUF;
  const double* __restrict__ lut64 = LUT64 + (size_t)blockIdx.y * M * K;
  TK tk;
  tk.init(smem + (size_t)warp * stride_keys, kp, k, lane);
  __syncthreads();
  const int64_t nb = (n + 31) / 32;
  for (int64_t b = (int64_t)blockIdx.x * nw + warp; b < nb; b += (int64_t)gridDim.x * nw) {
    const int64_t row = b * 32 + lane;
    Key key = key_empty();
    if (row < n) key = Key{rank_hi_from_dist(adc_exact(lut64, codes + (size_t)row * M, M, K)), (uint64_t)row};
    tk.push(key, row < n, lane);
  }
  topk_epilogue(tk, smem, kp, k, out);
}

template <typename KERN>
static cudaError_t set_smem(KERN kern, size_t smem) {
  if (smem > 48 * 1024)
    return cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  return cudaSuccess;
}

static bool adc_is_streaming(int M, int K) { return K <= 256 && (M == 8 || M == 16 || M == 32 || M == 64); }

typedef void (*AdcKern)(const uint8_t*, int64_t, int, const double*, int, int, TopkOut);
typedef void (*AdcGenKern)(const uint8_t*, int64_t, int, int, const double*, int, int, TopkOut);
template <class TK>
static AdcKern adc_kernel(int M) {
  switch (M) {
    case 8: return adc_scan_kernel<2, 8, TK>;
    case 16: return adc_scan_kernel<4, 4, TK>;
    case 32: return adc_scan_kernel<8, 2, TK>;
    default: return adc_scan_kernel<16, 1, TK>;
  }
}
static AdcKern pick_adc(int M, int k) { return k <= TOPK_REG_MAX_K ? adc_kernel<WarpTopKReg>(M) : adc_kernel<WarpTopK>(M); }
static AdcGenKern pick_adc_generic(int k) {
  return k <= TOPK_REG_MAX_K ? adc_scan_generic_kernel<WarpTopKReg> : adc_scan_generic_kernel<WarpTopK>;
}

bool adc_configure(AdcScanLaunch& L, int sms) {
  L.kp = topk_pad(L.k);
  L.threads = SCAN_THREADS;
  const bool streaming = adc_is_streaming(L.M, L.K);
  const size_t lut = streaming ? (size_t)L.M * ADC_KS * 4 : 0;
  while (((topk_block_smem(L.k, L.kp, L.threads / 32) + 15) / 16) * 16 + lut > 200 * 1024 && L.threads > 32) L.threads /= 2;
  L.smem_bytes = ((topk_block_smem(L.k, L.kp, L.threads / 32) + 15) / 16) * 16 + lut;
  int occ = 0;
  if (streaming) {
    AdcKern kern = pick_adc(L.M, L.k);
    if (set_smem(kern, L.smem_bytes) != cudaSuccess) return false;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, L.threads, L.smem_bytes);
  } else {
    AdcGenKern kern = pick_adc_generic(L.k);
    if (set_smem(kern, L.smem_bytes) != cudaSuccess) return false;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, L.threads, L.smem_bytes);
  }
  if (occ < 1) return false;
  if (occ > 8) occ = 8;
  int64_t grid = (int64_t)sms * (L.nq > 1 ? 1 : occ);
  if (L.k <= TOPK_REG_MAX_K && grid > TOPK_MAX_LISTS) grid = TOPK_MAX_LISTS;
  const int64_t cap = (L.n + 127) / 128;
  if (grid > cap) grid = cap;
  L.grid = (int)(grid < 1 ? 1 : grid);
  L.partial_keys = topk_partial_keys(L.k, L.grid, L.threads / 32);
  return true;
}

cudaError_t launch_adc_scan(const AdcScanLaunch& L, cudaStream_t st) {
  TopkOut o{L.partial, L.ctrl, L.partial_keys, L.ids_out, L.approx_out, L.counts_out, L.id_base, 1, L.out_stride > 0 ? L.out_stride : L.k};
  count_launch();
  cudaError_t e;
  if (!adc_is_streaming(L.M, L.K)) {
    AdcGenKern kern = pick_adc_generic(L.k);
    if ((e = set_smem(kern, L.smem_bytes)) != cudaSuccess) return e;
    kern<<<dim3(L.grid, L.nq), L.threads, L.smem_bytes, st>>>(L.codes, L.n, L.M, L.K, L.lut64, L.k, L.kp, o);
  } else {
    AdcKern kern = pick_adc(L.M, L.k);
    if ((e = set_smem(kern, L.smem_bytes)) != cudaSuccess) return e;
    kern<<<dim3(L.grid, L.nq), L.threads, L.smem_bytes, st>>>(L.codes, L.n, L.K, L.lut64, L.k, L.kp, o);
  }
  return cudaGetLastError();
}

}  // namespace vs
