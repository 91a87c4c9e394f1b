// pq.cu -- K3: PQ assignment (PqEncoder.encode and the PqTrainer assignment step) over n rows.
//
// Replaces J/pq/PqEncoder.java:18-37 (per vector, per subspace: argmin over K centroids of
// Distances.l2Squared on the sub-vector, strict '<' so the lowest centroid index wins ties) and
// the identical loop of J/pq/PqTrainer.java:56-68.
//
// The reference evaluates every sub-distance in its lane arithmetic -- pure fp64 when
// subDim < SIMD lanes (the production shape: subDim 8, 16 lanes).  Doing n*M*K fp64 distances is
// 10x off the fp32 rate, so the kernel NOMINATES in fp32 and decides in reference arithmetic:
//   pass 1: fp32 distance to every centroid (centroids of the processed subspaces live in shared
//           memory and are read as warp broadcasts), tracking the best and second-best estimate;
//   if the second best is outside the rounding band of the best, the best IS the reference argmin;
//   otherwise (near-tie or duplicate centroids) pass 2 re-walks the centroids and evaluates the
//   ones inside the band with the reference's own arithmetic, strict '<' in ascending index.
// Codes are therefore bit-identical to the reference's for every input without NaN/overflow
// special cases (those take the all-exact path).
#include "common.cuh"
#include "kernels.h"

namespace vs {

constexpr int PQ_THREADS = 512;

__device__ __forceinline__ float pq_band(float m1, int SD) {
  // |est - ref| <= (SD + SD/L + L + 4) * 2^-24 relative on both sides -> 3x slack
  const float rel = 3.0f * (float)(2 * SD + 24) * (1.0f / 16777216.0f);
  return m1 * (1.0f + rel) + 1e-30f;
}

template <int SD>
__global__ void __launch_bounds__(PQ_THREADS)
pq_assign_kernel(const float* __restrict__ X, int64_t n, int d, int M, int K,
                 const float* __restrict__ centroids, int lanes, uint8_t* __restrict__ codes_u8,
                 int32_t* __restrict__ assign_i32, int s_begin, int s_end) {
  extern __shared__ __align__(16) float cs[];  // [(s_end - s_begin)][K][SD]
  const int ns = s_end - s_begin;
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int nw = blockDim.x >> 5;
  {
    const float4* src = reinterpret_cast<const float4*>(centroids + (size_t)s_begin * K * SD);
    float4* dst = reinterpret_cast<float4*>(cs);
    const int total4 = ns * K * SD / 4;
    for (int i = threadIdx.x; i < total4; i += blockDim.x) dst[i] = src[i];
  }
  __syncthreads();
  const int64_t ntiles = (n + 31) / 32;
  const int64_t nitems = ntiles * ns;
  for (int64_t item = (int64_t)blockIdx.x * nw + warp; item < nitems; item += (int64_t)gridDim.x * nw) {
    const int64_t tile = item / ns;
    const int sl = (int)(item % ns);
    const int s = s_begin + sl;
    const int64_t row = tile * 32 + lane;
    const bool live = row < n;
    const float* xr = X + (size_t)(live ? row : n - 1) * d + (size_t)s * SD;
    float x[SD];
#pragma unroll
    for (int j = 0; j < SD / 4; j++) {
      float4 v = ld_stream_f4(reinterpret_cast<const float4*>(xr) + j);
      x[4 * j] = v.x; x[4 * j + 1] = v.y; x[4 * j + 2] = v.z; x[4 * j + 3] = v.w;
    }
    const float* c0 = cs + (size_t)sl * K * SD;
    float m1 = __int_as_float(0x7f800000), m2 = m1;
    int i1 = 0;
#pragma unroll 4
    for (int ci = 0; ci < K; ci++) {
      const float4* c = reinterpret_cast<const float4*>(c0 + (size_t)ci * SD);
      float e0 = 0.0f, e1 = 0.0f;
#pragma unroll
      for (int j = 0; j < SD / 4; j++) {
        const float4 cv = c[j];
        const float a = x[4 * j] - cv.x, b = x[4 * j + 1] - cv.y;
        const float g = x[4 * j + 2] - cv.z, h = x[4 * j + 3] - cv.w;
        e0 = fmaf(a, a, e0);
        e1 = fmaf(b, b, e1);
        e0 = fmaf(g, g, e0);
        e1 = fmaf(h, h, e1);
      }
      const float e = e0 + e1;
      if (e < m1) {
        m2 = m1;
        m1 = e;
        i1 = ci;
      } else if (e < m2) {
        m2 = e;
      }
    }
    int best = i1;
    const bool finite = m1 < __int_as_float(0x7f800000);
    if (!finite || !(m2 > pq_band(m1, SD))) {
      // near-tie, duplicate centroids, NaN or overflow: decide in reference arithmetic
      const float lim = finite ? pq_band(m1, SD) : __int_as_float(0x7f800000);
      const float* cg = centroids + (size_t)s * K * SD;
      double bestDist = __longlong_as_double(0x7ff0000000000000ll);
      best = 0;
      for (int ci = 0; ci < K; ci++) {
        bool in_band = true;
        if (finite) {
          const float* c = c0 + (size_t)ci * SD;
          float e0 = 0.0f, e1 = 0.0f;
#pragma unroll
          for (int j = 0; j < SD; j += 2) {
            const float a = x[j] - c[j], b = x[j + 1] - c[j + 1];
            e0 = fmaf(a, a, e0);
            e1 = fmaf(b, b, e1);
          }
          in_band = (e0 + e1) <= lim;
        }
        if (in_band) {
          const double dd = ref_sum_thread<REF_L2SQ>(xr, cg + (size_t)ci * SD, SD, lanes);
          if (dd < bestDist) {  // strict <: lowest ci wins ties, NaN never wins (PqEncoder.java:29)
            bestDist = dd;
            best = ci;
          }
        }
      }
    }
    if (live) {
      if (codes_u8) codes_u8[(size_t)row * M + s] = (uint8_t)(best & 0xFF);
      if (assign_i32) assign_i32[(size_t)s * n + row] = best;
    }
  }
}

// any subDim: reference arithmetic for every (row, subspace, centroid).  Slow correctness path.
__global__ void __launch_bounds__(256)
pq_assign_generic_kernel(const float* __restrict__ X, int64_t n, int d, int M, int K, int SD,
                         const float* __restrict__ centroids, int lanes, uint8_t* __restrict__ codes_u8,
                         int32_t* __restrict__ assign_i32, int s_begin, int s_end) {
  const int ns = s_end - s_begin;
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n * ns) return;
  const int64_t row = t / ns;
  const int s = s_begin + (int)(t % ns);
  const float* xr = X + (size_t)row * d + (size_t)s * SD;
  const float* cg = centroids + (size_t)s * K * SD;
  double bestDist = __longlong_as_double(0x7ff0000000000000ll);
  int best = 0;
  for (int ci = 0; ci < K; ci++) {
    const double dd = ref_sum_thread<REF_L2SQ>(xr, cg + (size_t)ci * SD, SD, lanes);
    if (dd < bestDist) {
      bestDist = dd;
      best = ci;
    }
  }
  if (codes_u8) codes_u8[(size_t)row * M + s] = (uint8_t)(best & 0xFF);
  if (assign_i32) assign_i32[(size_t)s * n + row] = best;
}

constexpr size_t PQ_SMEM_BUDGET = 200 * 1024;

template <int SD>
static cudaError_t pq_assign_t(const PqAssignLaunch& L, cudaStream_t st) {
  auto kern = pq_assign_kernel<SD>;
  const size_t per_s = (size_t)L.K * SD * 4;
  int chunk = (int)(PQ_SMEM_BUDGET / per_s);
  if (chunk < 1) return cudaErrorInvalidValue;
  int dev = 0, sms = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  for (int s0 = L.s_begin; s0 < L.s_end; s0 += chunk) {
    const int s1 = (s0 + chunk < L.s_end) ? s0 + chunk : L.s_end;
    const size_t smem = (size_t)(s1 - s0) * per_s;
    if (smem > 48 * 1024) {
      cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      if (e != cudaSuccess) return e;
    }
    int occ = 1;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, PQ_THREADS, smem);
    if (occ < 1) occ = 1;
    const int64_t nitems = ((L.n + 31) / 32) * (s1 - s0);
    int64_t grid = (int64_t)sms * occ;
    const int64_t need = (nitems + PQ_THREADS / 32 - 1) / (PQ_THREADS / 32);
    if (grid > need) grid = need;
    if (grid < 1) grid = 1;
    kern<<<(unsigned)grid, PQ_THREADS, smem, st>>>(L.X, L.n, L.d, L.M, L.K, L.centroids, L.lanes, L.codes_u8,
                                                   L.assign_i32, s0, s1);
    count_launch();
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
  }
  return cudaSuccess;
}

cudaError_t launch_pq_assign(const PqAssignLaunch& L, cudaStream_t st) {
  if (L.n <= 0 || L.s_end <= L.s_begin) return cudaSuccess;
  const bool aligned = (L.d % 4) == 0;  // sub-vectors start on 16-byte boundaries
  if (aligned && (size_t)L.K * L.subDim * 4 <= PQ_SMEM_BUDGET) {
    switch (L.subDim) {
      case 4: return pq_assign_t<4>(L, st);
      case 8: return pq_assign_t<8>(L, st);
      case 16: return pq_assign_t<16>(L, st);
      case 32: return pq_assign_t<32>(L, st);
      case 48: return pq_assign_t<48>(L, st);
      case 64: return pq_assign_t<64>(L, st);
      default: break;
    }
  }
  const int64_t total = L.n * (L.s_end - L.s_begin);
  pq_assign_generic_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(
      L.X, L.n, L.d, L.M, L.K, L.subDim, L.centroids, L.lanes, L.codes_u8, L.assign_i32, L.s_begin, L.s_end);
  count_launch();
  return cudaGetLastError();
}

}  // namespace vs
