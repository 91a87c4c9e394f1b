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
#include <atomic>

#include "common.cuh"
#include "kernels.h"

namespace vs {

constexpr int PQ_THREADS = 512;
// subDim 8 nomination: 2 = tcgen05 (pq_tc.cu, the default where the shape allows), 1 = mma.sync 3xTF32 (below),
// 0 = the FFMA kernel above.  Measured on B200, 10M x 128, M 16, K 256, one pass: tcgen05 16.7 ms (+ 3.4 ms to build
// the operand image, once per training run), FFMA 38 ms, mma.sync 82 ms (the legacy tensor path issues one
// m16n8k8 tf32 mma.sync per ~80 cycles per SM sub-partition; kept, parity-tested, as a reference point).
static std::atomic<int> g_pq_tensor_cores{2};
void pq_set_tensor_cores(int mode) { g_pq_tensor_cores.store(mode); }
constexpr size_t PQ_SMEM_BUDGET = 200 * 1024;

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

// ---- subDim 8: tensor-core nomination -------------------------------------------------------------------
// The production shape (d = 128, M = 16: 8-float sub-vectors, K = 256) makes the assignment a batch of
// [rows x 8] x [8 x K] products: exactly the k = 8 of one tf32 MMA.  Each warp owns 64 rows of one subspace and
// walks the K centroids in chunks of 32 with mma.sync.m16n8k8 (accumulators in registers, where the argmin
// needs them; the kernel is bound by that epilogue -- about five instructions per (row, centroid) -- not by
// the MMA, which is why the accumulators do not take the detour through TMEM here).  3xTF32: x = xh + xl,
// c = ch + cl with tf32-exact halves, <x,c> ~ xh.ch + xh.cl + xl.ch, so the estimate
//     e(row, ci) = |c|^2 + |x|^2 - 2<x,c> (+ a positive margin)
// is good to ~2^-20 relative.  The epilogue keeps, per row, the two smallest keys (estimate with the low 6
// mantissa bits replaced by the thread's column slot) and decides exactly like pq_assign_kernel: if the second
// best is outside the error band of the best, the best IS the reference argmin; otherwise the centroids
// inside the band are evaluated in the reference's own arithmetic, strict '<' in ascending index.
constexpr size_t PQ8_SMEM_BUDGET = 216 * 1024;
constexpr int PQ8_CSTRIDE = 12;  // floats per centroid row in shared memory: B-fragment loads hit 32 distinct banks

__device__ __forceinline__ void mma_tf32_16x8x8(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t tf32_hi(float x) { return __float_as_uint(x) & 0xffffe000u; }
__device__ __forceinline__ uint32_t tf32_lo(float x) { return __float_as_uint(x - __uint_as_float(tf32_hi(x))); }

__global__ void __launch_bounds__(256)
pq_assign8_mma_kernel(const float* __restrict__ X, int64_t n, int d, int M, int K,
                      const float* __restrict__ centroids, int lanes, uint8_t* __restrict__ codes_u8,
                      int32_t* __restrict__ assign_i32, int s_begin, int s_end) {
  extern __shared__ __align__(16) float cs[];  // [ns][K][12] centroids (padded rows), then [ns][K] squared norms
  constexpr int SD = 8;
  const int ns = s_end - s_begin;
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int nw = blockDim.x >> 5;
  const int g = lane >> 2, t = lane & 3;
  float* nrm = cs + (size_t)ns * K * PQ8_CSTRIDE;
  for (int i = threadIdx.x; i < ns * K; i += blockDim.x) {
    const float* src = centroids + ((size_t)s_begin * K + i) * SD;
    float* dst = cs + (size_t)i * PQ8_CSTRIDE;
    float ss = 0.0f;
#pragma unroll
    for (int j = 0; j < SD; j++) {
      const float v = src[j];
      dst[j] = v;
      ss = fmaf(v, v, ss);
    }
    nrm[i] = ss;
  }
  __syncthreads();
  const int64_t ntiles = (n + 63) / 64;
  const int64_t nitems = ntiles * ns;
  const unsigned FINITE_LIM = 0x7f800000u;
  for (int64_t item = (int64_t)blockIdx.x * nw + warp; item < nitems; item += (int64_t)gridDim.x * nw) {
    const int64_t tile = item / ns;
    const int sl = (int)(item % ns);
    const int s = s_begin + sl;
    const float* c0 = cs + (size_t)sl * K * PQ8_CSTRIDE;
    const float* n0 = nrm + (size_t)sl * K;
    // A fragments of the 4 row tiles (rows g and g + 8 of each, components t and t + 4), split into tf32 halves
    uint32_t ah[4][4], al[4][4];
    float xx[4][2];  // |x|^2 of rows (m, g) and (m, g + 8), reduced over the quad
#pragma unroll
    for (int m = 0; m < 4; m++) {
#pragma unroll
      for (int h = 0; h < 2; h++) {
        int64_t row = tile * 64 + m * 16 + h * 8 + g;
        if (row >= n) row = n - 1;
        const float* xr = X + (size_t)row * d + (size_t)s * SD;
        const float v0 = __ldg(xr + t), v1 = __ldg(xr + t + 4);
        ah[m][h] = tf32_hi(v0);
        al[m][h] = tf32_lo(v0);
        ah[m][2 + h] = tf32_hi(v1);
        al[m][2 + h] = tf32_lo(v1);
        float p = fmaf(v0, v0, v1 * v1);
        p += __shfl_xor_sync(FULL_MASK, p, 1);
        p += __shfl_xor_sync(FULL_MASK, p, 2);
        xx[m][h] = p;
      }
    }
    // largest |c|^2 of the subspace bounds the error band (computed once per item by the warp)
    float nmax = 0.0f;
    for (int ci = lane; ci < K; ci += 32) nmax = fmaxf(nmax, n0[ci]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) nmax = fmaxf(nmax, __shfl_xor_sync(FULL_MASK, nmax, o));
    unsigned k1[4][2], k2[4][2];
    float cinit[4][2], band[4][2];
#pragma unroll
    for (int m = 0; m < 4; m++)
#pragma unroll
      for (int h = 0; h < 2; h++) {
        k1[m][h] = 0x7fffffffu;
        k2[m][h] = 0x7fffffffu;
        // |e - reference| <= 0.6 band: dropped xl.cl and the tf32 truncation of xl, cl (3 * 2^-20 (|x|^2 + |c|^2)),
        // fp32 accumulation and norms (2^-21 ...), key truncation (2^-17 relative of e <= 2 (|x|^2 + |c|^2))
        band[m][h] = (xx[m][h] + nmax) * (1.0f / 32768.0f) + 1e-30f;
        // accumulator start: -(|x|^2 + margin)/2, so that |c|^2 - 2 acc = e + margin > 0 whatever the rounding
        cinit[m][h] = -0.5f * (xx[m][h] + 4.0f * band[m][h]);
      }
    const int nchunks = (K + 31) / 32;
    for (int ch = 0; ch < nchunks; ch++) {
      float acc[4][4][4];
#pragma unroll
      for (int m = 0; m < 4; m++)
#pragma unroll
        for (int nt = 0; nt < 4; nt++) {
          acc[m][nt][0] = cinit[m][0];
          acc[m][nt][1] = cinit[m][0];
          acc[m][nt][2] = cinit[m][1];
          acc[m][nt][3] = cinit[m][1];
        }
#pragma unroll
      for (int nt = 0; nt < 4; nt++) {
        int ci = ch * 32 + nt * 8 + g;  // B fragment: centroid ci (column g of the tile), components t and t + 4
        if (ci >= K) ci = K - 1;
        const float b0f = c0[(size_t)ci * PQ8_CSTRIDE + t], b1f = c0[(size_t)ci * PQ8_CSTRIDE + t + 4];
        const uint32_t bh0 = tf32_hi(b0f), bh1 = tf32_hi(b1f), bl0 = tf32_lo(b0f), bl1 = tf32_lo(b1f);
#pragma unroll
        for (int m = 0; m < 4; m++) {
          mma_tf32_16x8x8(acc[m][nt], al[m], bh0, bh1);
          mma_tf32_16x8x8(acc[m][nt], ah[m], bl0, bl1);
          mma_tf32_16x8x8(acc[m][nt], ah[m], bh0, bh1);
        }
      }
#pragma unroll
      for (int nt = 0; nt < 4; nt++) {
        const int col = ch * 32 + nt * 8 + 2 * t;  // this thread's two columns of the tile
        const float2 nc = make_float2(col < K ? n0[col] : __int_as_float(0x7f800000),
                                      col + 1 < K ? n0[col + 1] : __int_as_float(0x7f800000));
        const unsigned slot = (unsigned)(ch * 8 + nt * 2);  // 6-bit column slot within this thread's 64 columns
#pragma unroll
        for (int m = 0; m < 4; m++) {
#pragma unroll
          for (int i = 0; i < 4; i++) {
            const int h = i >> 1;
            const float e = fmaf(-2.0f, acc[m][nt][i], (i & 1) ? nc.y : nc.x);
            const unsigned key = (__float_as_uint(e) & 0xffffffc0u) | (slot + (unsigned)(i & 1));
            const unsigned hi = max(k1[m][h], key);
            k1[m][h] = min(k1[m][h], key);
            k2[m][h] = min(k2[m][h], hi);
          }
        }
      }
    }
    // combine the four threads of a quad (they hold disjoint column sets of the same rows)
#pragma unroll
    for (int m = 0; m < 4; m++)
#pragma unroll
      for (int h = 0; h < 2; h++) {
        // make the slot global before mixing threads: column = chunk * 32 + nt * 8 + 2 t + (i & 1)
        unsigned a1 = k1[m][h], a2 = k2[m][h];
        auto widen = [&](unsigned key) -> unsigned long long {  // (value bits, column) as one ordered 64-bit key
          const unsigned slot = key & 63u;
          const unsigned col = (slot >> 3) * 32 + ((slot >> 1) & 3) * 8 + 2 * t + (slot & 1);
          return ((unsigned long long)(key & 0xffffffc0u) << 32) | col;
        };
        unsigned long long w1 = widen(a1), w2 = widen(a2);
#pragma unroll
        for (int o = 1; o <= 2; o <<= 1) {
          const unsigned long long p1 = __shfl_xor_sync(FULL_MASK, w1, o), p2 = __shfl_xor_sync(FULL_MASK, w2, o);
          const unsigned long long lo1 = min(w1, p1), hi1 = max(w1, p1);
          w2 = min(hi1, min(w2, p2));
          w1 = lo1;
        }
        // every thread of the quad now holds the row's two best; thread t == (m & 3) ... keep it simple: t == 0 decides
        const int64_t row = tile * 64 + m * 16 + h * 8 + g;
        if (t == 0 && row < n) {
          const unsigned v1 = (unsigned)(w1 >> 32), v2 = (unsigned)(w2 >> 32);
          int best = (int)(w1 & 0xffffffffu);
          const bool finite = v1 < FINITE_LIM;
          const float e1 = __uint_as_float(v1), e2 = __uint_as_float(v2);
          if (!finite || !(e2 - e1 > 2.0f * band[m][h])) {
            // near-tie, duplicate centroids, NaN or overflow: decide in reference arithmetic
            const float* xr = X + (size_t)row * d + (size_t)s * SD;
            float x[SD];
#pragma unroll
            for (int j = 0; j < SD; j++) x[j] = xr[j];
            // estimates carry the margin 4 * band; plain fp32 distances do not
            const float lim = finite ? (e1 - 4.0f * band[m][h]) + 3.0f * band[m][h] : __int_as_float(0x7f800000);
            const float* cg = centroids + (size_t)s * K * SD;
            double bestDist = __longlong_as_double(0x7ff0000000000000ll);
            best = 0;
            for (int ci = 0; ci < K; ci++) {
              bool in_band = true;
              if (finite) {
                const float* c = c0 + (size_t)ci * PQ8_CSTRIDE;
                float f0 = 0.0f, f1 = 0.0f;
#pragma unroll
                for (int j = 0; j < SD; j += 2) {
                  const float a = x[j] - c[j], b = x[j + 1] - c[j + 1];
                  f0 = fmaf(a, a, f0);
                  f1 = fmaf(b, b, f1);
                }
                in_band = (f0 + f1) <= lim;
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
          if (codes_u8) codes_u8[(size_t)row * M + s] = (uint8_t)(best & 0xFF);
          if (assign_i32) assign_i32[(size_t)s * n + row] = best;
        }
      }
  }
}

static cudaError_t pq_assign8_mma(const PqAssignLaunch& L, cudaStream_t st) {
  const size_t per_s = (size_t)L.K * (PQ8_CSTRIDE + 1) * 4;
  int chunk = (int)(PQ8_SMEM_BUDGET / per_s);  // 16 subspaces of 256 centroids fit: one pass over the rows
  if (chunk < 1) return cudaErrorInvalidValue;
  int dev = 0, sms = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  cudaError_t e = cudaFuncSetAttribute(pq_assign8_mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)PQ8_SMEM_BUDGET);
  if (e != cudaSuccess) return e;
  for (int s0 = L.s_begin; s0 < L.s_end; s0 += chunk) {
    const int s1 = (s0 + chunk < L.s_end) ? s0 + chunk : L.s_end;
    const size_t smem = (size_t)(s1 - s0) * per_s;
    int occ = 1;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, pq_assign8_mma_kernel, 256, smem);
    if (occ < 1) occ = 1;
    const int64_t nitems = ((L.n + 63) / 64) * (s1 - s0);
    int64_t grid = (int64_t)sms * occ;
    const int64_t need = (nitems + 7) / 8;
    if (grid > need) grid = need;
    if (grid < 1) grid = 1;
    pq_assign8_mma_kernel<<<(unsigned)grid, 256, smem, st>>>(L.X, L.n, L.d, L.M, L.K, L.centroids, L.lanes, L.codes_u8,
                                                           L.assign_i32, s0, s1);
    count_launch();
    if ((e = cudaGetLastError()) != cudaSuccess) return e;
  }
  return cudaSuccess;
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
      case 8: {
        const int mode = g_pq_tensor_cores.load();
        if (mode == 2 && pq_tc_supported(L)) {
          int dev = 0, sms = 0;
          cudaGetDevice(&dev);
          cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
          return launch_pq_assign_tc(L, sms, st);
        }
        return (L.K <= 256 && mode == 1) ? pq_assign8_mma(L, st) : pq_assign_t<8>(L, st);
      }
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
