// misc.cu -- small kernels: synthetic row generator, the pair operations of the JMH surface.
#include "common.cuh"
#include "kernels.h"

namespace vs {

// ---- java.util.Random on the device (JDK core: 48-bit LCG, fully specified) -----------------------
constexpr uint64_t JR_MULT = 0x5DEECE66DULL;
constexpr uint64_t JR_ADD = 0xBULL;
constexpr uint64_t JR_MASK = (1ULL << 48) - 1;
constexpr int GEN_RUN = 16;  // consecutive draws per thread after one O(log n) skip-ahead

__device__ __forceinline__ uint64_t jr_skip(uint64_t state, uint64_t nsteps) {
  uint64_t a = JR_MULT, c = JR_ADD, acc_a = 1, acc_c = 0;
  while (nsteps) {
    if (nsteps & 1) {
      acc_a = (acc_a * a) & JR_MASK;
      acc_c = (acc_c * a + c) & JR_MASK;
    }
    c = ((a + 1) * c) & JR_MASK;
    a = (a * a) & JR_MASK;
    nsteps >>= 1;
  }
  return (acc_a * state + acc_c) & JR_MASK;
}

// element e = draw (first + e): nextFloat() = next(24) / (float)(1 << 24)
__global__ void generate_kernel(float* __restrict__ out, int64_t count, int64_t seed, int64_t first,
                                int kind) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t e0 = t * GEN_RUN;
  if (e0 >= count) return;
  uint64_t state = ((uint64_t)seed ^ JR_MULT) & JR_MASK;  // Random.initialScramble
  state = jr_skip(state, (uint64_t)(first + e0));
  const int64_t e1 = (e0 + GEN_RUN < count) ? e0 + GEN_RUN : count;
  for (int64_t e = e0; e < e1; e++) {
    state = (state * JR_MULT + JR_ADD) & JR_MASK;
    float f = __fmul_rn((float)(int32_t)(state >> 24), 1.0f / 16777216.0f);
    if (kind == 0)
      f = __fsub_rn(__fmul_rn(f, 2.0f), 1.0f);
    else if (kind == 2)
      f = __fmul_rn(f, 10.0f);
    out[e] = f;
  }
}

cudaError_t launch_generate(float* out, int64_t count, int64_t seed, int64_t first, int kind,
                            cudaStream_t st) {
  if (count <= 0) return cudaSuccess;
  const int64_t threads = (count + GEN_RUN - 1) / GEN_RUN;
  const int block = 256;
  const int64_t grid = (threads + block - 1) / block;
  generate_kernel<<<(unsigned)grid, block, 0, st>>>(out, count, seed, first, kind);
  count_launch();
  return cudaGetLastError();
}

// ---- Distances.* on one pair (J/util/Distances.java:31-153) -----------------------------------------
__global__ void pair_kernel(int op, const float* __restrict__ a, const float* __restrict__ b, int len,
                            int lanes, double* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int hl = lane & 15;
  if (threadIdx.x >= 16) return;
  const unsigned mask = 0x0000ffffu;
  double r;
  if (op == PAIR_L2 || op == PAIR_L2SQ) {
    r = ref_sum_halfwarp<REF_L2SQ>(a, b, len, lanes, hl, mask, 0);
    if (op == PAIR_L2) r = __dsqrt_rn(r);
  } else if (op == PAIR_DOT) {
    r = ref_sum_halfwarp<REF_DOT>(a, b, len, lanes, hl, mask, 0);
  } else if (op == PAIR_NORM) {
    r = __dsqrt_rn(ref_sum_halfwarp<REF_DOT>(a, a, len, lanes, hl, mask, 0));
  } else {
    const double aa = ref_sum_halfwarp<REF_DOT>(a, a, len, lanes, hl, mask, 0);
    const double bb = ref_sum_halfwarp<REF_DOT>(b, b, len, lanes, hl, mask, 0);
    const double ab = ref_sum_halfwarp<REF_DOT>(a, b, len, lanes, hl, mask, 0);
    r = ref_cosine_from_sums(ab, aa, bb);
  }
  if (hl == 0) *out = r;
}

cudaError_t launch_pair(int op, const float* a, const float* b, int len, int lanes, double* out,
                        cudaStream_t st) {
  pair_kernel<<<1, 32, 0, st>>>(op, a, b, len, lanes, out);
  count_launch();
  return cudaGetLastError();
}

// pqLutDistance of the JMH suite: float LUT, float running sum in subspace order
// (B/DistanceAndPqBenchmark.java:116-123)
__global__ void lut_distance_f32_kernel(const float* __restrict__ lut, int M, int K,
                                        const uint8_t* __restrict__ codes, float* __restrict__ out) {
  if (threadIdx.x != 0) return;
  float dist = 0.0f;
  for (int m = 0; m < M; m++) dist = __fadd_rn(dist, lut[(size_t)m * K + codes[m]]);
  *out = dist;
}

cudaError_t launch_lut_distance_f32(const float* lut, int M, int K, const uint8_t* codes, float* out,
                                    cudaStream_t st) {
  lut_distance_f32_kernel<<<1, 32, 0, st>>>(lut, M, K, codes, out);
  count_launch();
  return cudaGetLastError();
}

}  // namespace vs
