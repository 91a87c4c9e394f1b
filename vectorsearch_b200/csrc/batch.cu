// batch.cu -- K2: exact brute-force top-k for a BATCH of queries over a resident segment.
//
// Same contract as scan.cu (searchBruteForceSegment, J/fdb/FdbVectorIndex.java:676-721, once per
// query): every returned id and score is produced by the reference arithmetic of scan_rows.cuh.
// What changes is how rows are NOMINATED.  One query streams the segment from HBM (scan.cu); nq
// queries would stream it nq times, so here the segment is read once per 128 queries and the
// dense part -- the nq x n inner products -- runs on the 5th-generation tensor cores:
//
//   (1) row_prep_kernel   one pass at first use: one coefficient c per row so that a(q, x) orders rows like
//                         the metric does (L2: c = |x|^2, a = c - 2<q,x>;  cosine: c = -1/|x|, a = c<q,x>),
//                         plus max |x|^2; row_convert_kernel makes the optional fp16 operand copy.
//   (2) batch_gemm_kernel persistent, warp-specialised tcgen05 kernel: TMA (128-byte swizzle) stages
//                         a 128-query x 32-float block of Q and a 256-row x 32-float block of X per
//                         pipeline stage, one thread issues tcgen05.mma.kind::tf32 (fp32 bits read
//                         as tf32, 128 x 256 x 8 per instruction) into a double-buffered TMEM
//                         accumulator, four epilogue warps read it back with tcgen05.ld (thread =
//                         query, columns = rows) and reduce a(q, x) to one minimum per 64-row group.
//                         Output: gm[query][group], 1/64th of the score matrix.
//   (3) batch_select_kernel per query: T = k-th smallest group minimum (an upper bound of the k-th
//                         smallest a), tau = T + 2 * slack where slack bounds |a - exact| (tf32
//                         truncation: 2^-9 relative on every product); every row of the true top-k
//                         lies in a group whose minimum is <= tau.  Those groups (about k of them)
//                         are re-scored with the reference arithmetic and ranked exactly like scan.cu
//                         does, ties included.
//   (4) batch_fallback_kernel queries that cannot be nominated this way (non-finite query, candidate
//                         list overflow on adversarial data) are re-done by a full exact scan.
//                         Segments holding non-finite rows never enter this path (api.cu).
// Tensor cores only nominate; they never produce a returned value.

#include <mutex>

#include "kernels.h"
#include "scan_rows.cuh"
#include "tc05.cuh"

namespace vs {

constexpr int BQ_M = 128;       // queries per tile = TMEM lanes
constexpr int BQ_N = 256;       // rows per tile = TMEM columns of one accumulator
constexpr int BQ_THREADS = 384;     // warp 0 TMA, warp 1 MMA, warp 2 TMEM allocator, warps 4..11 epilogue
constexpr int BQ_EPI_WARPS = 8;
constexpr uint32_t BQ_A_BYTES = BQ_M * 128;
constexpr uint32_t BQ_B_BYTES = BQ_N * 128;
constexpr uint32_t BQ_STAGE_BYTES = BQ_A_BYTES + BQ_B_BYTES;
constexpr uint32_t BQ_AB_BYTES = BQ_EPI_WARPS * 2 * (BQ_N / 2) * 4;  // row coefficients of the warp's 128 columns, per accumulator
constexpr size_t BQ_GEMM_SMEM_BUDGET = 224 * 1024;  // dynamic shared memory of the nomination kernel
constexpr int BQ_SELECT_THREADS = 256;
constexpr int BQ_FB_SLOTS = 8;  // grid.y of the fallback scan
constexpr int BQ_SEL_MLP = 8;     // float4 loads of group minima in flight per thread of the selection passes
constexpr int BQ_TMIN_MAX_K = 64;  // up to this k the select threshold comes from per-thread minima

// instruction descriptors: D fp32, both operands K-major, N = 256, M = 128; A/B format 2 = tf32, 0 = fp16
constexpr uint32_t BQ_IDESC_TF32 = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(BQ_N >> 3) << 17) | ((uint32_t)(BQ_M >> 4) << 24);
constexpr uint32_t BQ_IDESC_F16 = (1u << 4) | (0u << 7) | (0u << 10) | ((uint32_t)(BQ_N >> 3) << 17) | ((uint32_t)(BQ_M >> 4) << 24);

// ---- (1) per-row nomination coefficients --------------------------------------------------------------------
template <bool COSINE>
__global__ void __launch_bounds__(256)
row_prep_kernel(const float* __restrict__ X, int64_t n, int d, const uint8_t* __restrict__ skip,
                float* __restrict__ coef, SegStats* __restrict__ stats) {
  const int lane = threadIdx.x & 31;
  const int64_t warp0 = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int64_t nwarps = (int64_t)gridDim.x * (blockDim.x >> 5);
  const int nv = d >> 2;
  unsigned int vmax = 0u;
  int bad = 0;
  for (int64_t r = warp0; r < n; r += nwarps) {
    const float4* xr = reinterpret_cast<const float4*>(X + (size_t)r * d);
    float ss = 0.0f;
    for (int i = lane; i < nv; i += 32) {
      const float4 x = ld_stream_f4(xr + i);
      ss = fmaf(x.x, x.x, ss);
      ss = fmaf(x.y, x.y, ss);
      ss = fmaf(x.z, x.z, ss);
      ss = fmaf(x.w, x.w, ss);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(FULL_MASK, ss, o);
    if (lane == 0) {
      const float inf = __int_as_float(0x7f800000);
      float v;
      // the maximum covers skipped rows too: it also fixes the scale of the fp16 operand copy, which outlives
      // a change of the skip mask
      if (ss < 1e30f) vmax = max(vmax, __float_as_uint(ss));
      if (skip != nullptr && skip[r]) {
        v = COSINE ? __int_as_float(0x7fc00000) : inf;  // never nominated (fminf drops NaN)
      } else {
        if (!(ss < 1e30f)) bad = 1;  // NaN, inf or out of the range the slack bound was derived for
        if (COSINE) v = ss > 0.0f ? -(1.0f / sqrtf(ss)) : 0.0f;
        else v = ss;
      }
      coef[r] = v;
    }
  }
  if (lane == 0) {
    if (vmax) atomicMax(&stats->xmax2_bits, vmax);
    if (bad) atomicOr(&stats->nonfinite, 1);
  }
}

cudaError_t launch_row_prep(const float* X, int64_t n, int d, const uint8_t* skip, bool cosine, float* coef,
                            SegStats* stats, int sms, cudaStream_t st) {
  cudaError_t e = cudaMemsetAsync(stats, 0, sizeof(SegStats), st);
  if (e != cudaSuccess) return e;
  int64_t grid = (n + 7) / 8;
  if (grid > (int64_t)sms * 8) grid = (int64_t)sms * 8;
  count_launch();
  if (cosine) row_prep_kernel<true><<<(int)grid, 256, 0, st>>>(X, n, d, skip, coef, stats);
  else row_prep_kernel<false><<<(int)grid, 256, 0, st>>>(X, n, d, skip, coef, stats);
  return cudaGetLastError();
}

// fp16 operand copy of the rows: Xh[r][c] = half(x[r][c] * sx), sx a power of two chosen so that no element
// overflows (|x| * sx <= 2^14); row pitch dp = d rounded up to 8, padding zero.  Nomination only.
__global__ void __launch_bounds__(256)
row_convert_kernel(const float* __restrict__ X, int64_t n, int d, int dp, float sx, __half* __restrict__ Xh) {
  const int cpr = dp >> 2;  // 4-element chunks per row
  const int64_t total = n * cpr;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / cpr;
    const int c = (int)(i - r * cpr) * 4;
    float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
    if (c < d) x = ld_stream_f4(reinterpret_cast<const float4*>(X + (size_t)r * d + c));
    const __half2 lo = __floats2half2_rn(x.x * sx, x.y * sx), hi = __floats2half2_rn(x.z * sx, x.w * sx);
    uint2 o;
    o.x = *reinterpret_cast<const unsigned int*>(&lo);
    o.y = *reinterpret_cast<const unsigned int*>(&hi);
    *reinterpret_cast<uint2*>(Xh + (size_t)r * dp + c) = o;
  }
}
cudaError_t launch_row_convert(const float* X, int64_t n, int d, int dp, float sx, void* Xh, int sms, cudaStream_t st) {
  int64_t grid = (n * (dp >> 2) + 255) / 256;
  if (grid > (int64_t)sms * 16) grid = (int64_t)sms * 16;
  count_launch();
  row_convert_kernel<<<(int)(grid < 1 ? 1 : grid), 256, 0, st>>>(X, n, d, dp, sx, static_cast<__half*>(Xh));
  return cudaGetLastError();
}

// Queries: Qh[q][c] = half(q[c] * sq) with a per-query power of two sq (|q| * sq <= 2^14), and
// qinv[q] = 1 / (sx * sq), the factor that brings the accumulator back to <q, x>.  One warp per query.
__global__ void __launch_bounds__(256)
query_convert_kernel(const float* __restrict__ Q, int nq, int nq_pad, int d, int dp, float sx, __half* __restrict__ Qh,
                     float* __restrict__ qinv) {
  pdl_trigger();
  const int lane = threadIdx.x & 31;
  const int qi = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (qi >= nq_pad) return;
  if (qi >= nq) {
    if (lane == 0) qinv[qi] = 0.0f;
    return;
  }
  const float* q = Q + (size_t)qi * d;
  float ss = 0.0f;
  for (int i = lane; i < d; i += 32) ss = fmaf(q[i], q[i], ss);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(FULL_MASK, ss, o);
  float sq = 1.0f;
  if (ss > 0.0f && ss < 1e30f) {
    int e;
    frexpf(sqrtf(ss) * 1.0001f, &e);  // |q| < 2^e
    sq = ldexpf(1.0f, 14 - e);
  }
  for (int i = lane; i < dp; i += 32) Qh[(size_t)qi * dp + i] = __float2half_rn(i < d ? q[i] * sq : 0.0f);
  if (lane == 0) qinv[qi] = 1.0f / (sx * sq);
}

// Epilogue core shared by the nomination kernels: reduce this thread's 128 accumulator columns (TMEM address
// taddr, columns taddr .. taddr + 127 of its lane) to (BQ_N / 2) / GROUP minima of a(q, x).  The four 32-column
// TMEM loads are software-pipelined: load h + 1 is in flight while load h is reduced.
template <bool COSINE, int GROUP>
__device__ __forceinline__ void bq_reduce_columns(uint32_t taddr, const float* __restrict__ myc, float m2, float qs,
                                                  float (&mins)[(BQ_N / 2) / GROUP]) {
  constexpr int NM = (BQ_N / 2) / GROUP;
  const float inf = __int_as_float(0x7f800000);
#pragma unroll
  for (int i = 0; i < NM; i++) mins[i] = inf;
  uint32_t v[2][32];
  tc_ld32(taddr, v[0]);
#pragma unroll
  for (int h = 0; h < 4; h++) {
    tc_wait_ld();
    tc_pin32(v[h & 1]);
    if (h < 3) tc_ld32(taddr + (uint32_t)(h + 1) * 32, v[(h + 1) & 1]);
    const float4* p4 = reinterpret_cast<const float4*>(myc + h * 32);
    float pm[2];  // columns 0..15 and 16..31 of this load
#pragma unroll
    for (int hh = 0; hh < 2; hh++) {
      float m0 = inf, m1 = inf;
#pragma unroll
      for (int j = 0; j < 4; j++) {
        const float4 p = p4[hh * 4 + j];  // coefficients of four rows; same address in every lane: broadcast
        const int c = hh * 16 + 4 * j;
        if (COSINE) {  // a = c <q,x>; the per-query scale is applied after the minimum
          m0 = fminf(m0, __uint_as_float(v[h & 1][c]) * p.x);
          m1 = fminf(m1, __uint_as_float(v[h & 1][c + 1]) * p.y);
          m0 = fminf(m0, __uint_as_float(v[h & 1][c + 2]) * p.z);
          m1 = fminf(m1, __uint_as_float(v[h & 1][c + 3]) * p.w);
        } else {       // a = c - 2 <q,x>
          m0 = fminf(m0, fmaf(__uint_as_float(v[h & 1][c]), m2, p.x));
          m1 = fminf(m1, fmaf(__uint_as_float(v[h & 1][c + 1]), m2, p.y));
          m0 = fminf(m0, fmaf(__uint_as_float(v[h & 1][c + 2]), m2, p.z));
          m1 = fminf(m1, fmaf(__uint_as_float(v[h & 1][c + 3]), m2, p.w));
        }
      }
      pm[hh] = fminf(m0, m1);
    }
    if (GROUP == 16) {
      mins[(2 * h) % NM] = pm[0];
      mins[(2 * h + 1) % NM] = pm[1];
    } else if (GROUP == 32) {
      mins[h % NM] = fminf(pm[0], pm[1]);
    } else {
      mins[(h / 2) % NM] = fminf(mins[(h / 2) % NM], fminf(pm[0], pm[1]));
    }
  }
  if (COSINE) {
#pragma unroll
    for (int i = 0; i < NM; i++) mins[i] *= qs;
  }
}

// ---- (2) tensor-core nomination ---------------------------------------------------------------------------------
// grid.x = nqb * nsplit; CTA (qb, split) owns query block qb and row tiles split, split + nsplit, ...
// (CTAs of the same split run side by side, so a row tile is fetched from HBM once and re-read from L2).
// STAT: the CTA's 128-query block (num_kb k-blocks of 16 KB) is loaded once and stays in shared memory, the
// pipeline stages carry row blocks only -- a third less L2 -> SM traffic, which is what bounds this kernel
// at d = 128.  Otherwise (long vectors) every stage carries the k-block of both operands.
constexpr int BQ_MAX_STAGES = 8;
// HALF: operands are the fp16 copies (k-block = 64 halfs, kind::f16, twice the tensor rate and half the bytes);
// qinv[query] undoes the power-of-two scaling of both copies.  Otherwise the fp32 data itself, read as tf32.
template <bool STAT, bool HALF, bool COSINE, int GROUP>
__global__ void __launch_bounds__(BQ_THREADS, 1)
batch_gemm_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmX,
                  const float* __restrict__ coef, const float* __restrict__ qinv, int64_t n, int num_kb, int nstages,
                  int nqb, int64_t tiles, float* __restrict__ gm, int64_t gm_stride) {
  // the swizzled operand tiles need 1024-byte alignment; no integer round-trip on the pointer, so every access
  // below stays a shared-memory instruction (LDS / STS) instead of a generic one
  extern __shared__ __align__(1024) uint8_t bq_smem[];
  uint8_t* base = bq_smem;
  if ((smem_u32(base) & 1023u) != 0) __trap();
  constexpr uint32_t STAGE_BYTES = STAT ? BQ_B_BYTES : BQ_STAGE_BYTES;
  constexpr int KB_ELEMS = HALF ? 64 : 32;  // elements per 128-byte k-block
  uint8_t* aq = base;                                                    // STAT: [num_kb][16 KB] query block
  uint8_t* stages = base + (STAT ? (size_t)num_kb * BQ_A_BYTES : 0);
  float* sm_coef = reinterpret_cast<float*>(stages + (size_t)nstages * STAGE_BYTES);
  uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(sm_coef) + BQ_AB_BYTES);
  uint64_t* full = bars;                          // [BQ_MAX_STAGES] TMA -> MMA
  uint64_t* empty = bars + BQ_MAX_STAGES;         // [BQ_MAX_STAGES] MMA -> TMA
  uint64_t* tfull = bars + 2 * BQ_MAX_STAGES;     // [2] MMA -> epilogue
  uint64_t* tempty = bars + 2 * BQ_MAX_STAGES + 2;  // [2] epilogue -> MMA
  uint64_t* afull = bars + 2 * BQ_MAX_STAGES + 4;   // [1] resident query block has landed
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(bars + 2 * BQ_MAX_STAGES + 5);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int qb = blockIdx.x % nqb;
  const int split = blockIdx.x / nqb;
  const int nsplit = gridDim.x / nqb;

  if (threadIdx.x == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmQ)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmX)) : "memory");
    for (int s = 0; s < nstages; s++) {
      mbar_init(full + s, 1);
      mbar_init(empty + s, 1);
    }
    for (int a = 0; a < 2; a++) {
      mbar_init(tfull + a, 1);
      mbar_init(tempty + a, BQ_EPI_WARPS);
    }
    mbar_init(afull, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(s_tmem)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(s_tmem);
  pdl_trigger();  // the selection kernel may be set up while this one runs
  pdl_wait();     // scaled queries / scratch reset of the launches before

  if (warp == 0) {
    if (lane == 0) {  // ===== TMA producer =====
      if (STAT) {
        mbar_expect_tx(afull, (uint32_t)num_kb * BQ_A_BYTES);
        for (int kb = 0; kb < num_kb; kb++) tma_load_2d(aq + (size_t)kb * BQ_A_BYTES, &tmQ, kb * KB_ELEMS, qb * BQ_M, afull);
      }
      int s = 0;
      uint32_t ph = 0;
      for (int64_t tile = split; tile < tiles; tile += nsplit) {
        for (int kb = 0; kb < num_kb; kb++) {
          mbar_wait(empty + s, ph ^ 1);
          mbar_expect_tx(full + s, STAGE_BYTES);
          uint8_t* st = stages + (size_t)s * STAGE_BYTES;
          if (!STAT) tma_load_2d(st, &tmQ, kb * KB_ELEMS, qb * BQ_M, full + s);
          tma_load_2d(st + (STAT ? 0 : BQ_A_BYTES), &tmX, kb * KB_ELEMS, (int)(tile * BQ_N), full + s);
          if (++s == nstages) {
            s = 0;
            ph ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {  // ===== MMA issuer =====
      int s = 0, it = 0;
      uint32_t ph = 0;
      if (STAT) mbar_wait(afull, 0);
      for (int64_t tile = split; tile < tiles; tile += nsplit, it++) {
        const int a = it & 1;
        mbar_wait(tempty + a, ((it >> 1) & 1) ^ 1);  // the epilogue has drained this accumulator
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)a * BQ_N;
        for (int kb = 0; kb < num_kb; kb++) {
          mbar_wait(full + s, ph);
          tc_fence_after();
          const uint32_t sb = smem_u32(stages + (size_t)s * STAGE_BYTES) + (STAT ? 0u : BQ_A_BYTES);
          const uint32_t sa = STAT ? smem_u32(aq + (size_t)kb * BQ_A_BYTES) : smem_u32(stages + (size_t)s * STAGE_BYTES);
          const uint64_t adesc = umma_desc_sw128(sa), bdesc = umma_desc_sw128(sb);
#pragma unroll
          for (int kk = 0; kk < 4; kk++) {  // 8 tf32 / 16 fp16 = 32 bytes per instruction: +2 in descriptor units
            if (HALF) tc_mma_f16(d_tmem, adesc + 2 * kk, bdesc + 2 * kk, BQ_IDESC_F16, (kb | kk) != 0 ? 1u : 0u);
            else tc_mma_tf32(d_tmem, adesc + 2 * kk, bdesc + 2 * kk, BQ_IDESC_TF32, (kb | kk) != 0 ? 1u : 0u);
          }
          tc_commit(empty + s);  // frees the stage once these MMAs have read it
          if (++s == nstages) {
            s = 0;
            ph ^= 1;
          }
        }
        tc_commit(tfull + a);
      }
    }
  } else if (warp >= 4) {  // ===== epilogue: TMEM -> per-group minima =====
    const int ew = warp - 4;
    const int lq = ew & 3;    // TMEM lane quarter this warp may read (= warp % 4)
    const int ch = ew >> 2;   // which 128 of the tile's 256 columns
    const float inf = __int_as_float(0x7f800000);
    const float qs = HALF ? qinv[qb * BQ_M + lq * 32 + lane] : 1.0f;  // exact power of two, > 0
    const float m2 = -2.0f * qs;
    const float dead = COSINE ? __int_as_float(0x7fc00000) : inf;
    // the warp's 128 coefficients of the NEXT tile travel in registers while this tile is reduced
    float cnext[4];
#pragma unroll
    for (int i = 0; i < 4; i++) {
      const int64_t r = (int64_t)split * BQ_N + ch * (BQ_N / 2) + i * 32 + lane;
      cnext[i] = (split < tiles && r < n) ? __ldg(coef + r) : dead;
    }
    int it = 0;
    for (int64_t tile = split; tile < tiles; tile += nsplit, it++) {
      const int a = it & 1;
      float* myc = sm_coef + (size_t)(ew * 2 + a) * (BQ_N / 2);
#pragma unroll
      for (int i = 0; i < 4; i++) myc[i * 32 + lane] = cnext[i];
      __syncwarp();
      {
        const int64_t nt = tile + nsplit;
#pragma unroll
        for (int i = 0; i < 4; i++) {
          const int64_t r = nt * BQ_N + ch * (BQ_N / 2) + i * 32 + lane;
          cnext[i] = (nt < tiles && r < n) ? __ldg(coef + r) : dead;
        }
      }
      mbar_wait(tfull + a, (it >> 1) & 1);
      tc_fence_after();
      constexpr int NM = (BQ_N / 2) / GROUP;  // minima this thread produces per tile (GROUP = 16 / 32 / 64 rows)
      float mins[NM];
      bq_reduce_columns<COSINE, GROUP>(tmem_base + ((uint32_t)(lq * 32) << 16) + (uint32_t)(a * BQ_N + ch * (BQ_N / 2)), myc, m2, qs, mins);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty + a);
      float* dst = gm + (size_t)(qb * BQ_M + lq * 32 + lane) * gm_stride + tile * (BQ_N / GROUP) + ch * NM;
      if (NM == 2) {
        *reinterpret_cast<float2*>(dst) = make_float2(mins[0], mins[1]);
      } else {
#pragma unroll
        for (int i = 0; i < NM; i += 4) *reinterpret_cast<float4*>(dst + i) = make_float4(mins[i], mins[(i + 1) % NM], mins[(i + 2) % NM], mins[(i + 3) % NM]);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
  }
}

// ---- (2b) the same nomination on CTA PAIRS (tcgen05 cta_group::2) ---------------------------------------------------
// Batches of more than 128 queries over short vectors.  Two CTAs on the SMs of one TPC form a cluster: each keeps
// its own 128-query block resident and loads HALF of every 256-row tile; one thread of the leader CTA issues
// tcgen05.mma.cta_group::2 (M = 256: both query blocks, N = 256), which reads A and its half of B from each
// CTA's shared memory and writes each CTA's 128 accumulator lanes into that CTA's TMEM.  Per flop that halves
// the L2 -> SM traffic of the row tiles and cuts the operand reads from shared memory by a third -- the two
// things suspected of bounding batch_gemm_kernel at d = 128 (tensor pipe 59 % active).  Measured (r1, 1M x 128,
// 1024 queries): 258 us against 250 us for the single-CTA kernel, and both drop to the same 212 us with the epilogue
// reduction stubbed out -- so neither operand path is the limiter at this shape.  Off by default ("batch_pairs"),
// parity-tested, kept as the starting point for deeper MMA/epilogue decoupling.
//   full[s] / afull live in the leader: both CTAs' TMA loads complete_tx on the leader's barrier (peer bit masked);
//   empty[s], tfull[a]: tcgen05.commit multicasts the arrival to both CTAs; tempty[a]: the epilogue warps of both
//   CTAs arrive on the leader's barrier (mapa + remote arrive).
constexpr int BQ_PAIR_MAX_STAGES = 12;
constexpr uint32_t BQ_BH_BYTES = (BQ_N / 2) * 128;  // one k-block of this CTA's half of the row tile
constexpr uint32_t BQ_IDESC_TF32_M256 = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(BQ_N >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
constexpr uint32_t BQ_IDESC_F16_M256 = (1u << 4) | (0u << 7) | (0u << 10) | ((uint32_t)(BQ_N >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// 2-SM TMA load: the bytes are credited to the LEADER CTA's barrier at the same shared-memory offset
__device__ __forceinline__ void tma_load_2d_pair(void* dst, const CUtensorMap* tm, int c0, int c1, uint64_t* bar) {
  asm volatile("cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
               ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(tm)), "r"(smem_u32(bar) & 0xFEFFFFFFu), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void tc_commit_pair(uint64_t* bar) {  // arrives on this barrier in BOTH CTAs
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(smem_u32(bar)), "h"((uint16_t)3) : "memory");
}
// arrive on the leader CTA's copy of `bar` (plain arrive: a cluster-scope release would add MEMBAR.ALL.GPU per tile;
// the TMEM reads it orders are fenced by tcgen05.fence::before_thread_sync)
__device__ __forceinline__ void mbar_arrive_leader(uint64_t* bar) {
  asm volatile("{\n.reg .b32 ra;\nmapa.shared::cluster.u32 ra, %0, %1;\nmbarrier.arrive.shared::cluster.b64 _, [ra];\n}"
               ::"r"(smem_u32(bar)), "r"(0) : "memory");
}
__device__ __forceinline__ void tc_mma_pair(bool half, uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  if (half)
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n}\n"
                 ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
  else
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n}\n"
                 ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}

// STAT as above; otherwise (long vectors: the query block does not fit beside the stages) every stage carries one
// k-block of the CTA's query block AND of its half of the row tile -- 32 KB per CTA for the flops the single-CTA
// kernel moves 48 KB for, which is what matters at d = 768 where that kernel runs into the L2 -> SM rate.
template <bool STAT, bool HALF, bool COSINE, int GROUP>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(BQ_THREADS, 1)
batch_gemm_pair_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmX /* 128-row box */,
                       const float* __restrict__ coef, const float* __restrict__ qinv, int64_t n, int num_kb, int nstages,
                       int nqp /* pairs of query blocks */, int64_t tiles, float* __restrict__ gm, int64_t gm_stride) {
  extern __shared__ __align__(1024) uint8_t bq_smem[];
  uint8_t* base = bq_smem;
  if ((smem_u32(base) & 1023u) != 0) __trap();
  constexpr int KB_ELEMS = HALF ? 64 : 32;
  constexpr uint32_t STAGE_BYTES = STAT ? BQ_BH_BYTES : BQ_A_BYTES + BQ_BH_BYTES;
  uint8_t* aq = base;                                   // STAT: [num_kb][16 KB] this CTA's query block
  // [nstages] this CTA's half of the row tile (16 KB), preceded by the query block's k-block when both stream
  uint8_t* stages = base + (STAT ? (size_t)num_kb * BQ_A_BYTES : 0);
  float* sm_coef = reinterpret_cast<float*>(stages + (size_t)nstages * STAGE_BYTES);
  uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(sm_coef) + BQ_AB_BYTES);
  uint64_t* full = bars;                                  // leader's copy is the live one
  uint64_t* empty = bars + BQ_PAIR_MAX_STAGES;            // one per CTA
  uint64_t* tfull = bars + 2 * BQ_PAIR_MAX_STAGES;        // one per CTA
  uint64_t* tempty = bars + 2 * BQ_PAIR_MAX_STAGES + 2;   // leader's copy is the live one
  uint64_t* afull = bars + 2 * BQ_PAIR_MAX_STAGES + 4;    // leader's copy is the live one
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(bars + 2 * BQ_PAIR_MAX_STAGES + 5);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int rank = (int)cluster_ctarank();
  const int pair = blockIdx.x >> 1;
  const int qp = pair % nqp;
  const int split = pair / nqp;
  const int nsplit = (gridDim.x >> 1) / nqp;
  const int qb = 2 * qp + rank;

  if (threadIdx.x == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmQ)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmX)) : "memory");
    for (int s = 0; s < nstages; s++) {
      mbar_init(full + s, 1);
      mbar_init(empty + s, 1);
    }
    for (int a = 0; a < 2; a++) {
      mbar_init(tfull + a, 1);
      mbar_init(tempty + a, 2 * BQ_EPI_WARPS);
    }
    mbar_init(afull, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(s_tmem)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  cluster_sync_all();  // both CTAs' barriers are initialised before anything signals them
  tc_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(s_tmem);
  pdl_trigger();
  pdl_wait();

  if (warp == 0) {
    if (lane == 0) {  // ===== TMA producer (both CTAs) =====
      if (STAT) {
        if (rank == 0) mbar_expect_tx(afull, 2u * (uint32_t)num_kb * BQ_A_BYTES);
        for (int kb = 0; kb < num_kb; kb++) tma_load_2d_pair(aq + (size_t)kb * BQ_A_BYTES, &tmQ, kb * KB_ELEMS, qb * BQ_M, afull);
      }
      int s = 0;
      uint32_t ph = 0;
      for (int64_t tile = split; tile < tiles; tile += nsplit) {
        for (int kb = 0; kb < num_kb; kb++) {
          mbar_wait(empty + s, ph ^ 1);
          if (rank == 0) mbar_expect_tx(full + s, 2u * STAGE_BYTES);
          uint8_t* st = stages + (size_t)s * STAGE_BYTES;
          if (!STAT) tma_load_2d_pair(st, &tmQ, kb * KB_ELEMS, qb * BQ_M, full + s);
          tma_load_2d_pair(st + (STAT ? 0 : BQ_A_BYTES), &tmX, kb * KB_ELEMS, (int)(tile * BQ_N + rank * (BQ_N / 2)), full + s);
          if (++s == nstages) {
            s = 0;
            ph ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0 && rank == 0) {  // ===== MMA issuer (leader CTA only) =====
      int s = 0, it = 0;
      uint32_t ph = 0;
      if (STAT) mbar_wait(afull, 0);
      for (int64_t tile = split; tile < tiles; tile += nsplit, it++) {
        const int a = it & 1;
        mbar_wait(tempty + a, ((it >> 1) & 1) ^ 1);  // both CTAs' epilogues have drained this accumulator
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)a * BQ_N;
        for (int kb = 0; kb < num_kb; kb++) {
          mbar_wait(full + s, ph);
          tc_fence_after();
          const uint32_t sst = smem_u32(stages + (size_t)s * STAGE_BYTES);
          const uint64_t adesc = umma_desc_sw128(STAT ? smem_u32(aq + (size_t)kb * BQ_A_BYTES) : sst);
          const uint64_t bdesc = umma_desc_sw128(sst + (STAT ? 0u : BQ_A_BYTES));
#pragma unroll
          for (int kk = 0; kk < 4; kk++)
            tc_mma_pair(HALF, d_tmem, adesc + 2 * kk, bdesc + 2 * kk, HALF ? BQ_IDESC_F16_M256 : BQ_IDESC_TF32_M256, (kb | kk) != 0 ? 1u : 0u);
          tc_commit_pair(empty + s);
          if (++s == nstages) {
            s = 0;
            ph ^= 1;
          }
        }
        tc_commit_pair(tfull + a);
      }
    }
  } else if (warp >= 4) {  // ===== epilogue (both CTAs, each on its own 128 accumulator lanes) =====
    const int ew = warp - 4;
    const int lq = ew & 3;
    const int ch = ew >> 2;
    const float inf = __int_as_float(0x7f800000);
    const float qs = HALF ? qinv[qb * BQ_M + lq * 32 + lane] : 1.0f;
    const float m2 = -2.0f * qs;
    const float dead = COSINE ? __int_as_float(0x7fc00000) : inf;
    float cnext[4];
#pragma unroll
    for (int i = 0; i < 4; i++) {
      const int64_t r = (int64_t)split * BQ_N + ch * (BQ_N / 2) + i * 32 + lane;
      cnext[i] = (split < tiles && r < n) ? __ldg(coef + r) : dead;
    }
    int it = 0;
    for (int64_t tile = split; tile < tiles; tile += nsplit, it++) {
      const int a = it & 1;
      float* myc = sm_coef + (size_t)(ew * 2 + a) * (BQ_N / 2);
#pragma unroll
      for (int i = 0; i < 4; i++) myc[i * 32 + lane] = cnext[i];
      __syncwarp();
      {
        const int64_t nt = tile + nsplit;
#pragma unroll
        for (int i = 0; i < 4; i++) {
          const int64_t r = nt * BQ_N + ch * (BQ_N / 2) + i * 32 + lane;
          cnext[i] = (nt < tiles && r < n) ? __ldg(coef + r) : dead;
        }
      }
      mbar_wait(tfull + a, (it >> 1) & 1);
      tc_fence_after();
      constexpr int NM = (BQ_N / 2) / GROUP;  // minima this thread produces per tile (GROUP = 16 / 32 / 64 rows)
      float mins[NM];
      bq_reduce_columns<COSINE, GROUP>(tmem_base + ((uint32_t)(lq * 32) << 16) + (uint32_t)(a * BQ_N + ch * (BQ_N / 2)), myc, m2, qs, mins);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_leader(tempty + a);
      float* dst = gm + (size_t)(qb * BQ_M + lq * 32 + lane) * gm_stride + tile * (BQ_N / GROUP) + ch * NM;
      if (NM == 2) {
        *reinterpret_cast<float2*>(dst) = make_float2(mins[0], mins[1]);
      } else {
#pragma unroll
        for (int i = 0; i < NM; i += 4) *reinterpret_cast<float4*>(dst + i) = make_float4(mins[i], mins[(i + 1) % NM], mins[(i + 2) % NM], mins[(i + 3) % NM]);
      }
    }
  }
  tc_fence_before();
  cluster_sync_all();  // nobody leaves (or frees TMEM) while the peer may still signal or read
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
  }
}

// ---- candidate groups: fp16 pre-filter, then exact scores ------------------------------------------------------------
// A candidate group holds gsz rows of which one or two matter.  With the fp16 operand copy at hand the warp first
// recomputes a(q, x) for the group's rows from that copy (half the bytes of the fp32 rows; exact products, fp32
// sums: the same slack covers it) and only the rows with a <= tau -- the condition every row of the true top-k
// satisfies -- are fetched in fp32 and scored with the reference arithmetic, one row per half-warp.
struct HalfOperands {
  const __half* Xh;    // [n][dp] scaled fp16 rows, nullptr: no copy (tf32 operands), whole groups are scored exactly
  const __half* Qh;    // [nq][dp] scaled fp16 queries
  const float* qinv;   // [nq] 1 / (row scale * query scale)
  const float* coef;   // [n] row coefficients of the metric
  int dp;
};

constexpr int BQ_PLIST = 128;  // survivors a warp collects before it scores them (ints of warp-private shared memory)

// Rolling L2 prefetch of a warp's candidate rows, eight rows (one pre-filter round) at a time and c_bq_pf_ahead rounds
// ahead of the loads.  (Asking for every candidate group of every query up front, as this kernel once did, is
// fine at d = 128 but pushes ~1 GB through a 126 MB L2 at d = 768, k = 50: most lines were evicted before use
// and fetched from HBM twice.)
__constant__ int c_bq_pf_ahead = 1;
int batch_set_prefetch_rounds(int rounds) { return (int)cudaMemcpyToSymbol(c_bq_pf_ahead, &rounds, sizeof(int)); }
struct PrefetchCursor {
  const int* list;
  const char* xb;   // fp16 rows
  int64_t rbytes;   // bytes per row
  int64_t n;
  int gi, step, cnt, gsz, r0;
  __device__ __forceinline__ void advance(int lane) {  // request the cursor's round, move on
    if (gi >= cnt) return;
    if (lane == 0) {
      const int64_t row = (int64_t)list[gi] * gsz + r0;
      int64_t rows = n - row;
      if (rows > 8) rows = 8;
      if (rows > gsz - r0) rows = gsz - r0;
      if (rows > 0)
        asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(xb + row * rbytes), "r"((uint32_t)(rows * rbytes)) : "memory");
    }
    r0 += 8;
    if (r0 >= gsz) {
      r0 = 0;
      gi += step;
    }
  }
};

// Pre-filter of candidate group `gi` (rows row0 .. row0 + gsz): appends (gi << 6 | row in group) of every row with
// a <= tau to plist and asks L2 for the fp32 row, so that the exact scores (bq_score_survivors, after several
// groups) do not wait on HBM row by row.  Four lanes per row, eight rows per round, up to eight 16-byte loads in
// flight per lane: the rows come from L2 (bulk prefetch of the whole group) and the loop is pure latency.
template <bool COSINE>
__device__ __forceinline__ int bq_group_prefilter(const HalfOperands& H, const float* __restrict__ X, int64_t n, int d,
                                                  const float* __restrict__ qf /* smem [dp] */, float qi_scale, float tau,
                                                  int64_t row0, int gsz, int gi, int* __restrict__ plist, int pcnt,
                                                  PrefetchCursor& pf, int lane) {
  const int t = lane & 3, g = lane >> 2;
  const int chunks = H.dp >> 3;  // 16-byte chunks (8 halfs) per row
  for (int r0 = 0; r0 < gsz; r0 += 8) {
    pf.advance(lane);
    const int64_t row = row0 + r0 + g;
    const bool live = row < n && r0 + g < gsz;
    float dot = 0.0f;
    if (live) {
      const uint4* xr = reinterpret_cast<const uint4*>(H.Xh + (size_t)row * H.dp);
      float d0 = 0.0f, d1 = 0.0f;
      for (int c0 = t; c0 < chunks; c0 += 32) {
        uint4 v[8];
#pragma unroll
        for (int j = 0; j < 8; j++) {
          const int c = c0 + 4 * j;
          v[j] = c < chunks ? __ldg(xr + c) : make_uint4(0u, 0u, 0u, 0u);
        }
#pragma unroll
        for (int j = 0; j < 8; j++) {
          const int c = c0 + 4 * j;
          if (c < chunks) {
            const float4 qa = *reinterpret_cast<const float4*>(qf + c * 8), qb = *reinterpret_cast<const float4*>(qf + c * 8 + 4);
            const float2 x0 = __half22float2(*reinterpret_cast<const __half2*>(&v[j].x)), x1 = __half22float2(*reinterpret_cast<const __half2*>(&v[j].y));
            const float2 x2 = __half22float2(*reinterpret_cast<const __half2*>(&v[j].z)), x3 = __half22float2(*reinterpret_cast<const __half2*>(&v[j].w));
            d0 = fmaf(x0.x, qa.x, d0);
            d1 = fmaf(x0.y, qa.y, d1);
            d0 = fmaf(x1.x, qa.z, d0);
            d1 = fmaf(x1.y, qa.w, d1);
            d0 = fmaf(x2.x, qb.x, d0);
            d1 = fmaf(x2.y, qb.y, d1);
            d0 = fmaf(x3.x, qb.z, d0);
            d1 = fmaf(x3.y, qb.w, d1);
          }
        }
      }
      dot = d0 + d1;
    }
    dot += __shfl_xor_sync(FULL_MASK, dot, 1);
    dot += __shfl_xor_sync(FULL_MASK, dot, 2);
    bool pass = false;
    if (live && t == 0) {
      const float c = __ldg(H.coef + row);
      const float a = COSINE ? (c * dot) * qi_scale : fmaf(dot, -2.0f * qi_scale, c);
      pass = a <= tau;  // dead rows carry +inf / NaN coefficients and never pass
    }
    const unsigned m = __ballot_sync(FULL_MASK, pass);
    if (pass) {
      plist[pcnt + __popc(m & ((1u << lane) - 1u))] = (gi << 6) | (r0 + g);
      if ((d & 3) == 0)
        asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(X + (size_t)row * d), "r"((uint32_t)d * 4u) : "memory");
    }
    pcnt += __popc(m);
  }
  return pcnt;
}

// exact scores of the survivors: one row per half-warp, lane hl is SIMD lane hl of the modelled JVM
template <bool COSINE, class TK>
__device__ __forceinline__ void bq_score_survivors(const float* __restrict__ X, int d, const float* __restrict__ q,
                                                   const uint8_t* __restrict__ skip, int lanes, double qq,
                                                   const int* __restrict__ list, int gsz, const int* __restrict__ plist, int pcnt,
                                                   TK& tk, int lane) {
  __syncwarp();
  const int hl = lane & 15, hw = lane >> 4;
  const unsigned hmask = (lane < 16) ? 0x0000ffffu : 0xffff0000u;
  const int base_lane = lane & 16;
  for (int i = 0; i < pcnt; i += 2) {
    const bool have = i + hw < pcnt;
    Key key = key_empty();
    bool ok = false;
    if (have) {  // uniform across the half-warp
      const int e = plist[i + hw];
      const int64_t row = (int64_t)list[e >> 6] * gsz + (e & 63);
      ok = !(skip != nullptr && skip[row]);
      if (ok) {
        const float* x = X + (size_t)row * d;
        double score;
        if (COSINE) {
          const double dotx = ref_sum_halfwarp<REF_DOT>(q, x, d, lanes, hl, hmask, base_lane);
          const double xx = ref_sum_halfwarp<REF_DOT>(x, x, d, lanes, hl, hmask, base_lane);
          score = ref_cosine_from_sums(dotx, qq, xx);
        } else {
          score = -__dsqrt_rn(ref_sum_halfwarp<REF_L2SQ>(q, x, d, lanes, hl, hmask, base_lane));
        }
        key = Key{rank_hi_from_score(score), (uint64_t)row};
      }
    }
    tk.push(key, ok && hl == 0, lane);
  }
  __syncwarp();
}

// all of a warp's candidate groups list[first], list[first + step], ...: pre-filter, then exact scores in batches
template <bool COSINE, class TK>
__device__ __forceinline__ void bq_groups_prefilter_exact(const HalfOperands& H, const float* __restrict__ X, int64_t n, int d,
                                                          const float* __restrict__ q, const float* __restrict__ qf, float qi_scale,
                                                          const uint8_t* __restrict__ skip, int lanes, double qq, float tau,
                                                          const int* __restrict__ list, int first, int step, int cnt, int gsz,
                                                          int* __restrict__ plist /* warp smem [BQ_PLIST] */, TK& tk, int lane) {
  int pcnt = 0;
  PrefetchCursor pf{list, reinterpret_cast<const char*>(H.Xh), (int64_t)H.dp * 2, n, first, step, cnt, gsz, 0};
  const int ahead = c_bq_pf_ahead;
  if (ahead < 0) pf.gi = cnt;  // no prefetch
  for (int i = 0; i < ahead; i++) pf.advance(lane);
  for (int gi = first; gi < cnt; gi += step) {
    if (pcnt + gsz > BQ_PLIST) {
      bq_score_survivors<COSINE, TK>(X, d, q, skip, lanes, qq, list, gsz, plist, pcnt, tk, lane);
      pcnt = 0;
    }
    pcnt = bq_group_prefilter<COSINE>(H, X, n, d, qf, qi_scale, tau, (int64_t)list[gi] * gsz, gsz, gi, plist, pcnt, pf, lane);
  }
  bq_score_survivors<COSINE, TK>(X, d, q, skip, lanes, qq, list, gsz, plist, pcnt, tk, lane);
}

// ---- (3) threshold, candidate groups, exact ranking --------------------------------------------------------------
// |a(q, x) - exact| <= slack.  Operand rounding: tf32 truncates both operands (2^-10 each), the fp16 copies are
// rounded to nearest (2^-11 each, plus a subnormal floor that the power-of-two scaling keeps below 2^-37 relative);
// d * 2^-22 covers the tensor core's fp32 accumulation: twice the worst case of one truncation (round toward zero,
// 2^-23 relative to the running sum) per accumulated product.  tests/test_gpu_parity.py::test_batch_nomination_bound
// measures the actual deviation against this bound through vs_debug_batch_groupmins.
__host__ __device__ __forceinline__ double batch_slack(bool cosine, bool half, int d, double xmax, double qn) {
  const double c1 = (half ? (1.0 / 1024.0) * 1.01 + sqrt((double)d) * (1.0 / 68719476736.0) : (1.0 / 512.0) * 1.02) +
                    (double)d * (1.0 / 4194304.0);
  const double c2 = (double)(d + 64) * (1.0 / 8388608.0);                  // fp32 rounding of alpha, of a, of the reference sums
  if (cosine) return (c1 + c2) * qn;
  return 2.0 * c1 * xmax * qn + c2 * (xmax + qn) * (xmax + qn);
}

// fb = [count][nq query indices][nq flags]: list query qi once, whichever CTA asks first
__device__ __forceinline__ void batch_list_fallback(int32_t* fb, int nq_total, int qi) {
  if (atomicExch(fb + 1 + nq_total + qi, 1) == 0) fb[1 + atomicAdd(fb, 1)] = qi;
}

// development-only phase timestamps of the selection kernel (build with -DVS_BQ_STAMPS; tools/select_stamps.py)
#ifdef VS_BQ_STAMPS
static __device__ unsigned long long g_bq_stamps[8 * 1024];
__device__ __forceinline__ void bq_stamp(int ph) {
  if (threadIdx.x == 0 && blockIdx.x == 0 && blockIdx.y < 1024) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    g_bq_stamps[blockIdx.y * 8 + ph] = t;
  }
}
int debug_read_stamps_batch(void* dst, size_t bytes) {
  cudaDeviceSynchronize();
  return (int)cudaMemcpyFromSymbol(dst, g_bq_stamps, bytes);
}
__device__ __forceinline__ void bq_stamp_any(int ph) {  // whichever CTA runs this phase (the last CTA of a query)
  if (threadIdx.x == 0 && blockIdx.y < 1024) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    g_bq_stamps[blockIdx.y * 8 + ph] = t;
  }
}
// scan_half_kernel: up to 32 phases per query at g_bq_stamps[4096 + 32 * query + phase]; CTA 0's thread 0, or (any = true)
// thread 0 of whichever CTA runs the phase
__device__ __forceinline__ void sh_stamp(int ph, bool any = false) {
  if (threadIdx.x == 0 && (any || blockIdx.x == 0) && blockIdx.y < 2) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    g_bq_stamps[4096 + blockIdx.y * 32 + ph] = t;
  }
}
// ... and per CTA of query 0: g_bq_stamps[1024 + 2 * CTA + phase] (0 = start, 1 = warp 0 left the loop)
__device__ __forceinline__ void sh_stamp_cta(int ph) {
  if (threadIdx.x == 0 && blockIdx.y == 0 && blockIdx.x < 1024) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    g_bq_stamps[1024 + blockIdx.x * 2 + ph] = t;
  }
}
#else
__device__ __forceinline__ void bq_stamp(int) {}
__device__ __forceinline__ void bq_stamp_any(int) {}
__device__ __forceinline__ void sh_stamp(int, bool = false) {}
__device__ __forceinline__ void sh_stamp_cta(int) {}
#endif

// ---- (2c) one or two queries: nomination by a CUDA-core scan of the fp16 copy -----------------------------------------
// A single query cannot amortise the tensor-core tile, but it does not have to read the fp32 rows either: the same
// a(q, x) the nomination GEMM produces is computed here from the fp16 operand copy (exact fp16 products, fp32 sums --
// the slack of the batched path covers any summation order), i.e. from HALF the bytes scan.cu streams, and only the
// rows that can be in the top-k are scored from the fp32 rows in the reference's arithmetic.
//   * per warp: a private TMA ring over [TR rows of Xh | their TR coefficients], eight lanes per row (a quarter-warp
//     reads 128 contiguous bytes of one row: no bank conflicts), four rows per step, the warp's KK smallest (a, row) in
//     registers (one 64-bit key per lane, insertion by ballot + shuffle);
//   * per CTA: the KK smallest of its warps' keys are published;
//   * last CTA: B = k-th smallest list head (k distinct rows at or below it), tau = B + 2 * slack; every published key
//     with a <= tau is a candidate and is scored exactly (one row per half-warp) into the register top-k of scan.cu.
//     A row that a warp or a CTA dropped has a >= that CTA's KK-th key, so the result is complete iff every FULL list
//     ends above tau; otherwise (rows of one cluster stored side by side, thousands of duplicates) the query goes to
//     the exact fallback scan, like a candidate overflow of the batched selection (the scan writes a flag per query,
//     batch_fallback_kernel is launched behind every scan and exits at once when no flag is set).
//   * sixteen warps per SM wherever their rings fit: as two 256-thread CTAs (k <= 16: a CTA's set-up and epilogue
//     then overlap the neighbour's streaming) or one 512-thread CTA, else eight (long vectors); see batch_configure.
constexpr int SH_THREADS = 512;   // the largest CTA: 16 warps (the scan is latency-bound per warp)
constexpr int SH_MIN_THREADS = 256;
constexpr int SH_CAND = 1024;  // candidate rows the last CTA can hold
// CTAs per query (published lists the last CTA takes): two 256-thread CTAs per SM for short lists (k <= 16), see
// batch_configure; the last CTA holds every published key in registers, so long lists keep one CTA per SM
__host__ __device__ constexpr int sh_max_lists(int kk) { return kk == 16 ? 320 : 160; }
__device__ __forceinline__ uint32_t sh_ord(float f) {
  const uint32_t u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float sh_unord(uint32_t o) { return __uint_as_float((o & 0x80000000u) ? (o & 0x7fffffffu) : ~o); }
// d0 += x.lo * q.lo, d1 += x.hi * q.hi for the two fp16 values packed in each word: fma.rn.f32.f16 (SASS FHFMA, with
// .H0 / .H1 operand selectors) multiplies two fp16 values exactly and adds in fp32 with one rounding -- no conversion
// instructions, one issue slot per element
__device__ __forceinline__ void sh_fh2(float& d0, float& d1, uint32_t x, uint32_t q) {
  asm("{\n.reg .b16 xl, xh, ql, qh;\nmov.b32 {xl, xh}, %2;\nmov.b32 {ql, qh}, %3;\n"
      "fma.rn.f32.f16 %0, xl, ql, %0;\nfma.rn.f32.f16 %1, xh, qh, %1;\n}"
      : "+f"(d0), "+f"(d1) : "r"(x), "r"(q));
}
__device__ __forceinline__ void sh_fh8(const uint4 v, const uint4 q, float& d0, float& d1) {
  sh_fh2(d0, d1, v.x, q.x);
  sh_fh2(d0, d1, v.y, q.y);
  sh_fh2(d0, d1, v.z, q.z);
  sh_fh2(d0, d1, v.w, q.w);
}

// CPL > 0: the lane's CPL 16-byte chunks of the scaled query live in registers (dp <= 64 * CPL); 0: in shared memory.
template <bool COSINE, int KK, int CPL>
__global__ void __launch_bounds__(SH_THREADS, 1)
scan_half_kernel(const __half* __restrict__ Xh, int64_t n, int dp, const float* __restrict__ coef, float x_scale,
                 const SegStats* __restrict__ stats, const float* __restrict__ X, int d, const float* __restrict__ Q,
                 const uint8_t* __restrict__ skip, int lanes, int k, int TR, int NS, int32_t* __restrict__ fb, int nq_total,
                 unsigned long long* __restrict__ pub, TopkOut out) {
  extern __shared__ __align__(128) unsigned char shm[];
  pdl_trigger();
  sh_stamp(0);
  sh_stamp_cta(0);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  const int t = lane & 7, g = lane >> 3;
  const int qi = blockIdx.y;
  const uint32_t row_bytes = (uint32_t)dp * 2u;
  const uint32_t rows_bytes = (uint32_t)TR * row_bytes;
  const uint32_t stage_bytes = (rows_bytes + (uint32_t)TR * 4u + 127u) & ~127u;
  unsigned char* ring = shm + (size_t)warp * NS * stage_bytes;
  float* qf = reinterpret_cast<float*>(shm + (size_t)nw * NS * stage_bytes);
  uint64_t* bars = reinterpret_cast<uint64_t*>(qf + ((dp + 3) & ~3)) + warp * NS;
  const float* __restrict__ q = Q + (size_t)qi * d;

  if (lane == 0) {
    for (int s = 0; s < NS; s++) mbar_init(bars + s, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncwarp();
  const int64_t ntiles = (n + TR - 1) / TR;
  const int64_t first = (int64_t)blockIdx.x * nw + warp, step = (int64_t)gridDim.x * nw;
  const int64_t whole_tiles = n / TR;  // tiles below this index have all TR rows
  // lane 0: rows (and, for a whole tile, coefficients) of `tile` into stage s.  Whole tiles -- all but the segment's last --
  // take the short path: constant sizes, source addresses advanced by constants (the general form costs ~75
  // instructions of 64-bit arithmetic per tile, a quarter of what scoring the tile's sixteen rows takes)
  const uint32_t ring_u32 = smem_u32(ring), bars_u32 = smem_u32(bars);
  const size_t x_step = (size_t)step * TR * row_bytes, c_step = (size_t)step * TR * 4;
  const unsigned char* x_next = reinterpret_cast<const unsigned char*>(Xh) + (size_t)first * TR * row_bytes;  // of the next tile to request
  const unsigned char* c_next = reinterpret_cast<const unsigned char*>(coef) + (size_t)first * TR * 4;
  auto request = [&](int64_t tile, int s) {
    const uint32_t dst = ring_u32 + (uint32_t)s * stage_bytes, bar = bars_u32 + (uint32_t)s * 8u;
    if (tile < whole_tiles) {
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(rows_bytes + (uint32_t)TR * 4u) : "memory");
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                   ::"r"(dst), "l"(x_next), "r"(rows_bytes), "r"(bar) : "memory");
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                   ::"r"(dst + rows_bytes), "l"(c_next), "r"((uint32_t)TR * 4u), "r"(bar) : "memory");
    } else {
      const uint32_t rb = (uint32_t)(n - tile * TR) * row_bytes;
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(rb) : "memory");
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                   ::"r"(dst), "l"(x_next), "r"(rb), "r"(bar) : "memory");
    }
    x_next += x_step;
    c_next += c_step;
  };
  // (nothing read up to the epilogue depends on the predecessor in the stream: the ring fills at once)
  if (lane == 0)
    for (int s = 0; s < NS; s++)
      if (first + (int64_t)s * step < ntiles) request(first + (int64_t)s * step, s);
  sh_stamp(1);
  __shared__ float s_q2;
  __shared__ double s_qq, s_slack2;
  if (warp == 0) {
    float ss = 0.0f;
    for (int i = lane; i < d; i += 32) ss = fmaf(q[i], q[i], ss);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(FULL_MASK, ss, o);
    if (lane == 0) s_q2 = ss;
  }
  __syncthreads();
  // the query scaled by a power of two and rounded to fp16, like the rows (query_convert_kernel's arithmetic, done
  // by every CTA for itself: one launch and one dependency less on the latency path of a query)
  sh_stamp(2);
  const float q2 = s_q2;
  // twice the nomination slack of this query (the last CTA's candidate band): off the epilogue's critical path
  if (threadIdx.x == blockDim.x - 1)
    s_slack2 = 2.0 * batch_slack(COSINE, true, d, sqrt((double)__uint_as_float(stats->xmax2_bits)), sqrt((double)q2));
  float sq = 1.0f;
  if (q2 > 0.0f && q2 < 1e30f) {
    int e;
    frexpf(sqrtf(q2) * 1.0001f, &e);  // |q| < 2^e
    sq = ldexpf(1.0f, 14 - e);
  }
  const float qis = 1.0f / (x_scale * sq);
  uint4 qr[CPL > 0 ? CPL : 1];   // the lane's chunks t, t + 8, ... of the scaled fp16 query (zero beyond d)
  int xoff[CPL > 0 ? CPL : 1];   // byte offsets of those chunks inside a row (clamped: the query chunk is zero there)
  const int chunks = dp >> 3;
  uint4* qh_s = reinterpret_cast<uint4*>(qf);  // CPL == 0: the whole scaled fp16 query in shared memory
  auto q_chunk = [&](int c) -> uint4 {  // elements 8c .. 8c + 7
    float f[8];
#pragma unroll
    for (int i = 0; i < 8; i++) f[i] = (c * 8 + i < d) ? q[c * 8 + i] * sq : 0.0f;
    const __half2 h0 = __floats2half2_rn(f[0], f[1]), h1 = __floats2half2_rn(f[2], f[3]);
    const __half2 h2 = __floats2half2_rn(f[4], f[5]), h3 = __floats2half2_rn(f[6], f[7]);
    uint4 o;
    o.x = *reinterpret_cast<const unsigned int*>(&h0);
    o.y = *reinterpret_cast<const unsigned int*>(&h1);
    o.z = *reinterpret_cast<const unsigned int*>(&h2);
    o.w = *reinterpret_cast<const unsigned int*>(&h3);
    return o;
  };
  if (CPL > 0) {
#pragma unroll
    for (int j = 0; j < (CPL > 0 ? CPL : 1); j++) {
      const int c = t + 8 * j;
      qr[j] = c < chunks ? q_chunk(c) : make_uint4(0u, 0u, 0u, 0u);
      xoff[j] = (c < chunks ? c : chunks - 1) * 16;
    }
  } else {
    for (int i = threadIdx.x; i < chunks; i += blockDim.x) qh_s[i] = q_chunk(i);
  }

  __syncthreads();
  sh_stamp(3);
  const bool bad_query = !(q2 < 1e30f);  // non-finite (or absurdly large): exact scan instead (uniform over the grid)

  const float inf = __int_as_float(0x7f800000);
  // lane i: the warp's i-th smallest a (order-preserving bits) and its row; 0xffffffff = empty
  uint32_t my_a = 0xffffffffu, my_r = 0xffffffffu, thr_ord = 0xffffffffu;
  float thr = inf;
  // What the CTA publishes is the KK smallest keys of ALL its warps, so a row at or above ANY warp's KK-th value is of
  // no use to any of them: the warps share the smallest such value (s_cta_thr) and filter with it even before their own
  // list is full -- a tenth of the insertions sixteen private thresholds cost.
  __shared__ unsigned int s_cta_thr;
  if (threadIdx.x == 0) s_cta_thr = 0xff800000u;  // +inf
  __syncthreads();
  const float m2q = -2.0f * qis;
  // <q, x> of the row whose first byte is at rowp, summed over the row's eight lanes (every one of them gets the sum)
  auto row_dot = [&](const unsigned char* rowp) -> float {
    float d0 = 0.0f, d1 = 0.0f;
    if (CPL > 0) {
      uint4 v[CPL > 0 ? CPL : 1];
#pragma unroll
      for (int j = 0; j < (CPL > 0 ? CPL : 1); j++) v[j] = *reinterpret_cast<const uint4*>(rowp + xoff[j]);
#pragma unroll
      for (int j = 0; j < (CPL > 0 ? CPL : 1); j++) sh_fh8(v[j], qr[j], d0, d1);
    } else {
#pragma unroll 4
      for (int c = t; c < chunks; c += 8) sh_fh8(*reinterpret_cast<const uint4*>(rowp + c * 16), qh_s[c], d0, d1);
    }
    return d0 + d1;
  };
  auto insert = [&](unsigned m, float a, uint32_t row) {  // lanes of mask m offer (a, row)
    const uint32_t before = thr_ord;
    while (m) {
      const int src = __ffs(m) - 1;
      m &= m - 1;
      const uint32_t ca = __shfl_sync(FULL_MASK, sh_ord(a), src), cr = __shfl_sync(FULL_MASK, row, src);
      if (!(ca < thr_ord)) continue;  // the threshold may have tightened since the ballot
      const int pos = __popc(__ballot_sync(FULL_MASK, my_a <= ca));  // sorted: the smaller keys are a prefix
      const uint32_t ua = __shfl_up_sync(FULL_MASK, my_a, 1), ur = __shfl_up_sync(FULL_MASK, my_r, 1);
      if (lane > pos) {
        my_a = ua;
        my_r = ur;
      } else if (lane == pos) {
        my_a = ca;
        my_r = cr;
      }
      thr_ord = __shfl_sync(FULL_MASK, my_a, KK - 1);
    }
    if (thr_ord != before) {
      thr = sh_unord(thr_ord);
      if (lane == 0) atomicMin(&s_cta_thr, thr_ord);
    }
  };
  int stage = 0;
  uint32_t parity = 0;
  for (int64_t tile = first; tile < ntiles; tile += step) {
    mbar_wait_suspend(bars + stage, parity);
    const unsigned char* sp = ring + (size_t)stage * stage_bytes;
    const float* cf = reinterpret_cast<const float*>(sp + rows_bytes);
    const uint32_t row0 = (uint32_t)(tile * TR);
    if (!bad_query) {
      if ((tile + 1) * TR <= n) {  // whole tile: coefficients came with the rows, no row is out of range
        const unsigned char* lp = sp + (size_t)g * row_bytes;
        for (int r0 = 0; r0 < TR; r0 += 8, lp += 8 * row_bytes) {  // two sub-steps of four rows: independent chains
          const float thr_eff = fminf(thr, sh_unord(*reinterpret_cast<volatile unsigned int*>(&s_cta_thr)));
          float da = row_dot(lp), db = row_dot(lp + 4 * row_bytes);
          da += __shfl_xor_sync(FULL_MASK, da, 1);
          db += __shfl_xor_sync(FULL_MASK, db, 1);
          da += __shfl_xor_sync(FULL_MASK, da, 2);
          db += __shfl_xor_sync(FULL_MASK, db, 2);
          da += __shfl_xor_sync(FULL_MASK, da, 4);
          db += __shfl_xor_sync(FULL_MASK, db, 4);
          const float ca = cf[r0 + g], cb = cf[r0 + 4 + g];
          const float aa = COSINE ? (ca * da) * qis : fmaf(da, m2q, ca);
          const float ab = COSINE ? (cb * db) * qis : fmaf(db, m2q, cb);
          // (dead rows carry +inf / NaN coefficients and never pass)
          const unsigned ma = __ballot_sync(FULL_MASK, t == 0 && aa < thr_eff), mb = __ballot_sync(FULL_MASK, t == 0 && ab < thr_eff);
          if (ma | mb) {
            insert(ma, aa, row0 + r0 + g);
            insert(mb, ab, row0 + r0 + 4 + g);
          }
        }
      } else {  // the segment's last, partial tile
        for (int r0 = 0; r0 < TR; r0 += 4) {
          const int ri = r0 + g;
          const int64_t row = tile * TR + ri;
          float da = row_dot(sp + (size_t)ri * row_bytes);
          da += __shfl_xor_sync(FULL_MASK, da, 1);
          da += __shfl_xor_sync(FULL_MASK, da, 2);
          da += __shfl_xor_sync(FULL_MASK, da, 4);
          float aa = inf;
          if (row < n) {
            const float ca = __ldg(coef + row);
            aa = COSINE ? (ca * da) * qis : fmaf(da, m2q, ca);
          }
          insert(__ballot_sync(FULL_MASK, t == 0 && row < n && aa < thr), aa, row0 + ri);  // (dead rows: +inf / NaN)
        }
      }
    }
    __syncwarp();
    if (lane == 0) {
      const int64_t nt = tile + (int64_t)NS * step;
      if (nt < ntiles) {
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        request(nt, stage);
      }
    }
    if (++stage == NS) {
      stage = 0;
      parity ^= 1;
    }
  }
  sh_stamp(4);
  sh_stamp_cta(1);
  pdl_wait();  // nothing above wrote global memory; the scratch and the outputs may be in use by earlier kernels until here
  if (bad_query) {
    if (blockIdx.x == 0 && threadIdx.x == 0) fb[qi] = 1;  // (every launch writes the flag of every query: nothing to zero)
    return;
  }

  sh_stamp(5);
  // ---- epilogue (the rings are dead: their shared memory is reused) ---------------------------------------------------
  __shared__ int s_m;
  __shared__ unsigned int s_pos[SH_THREADS / 32];
  // The CTA publishes the KK smallest keys of its warps' (sorted) lists.  All of them lie at or below T = min(V, the
  // smallest KK-th key of any warp): with c = nw / 2 and p = KK / c - 1, V = the c-th smallest of the warps' (p + 1)-th
  // keys has c warps with p + 1 keys each at or below it, i.e. KK keys (rows are distinct).  Only the keys at or below T
  // -- two or three dozen of the nw * KK -- are compacted (one shared-memory atomic per warp) and ranked.
  const int cw = nw >= 2 ? nw / 2 : 1, pp = KK / cw - 1;
  if (threadIdx.x == 0) s_m = 0;
  if (lane == pp) s_pos[warp] = my_a;
  __syncthreads();
  sh_stamp(6);
  uint64_t* keys = reinterpret_cast<uint64_t*>(shm);       // [nw * KK]
  uint64_t* outk = keys + nw * KK;                          // [KK]
  uint32_t* heads = reinterpret_cast<uint32_t*>(outk + KK);  // [sh_max_lists(KK)]: the a of every list's smallest key
  uint32_t* cand = heads + sh_max_lists(KK);                 // [SH_CAND]
  ulonglong2* ekeys = reinterpret_cast<ulonglong2*>(cand + SH_CAND);      // [nw * k], then [k]
  {
    const uint32_t pv = lane < nw ? s_pos[lane] : 0xffffffffu;
    int rank = 0;
    for (int j = 0; j < nw; j++) {
      const uint32_t o = __shfl_sync(FULL_MASK, pv, j);
      rank += (o < pv || (o == pv && j < lane)) ? 1 : 0;
    }
    const unsigned bsel = __ballot_sync(FULL_MASK, lane < nw && rank == cw - 1);
    const uint32_t V = __shfl_sync(FULL_MASK, pv, __ffs(bsel) - 1);
    const uint32_t T = min(V, s_cta_thr);
    const bool ok = lane < KK && my_a != 0xffffffffu && my_a <= T;
    const unsigned bm = __ballot_sync(FULL_MASK, ok);
    int base = 0;
    if (lane == 0 && bm != 0u) base = atomicAdd(&s_m, __popc(bm));
    base = __shfl_sync(FULL_MASK, base, 0);
    if (ok) keys[base + __popc(bm & ((1u << lane) - 1u))] = ((uint64_t)my_a << 32) | (uint64_t)my_r;
  }
  if (threadIdx.x < KK) outk[threadIdx.x] = KEY_EMPTY64;
  __syncthreads();
  {
    const int m = s_m;
    for (int i = threadIdx.x; i < m; i += blockDim.x) {
      const uint64_t me = keys[i];
      int rank = 0;
#pragma unroll 4
      for (int j = 0; j < m; j++) rank += keys[j] < me ? 1 : 0;  // keys are distinct (row ids)
      if (rank < KK) outk[rank] = me;
    }
  }
  __syncthreads();
  sh_stamp(7);
  const int nl = gridDim.x;
  unsigned long long* all = pub + (size_t)qi * nl * KK;
  if (threadIdx.x < KK) __stcg(all + (size_t)blockIdx.x * KK + threadIdx.x, (unsigned long long)outk[threadIdx.x]);
  __threadfence();
  __syncthreads();
  sh_stamp(8);
  __shared__ unsigned int s_last;
  __shared__ int s_cnt, s_over;
  __shared__ unsigned int s_B;  // order-preserving bits of the k-th smallest list head's a (atomicMax from 0); 0xffffffff = fewer than k lists
  unsigned long long* ctrl = out.ctrl + 4 * (size_t)qi;
  if (threadIdx.x == 0) {
    const unsigned long long tk = atomicAdd(ctrl, 1ull);
    s_last = (tk == gridDim.x - 1) ? 1u : 0u;
    s_cnt = 0;
    s_over = 0;
    s_B = 0u;
  }
  __syncthreads();
  sh_stamp(9);
  if (!s_last) return;
  __threadfence();
  sh_stamp(10, true);
  // every published key is requested now; the bound is found meanwhile
  constexpr int PRE = (sh_max_lists(KK) * KK + SH_MIN_THREADS - 1) / SH_MIN_THREADS;
  uint64_t pre[PRE];
  const int total = nl * KK;
#pragma unroll
  for (int i = 0; i < PRE; i++) {
    const int idx = threadIdx.x + i * blockDim.x;
    pre[i] = idx < total ? (uint64_t)__ldcg(all + idx) : KEY_EMPTY64;
  }
  const int nl4 = (nl + 3) & ~3;
  for (int i = threadIdx.x; i < nl4; i += blockDim.x) heads[i] = i < nl ? (uint32_t)((uint64_t)__ldcg(all + (size_t)i * KK) >> 32) : 0xffffffffu;
  if (COSINE && threadIdx.x == 32) s_qq = ref_sum_thread<REF_DOT>(q, q, d, lanes);
  __syncthreads();
  sh_stamp(11, true);
  // B = the k-th smallest head (with multiplicity): k distinct rows lie at or below it; only its a matters, and it is
  // the largest head with fewer than k heads below it (none when there are fewer than k lists).  Heads are ranked on
  // their 32 order-preserving bits.  Many lists per warp (two CTAs per SM: 296 lists, 8 warps): every warp first keeps
  // the k smallest of its share (ranks by shuffles, ties by index) -- whatever is among the k smallest of all is among
  // the k smallest of its warp.  Then all-pairs counts, four heads per shared-memory load, S adjacent lanes sharing a
  // head's comparisons so that every thread of the CTA has work.
  {
    const uint32_t* hv = heads;
    int nh = nl;
    const int per = (nl + nw - 1) / nw;  // <= 64: nl <= 320 lists, nw >= 8 warps
    if (per > 2 * k) {
      uint32_t* lvl = cand;  // [nw * k] (the candidate list is written after the bound is known)
      const int h0 = warp * per, i1 = lane + 32;
      const uint32_t v0 = (lane < per && h0 + lane < nl) ? heads[h0 + lane] : 0xffffffffu;
      const uint32_t v1 = (i1 < per && h0 + i1 < nl) ? heads[h0 + i1] : 0xffffffffu;
      int r0 = 0, r1 = 0;
      for (int j = 0; j < per; j++) {
        const uint32_t o = __shfl_sync(FULL_MASK, j < 32 ? v0 : v1, j & 31);
        r0 += (o < v0 || (o == v0 && j < lane)) ? 1 : 0;
        r1 += (o < v1 || (o == v1 && j < i1)) ? 1 : 0;
      }
      if (lane < per && r0 < k) lvl[warp * k + r0] = v0;
      if (i1 < per && r1 < k) lvl[warp * k + r1] = v1;
      nh = nw * k;
      if (threadIdx.x < 4) lvl[nh + threadIdx.x] = 0xffffffffu;  // padding of the last quad
      hv = lvl;
      __syncthreads();
    }
    if (nh < k) {
      if (threadIdx.x == 0) s_B = 0xffffffffu;
    } else {
      const int S = nh > 64 ? 4 : 1;
      const int quads = (nh + 3) >> 2, qper = (quads + S - 1) / S;
      const uint4* h4 = reinterpret_cast<const uint4*>(hv);
      const int items = ((nh * S + 31) & ~31);  // whole warps (the reduction below shuffles)
      for (int it = threadIdx.x; it < items; it += blockDim.x) {
        const int i = it / S, part = it - i * S;
        const uint32_t h = i < nh ? hv[i] : 0u;
        int lt = 0;
        const int j1 = min(quads, (part + 1) * qper);
#pragma unroll 4
        for (int j = part * qper; j < j1; j++) {
          const uint4 v = h4[j];
          lt += (v.x < h ? 1 : 0) + (v.y < h ? 1 : 0) + (v.z < h ? 1 : 0) + (v.w < h ? 1 : 0);
        }
        for (int o = 1; o < S; o <<= 1) lt += __shfl_xor_sync(FULL_MASK, lt, o);
        if (i < nh && part == 0 && lt <= k - 1) atomicMax(&s_B, h);  // (an empty head, 0xffffffff, when fewer than k lists hold a key)
      }
    }
  }
  __syncthreads();
  sh_stamp(12, true);
  float tau = inf;
  if (s_B != 0xffffffffu) {
    tau = f32_next_up(__double2float_ru((double)sh_unord(s_B) + s_slack2 + 1e-37));
    if (!(tau == tau)) tau = inf;
  }
#pragma unroll
  for (int i = 0; i < PRE; i++) {
    const uint64_t key = pre[i];
    if (key == KEY_EMPTY64) continue;
    if (sh_unord((uint32_t)(key >> 32)) <= tau) {
      const int slot = atomicAdd(&s_cnt, 1);
      if (slot < SH_CAND) cand[slot] = (uint32_t)key;
      if ((threadIdx.x + i * blockDim.x) % KK == KK - 1) s_over = 1;  // a full list ends inside the band
    }
  }
  __syncthreads();
  sh_stamp(13, true);
  int cnt = s_cnt;
  if (threadIdx.x == 0) fb[qi] = (s_over || cnt > SH_CAND) ? 1 : 0;  // 1: the fallback scan overwrites what follows
  if (cnt > SH_CAND) cnt = SH_CAND;
  // exact scores of the candidates: one row per half-warp, lane hl is SIMD lane hl of the modelled JVM
  WarpTopKReg tk;
  tk.init(nullptr, 32, k, lane);
  {
    const double qq = COSINE ? s_qq : 0.0;
    const int hl = lane & 15, hw = lane >> 4;
    const unsigned hmask = (lane < 16) ? 0x0000ffffu : 0xffff0000u;
    const int base_lane = lane & 16;
    for (int i = warp * 2; i < cnt; i += nw * 2) {
      const bool have = i + hw < cnt;
      Key key = key_empty();
      bool ok = false;
      if (have) {  // uniform across the half-warp
        const int64_t row = (int64_t)cand[i + hw];
        ok = !(skip != nullptr && skip[row]);
        if (ok) {
          const float* x = X + (size_t)row * d;
          double score;
          if (COSINE) {
            const double dotx = ref_sum_halfwarp<REF_DOT>(q, x, d, lanes, hl, hmask, base_lane);
            const double xx = ref_sum_halfwarp<REF_DOT>(x, x, d, lanes, hl, hmask, base_lane);
            score = ref_cosine_from_sums(dotx, qq, xx);
          } else {
            score = -__dsqrt_rn(ref_sum_halfwarp<REF_L2SQ>(q, x, d, lanes, hl, hmask, base_lane));
          }
          key = Key{rank_hi_from_score(score), (uint64_t)row};
        }
      }
      tk.push(key, ok && hl == 0, lane);
    }
  }
  sh_stamp(14, true);
  ulonglong2* eout = ekeys + nw * k;
  if (threadIdx.x == 0) s_m = 0;
  if (threadIdx.x < k) st_key(eout + threadIdx.x, key_empty());
  __syncthreads();
  if (lane < k && tk.my_lo != KEY_EMPTY64) st_key(ekeys + atomicAdd(&s_m, 1), Key{tk.my_hi, tk.my_lo});
  __syncthreads();
  block_rank_select(ekeys, s_m, k, eout);
  __syncthreads();
  topk_write_out(eout, k, out, qi);
  sh_stamp(15, true);
}

// grid (P, nq): CTA (p, qi) owns the p-th slice of the row groups (gsz rows each) of query qi.
template <int TPR, int U, bool COSINE, class TK>
__global__ void __launch_bounds__(BQ_SELECT_THREADS)
batch_select_kernel(const float* __restrict__ X, int64_t n, int d, const float* __restrict__ Q,
                    const uint8_t* __restrict__ skip, const float* gm /* written by the predecessor: no __restrict__ */, int64_t gm_stride,
                    int64_t ngroups, int gsz, const SegStats* __restrict__ stats, int half, HalfOperands H, int k, int kp,
                    int cap, int32_t* __restrict__ fb, int nq_total, TopkOut out) {
  extern __shared__ __align__(128) ulonglong2 smem[];
  pdl_trigger();
  pdl_wait();  // group minima of the nomination kernel
  constexpr int L = TPR * 4;
  constexpr int RPB = (32 / TPR) * U;  // rows per scoring batch
  const int BPG = gsz / RPB;           // scoring batches per group (the host keeps gsz >= RPB)
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int nw = blockDim.x >> 5;
  const int qi = blockIdx.y;
  const int stride1 = kp + TOPK_BUF;
  const float* __restrict__ q = Q + (size_t)qi * d;
  const float* gmq = gm + (size_t)qi * gm_stride;

  // shared memory: [collectors (phase 1 and phase 3 alias)][group list: cap ints][query: d floats]
  size_t coll_bytes = (size_t)nw * topk_warp_smem(kp);
  const size_t b3 = topk_block_smem(k, kp, nw);
  if (b3 > coll_bytes) coll_bytes = b3;
  coll_bytes = (coll_bytes + 15) & ~size_t(15);
  int* list = reinterpret_cast<int*>(reinterpret_cast<unsigned char*>(smem) + coll_bytes);
  float* qs = reinterpret_cast<float*>(list + ((cap + 3) & ~3));

  __shared__ double s_qq;
  __shared__ float s_q2;
  __shared__ int s_cnt;
  bq_stamp(0);
  if (H.Xh != nullptr) {
    for (int i = threadIdx.x; i < H.dp; i += blockDim.x) qs[i] = __half2float(H.Qh[(size_t)qi * H.dp + i]);
  } else {
    for (int i = threadIdx.x; i < d; i += blockDim.x) qs[i] = q[i];
  }
  if (COSINE && threadIdx.x == 32) s_qq = ref_sum_thread_L<L, REF_DOT>(q, q, d);
  if (warp == 0) {
    float ss = 0.0f;
    for (int i = lane; i < d; i += 32) ss = fmaf(q[i], q[i], ss);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(FULL_MASK, ss, o);
    if (lane == 0) {
      s_q2 = ss;
      s_cnt = 0;
    }
  }
  __syncthreads();
  const float q2 = s_q2;
  if (!(q2 < 1e30f)) {  // non-finite (or absurdly large) query: exact scan instead
    if (blockIdx.x == 0 && threadIdx.x == 0) batch_list_fallback(fb, nq_total, qi);
    return;
  }
  const double qq = COSINE ? s_qq : 0.0;
  const float qn = COSINE ? (float)sqrt(qq) : 0.0f;
  const double slack = batch_slack(COSINE, half != 0, d, sqrt((double)__uint_as_float(stats->xmax2_bits)), sqrt((double)q2));

  bq_stamp(1);
  // phase 1: T = an upper bound of the k-th smallest group minimum of this slice
  const int64_t g0 = ngroups * blockIdx.x / gridDim.x, g1 = ngroups * (blockIdx.x + 1) / gridDim.x;
  float tau = __int_as_float(0x7f800000);
  if (k <= BQ_TMIN_MAX_K && k <= (int)blockDim.x) {
    // small k: the k-th smallest of the per-thread minima.  The minima belong to distinct groups, so k of
    // them lie at or below it -- a valid bound, and nearly the exact one because the few best groups
    // rarely share a thread.  One strided pass, one 256-way rank count.
    float* tmin = reinterpret_cast<float*>(smem);
    float tm = __int_as_float(0x7f800000);
    {
      const float4* gm4 = reinterpret_cast<const float4*>(gmq);
      const int64_t c0 = g0 >> 2, c1 = (g1 + 3) >> 2;
      const float4 inf4 = make_float4(tm, tm, tm, tm);
      for (int64_t cb = c0 + threadIdx.x; cb < c1; cb += (int64_t)BQ_SEL_MLP * blockDim.x) {
        float4 v[BQ_SEL_MLP];  // BQ_SEL_MLP independent loads in flight per thread: this pass is pure latency
#pragma unroll
        for (int j = 0; j < BQ_SEL_MLP; j++) {
          const int64_t c = cb + (int64_t)j * blockDim.x;
          v[j] = c < c1 ? gm4[c] : inf4;
        }
#pragma unroll
        for (int j = 0; j < BQ_SEL_MLP; j++) {
          const int64_t g = (cb + (int64_t)j * blockDim.x) << 2;
          if (g >= g0 && g + 3 < g1) {
            tm = fminf(fminf(tm, v[j].x), fminf(v[j].y, fminf(v[j].z, v[j].w)));
          } else {  // slice boundaries
            if (g >= g0 && g < g1) tm = fminf(tm, v[j].x);
            if (g + 1 >= g0 && g + 1 < g1) tm = fminf(tm, v[j].y);
            if (g + 2 >= g0 && g + 2 < g1) tm = fminf(tm, v[j].z);
            if (g + 3 >= g0 && g + 3 < g1) tm = fminf(tm, v[j].w);
          }
        }
      }
    }
    tmin[threadIdx.x] = tm;
    __shared__ float s_T;
    if (threadIdx.x == 0) s_T = __int_as_float(0x7f800000);
    __syncthreads();
    int rank = 0;
    for (int j = 0; j < (int)blockDim.x; j++) {
      const float o = tmin[j];
      rank += (o < tm) | ((o == tm) & (j < (int)threadIdx.x));
    }
    if (rank == k - 1) s_T = tm;
    __syncthreads();
    const float T = s_T;
    if (T < __int_as_float(0x7f800000)) {
      tau = f32_next_up(__double2float_ru((double)T + 2.0 * slack + 1e-37));
      if (!(tau == tau)) tau = __int_as_float(0x7f800000);
    }
    __syncthreads();  // tmin aliases the collectors
  } else {
    WarpTopK c1;
    c1.init(smem + (size_t)warp * stride1, kp, k, lane);
    for (int64_t gb = g0 + (int64_t)warp * 32; gb < g1; gb += (int64_t)nw * 32) {
      const int64_t g = gb + lane;
      const bool ok = g < g1;
      const float v = ok ? gmq[g] : 0.0f;
      c1.push(Key{f64_ordered((double)v), (uint64_t)g}, ok, lane);
    }
    c1.flush(lane);
    block_combine_lists(smem, stride1, nw, kp, warp, lane);
    const Key kth = ld_key(smem + (k - 1));
    __syncthreads();  // everyone has read the k-th key: the collectors may be reused
    if (!key_is_empty(kth)) {
      const double T = f64_from_ordered(kth.hi);
      tau = f32_next_up(__double2float_ru(T + 2.0 * slack + 1e-37));
      if (!(tau == tau)) tau = __int_as_float(0x7f800000);
    }
  }
  bq_stamp(2);
  // phase 2: groups that can hold a top-k row
  {
    const float4* gm4 = reinterpret_cast<const float4*>(gmq);
    const int64_t c0 = g0 >> 2, c1 = (g1 + 3) >> 2;
    const float inf = __int_as_float(0x7f800000);
    const float4 inf4 = make_float4(inf, inf, inf, inf);
    for (int64_t cb = c0 + threadIdx.x; cb < c1; cb += (int64_t)BQ_SEL_MLP * blockDim.x) {
      float4 v[BQ_SEL_MLP];
#pragma unroll
      for (int j = 0; j < BQ_SEL_MLP; j++) {
        const int64_t c = cb + (int64_t)j * blockDim.x;
        v[j] = c < c1 ? gm4[c] : inf4;
      }
#pragma unroll
      for (int j = 0; j < BQ_SEL_MLP; j++) {
        const float m4 = fminf(fminf(v[j].x, v[j].y), fminf(v[j].z, v[j].w));
        if (!(m4 <= tau)) continue;  // almost always: nothing of this chunk qualifies
        const float e[4] = {v[j].x, v[j].y, v[j].z, v[j].w};
#pragma unroll
        for (int i = 0; i < 4; i++) {
          const int64_t g = ((cb + (int64_t)j * blockDim.x) << 2) + i;
          if (g >= g0 && g < g1 && e[i] <= tau) {
            const int idx = atomicAdd(&s_cnt, 1);
            if (idx < cap) list[idx] = (int)g;
          }
        }
      }
    }
  }
  __syncthreads();
  int cnt = s_cnt;
  if (cnt > cap) {  // adversarial data (e.g. thousands of duplicates): full exact scan for this query
    if (threadIdx.x == 0) batch_list_fallback(fb, nq_total, qi);
    cnt = cap;      // keep the ticket protocol of the epilogue intact; the fallback overwrites the result
  }
  bq_stamp(3);
  // phase 3: exact scores of the candidate groups' rows, ranked like scan.cu.  All their lines are requested
  // into L2 first, so the scoring rounds below do not each wait on HBM.
  {
    // one bulk L2 prefetch per candidate group (its rows are contiguous): the fp16 rows the pre-filter reads, or
    // the fp32 rows when whole groups are scored exactly
    // (the fp16 pre-filter prefetches its rows itself, a few rounds ahead of its loads)
    const bool hp = H.Xh != nullptr;
    const char* xb = reinterpret_cast<const char*>(X);
    const int64_t rbytes = (int64_t)d * 4;
    const int64_t xbytes = n * rbytes, gbytes = (int64_t)gsz * rbytes;
    for (int i = threadIdx.x; i < (hp ? 0 : cnt); i += blockDim.x) {
      const int64_t off = (int64_t)list[i] * gbytes;
      const int64_t len = off + gbytes <= xbytes ? gbytes : xbytes - off;
      if (len > 0) asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(xb + off), "r"((uint32_t)len) : "memory");
    }
  }
  bq_stamp(4);
  TK tk;
  tk.init(smem + (size_t)warp * stride1, kp, k, lane);
  const int ub = d - (d % L);
  const int nv = ub / L;
  float fthr = __int_as_float(0x7f800000);
  uint64_t seen_hi = KEY_EMPTY64, seen_lo = KEY_EMPTY64;
  if (H.Xh != nullptr) {
    // fp16 pre-filter per group, exact scores for the rows that pass (qs holds the scaled fp16 query as floats)
    int* plist = reinterpret_cast<int*>(qs + ((H.dp + 3) & ~3)) + warp * BQ_PLIST;
    bq_groups_prefilter_exact<COSINE, TK>(H, X, n, d, q, qs, H.qinv[qi], skip, L, qq, tau, list, warp, nw, cnt, gsz, plist, tk, lane);
  } else {
  const int total = cnt * BPG;
  for (int b = warp; b < total; b += nw) {
    const int64_t row_base = (int64_t)list[b / BPG] * gsz + (int64_t)(b % BPG) * RPB;
    if (row_base >= n) continue;
    scan_batch_ldg<TPR, U, COSINE, TK>(X, n, d, q, qs, skip, row_base, nv, ub, qq, qn, fthr, tk, lane);
    if (tk.thr.hi != seen_hi || tk.thr.lo != seen_lo) {
      seen_hi = tk.thr.hi;
      seen_lo = tk.thr.lo;
      fthr = scan_filter_threshold<COSINE>(tk.thr);
    }
  }
  }
  bq_stamp(5);
  topk_epilogue(tk, smem, kp, k, out);
  bq_stamp(6);
}

// Large batches, k <= 32: one WARP per query.  Every step of the selection is latency-bound (a pass over the
// query's group minima, then a dozen 32 KB row groups fetched from HBM), so what matters is how many queries
// an SM works on at once: 16 here against 2 CTAs of the kernel above, with no block-wide barrier anywhere.
template <int TPR, int U, bool COSINE>
__global__ void __launch_bounds__(BQ_SELECT_THREADS)
batch_select_warp_kernel(const float* __restrict__ X, int64_t n, int d, const float* __restrict__ Q,
                         const uint8_t* __restrict__ skip, const float* gm /* written by the predecessor: no __restrict__ */, int64_t gm_stride,
                         int64_t ngroups, int gsz, const SegStats* __restrict__ stats, int half, HalfOperands H, int k, int cap,
                         int32_t* __restrict__ fb, int nq, int64_t* __restrict__ ids, double* __restrict__ scores,
                         int32_t* __restrict__ counts, int64_t id_base, int64_t out_stride) {
  extern __shared__ __align__(128) ulonglong2 smem[];
  pdl_trigger();
  pdl_wait();  // group minima of the nomination kernel
  constexpr int L = TPR * 4;
  constexpr int RPB = (32 / TPR) * U;
  const int BPG = gsz / RPB;
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int qi = blockIdx.x * (blockDim.x >> 5) + warp;
  if (qi >= nq) return;
  const int qpad = (((H.Xh != nullptr ? H.dp : d) + 3) & ~3) + BQ_PLIST;  // query floats + the pre-filter's survivor list
  int* list = reinterpret_cast<int*>(smem) + (size_t)warp * (cap + qpad);
  float* qs = reinterpret_cast<float*>(list + cap);
  int* plist = reinterpret_cast<int*>(qs + qpad - BQ_PLIST);
  const float* __restrict__ q = Q + (size_t)qi * d;
  const float* gmq = gm + (size_t)qi * gm_stride;
  const float inf = __int_as_float(0x7f800000);

  float ss = 0.0f;
  for (int i = lane; i < d; i += 32) {
    const float v = q[i];
    if (H.Xh == nullptr) qs[i] = v;
    ss = fmaf(v, v, ss);
  }
  if (H.Xh != nullptr)  // the pre-filter multiplies the fp16 operand copies: stage the scaled fp16 query as floats
    for (int i = lane; i < H.dp; i += 32) qs[i] = __half2float(H.Qh[(size_t)qi * H.dp + i]);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(FULL_MASK, ss, o);
  if (!(ss < 1e30f)) {  // non-finite query: exact scan instead
    if (lane == 0) batch_list_fallback(fb, nq, qi);
    return;
  }
  double qq = 0.0;
  if (COSINE) {
    if (lane == 0) qq = ref_sum_thread_L<L, REF_DOT>(q, q, d);
    qq = __shfl_sync(FULL_MASK, qq, 0);
  }
  const float qn = COSINE ? (float)sqrt(qq) : 0.0f;
  const double slack = batch_slack(COSINE, half != 0, d, sqrt((double)__uint_as_float(stats->xmax2_bits)), sqrt((double)ss));
  __syncwarp();

  // T = k-th smallest of the 32 per-lane minima (distinct groups, so k of them lie at or below it)
  const float4* gm4 = reinterpret_cast<const float4*>(gmq);
  const int64_t nch = (ngroups + 3) >> 2;  // the tile padding beyond ngroups holds +inf
  float tm = inf;
#pragma unroll 8
  for (int64_t c = lane; c < nch; c += 32) {
    const float4 v = gm4[c];
    tm = fminf(fminf(tm, v.x), fminf(v.y, fminf(v.z, v.w)));
  }
  int rank = 0;
#pragma unroll
  for (int j = 0; j < 32; j++) {
    const float o = __shfl_sync(FULL_MASK, tm, j);
    rank += (o < tm) | ((o == tm) & (j < lane));
  }
  const unsigned who = __ballot_sync(FULL_MASK, rank == k - 1);
  const float T = __shfl_sync(FULL_MASK, tm, __ffs(who) - 1);
  float tau = inf;
  if (T < inf) {
    tau = f32_next_up(__double2float_ru((double)T + 2.0 * slack + 1e-37));
    if (!(tau == tau)) tau = inf;
  }
  // groups that can hold a top-k row, compacted with ballots
  int cnt = 0;
  for (int64_t c0 = 0; c0 < nch; c0 += 32) {
    const int64_t c = c0 + lane;
    float4 v = make_float4(inf, inf, inf, inf);
    if (c < nch) v = gm4[c];
    const float e[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int j = 0; j < 4; j++) {
      const int64_t g = (c << 2) + j;
      const bool p = c < nch && g < ngroups && e[j] <= tau;
      const unsigned m = __ballot_sync(FULL_MASK, p);
      if (p) {
        const int pos = cnt + __popc(m & ((1u << lane) - 1u));
        if (pos < cap) list[pos] = (int)g;
      }
      cnt += __popc(m);
    }
  }
  if (cnt > cap) {  // adversarial data: full exact scan for this query (its result below is overwritten)
    if (lane == 0) batch_list_fallback(fb, nq, qi);
    cnt = cap;
  }
  __syncwarp();
  {
    const bool hp = H.Xh != nullptr;
    const char* xb = reinterpret_cast<const char*>(X);
    const int64_t rbytes = (int64_t)d * 4;
    const int64_t xbytes = n * rbytes, gbytes = (int64_t)gsz * rbytes;
    for (int i = lane; i < (hp ? 0 : cnt); i += 32) {
      const int64_t off = (int64_t)list[i] * gbytes;
      const int64_t len = off + gbytes <= xbytes ? gbytes : xbytes - off;
      if (len > 0) asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(xb + off), "r"((uint32_t)len) : "memory");
    }
  }
  WarpTopKReg tk;
  tk.init(nullptr, 32, k, lane);
  const int ub = d - (d % L);
  const int nv = ub / L;
  float fthr = inf;
  uint64_t seen_hi = KEY_EMPTY64, seen_lo = KEY_EMPTY64;
  if (H.Xh != nullptr) {
    bq_groups_prefilter_exact<COSINE, WarpTopKReg>(H, X, n, d, q, qs, H.qinv[qi], skip, L, qq, tau, list, 0, 1, cnt, gsz, plist, tk, lane);
  } else {
  const int total = cnt * BPG;
  for (int b = 0; b < total; b++) {
    const int64_t row_base = (int64_t)list[b / BPG] * gsz + (int64_t)(b % BPG) * RPB;
    if (row_base >= n) continue;
    scan_batch_ldg<TPR, U, COSINE, WarpTopKReg>(X, n, d, q, qs, skip, row_base, nv, ub, qq, qn, fthr, tk, lane);
    if (tk.thr.hi != seen_hi || tk.thr.lo != seen_lo) {
      seen_hi = tk.thr.hi;
      seen_lo = tk.thr.lo;
      fthr = scan_filter_threshold<COSINE>(tk.thr);
    }
  }
  }
  // lane i holds the i-th best key
  const bool ok = lane < k && tk.my_lo != KEY_EMPTY64;
  if (lane < k) {
    ids[(size_t)qi * out_stride + lane] = ok ? id_base + (int64_t)tk.my_lo : -1;
    scores[(size_t)qi * out_stride + lane] = ok ? score_from_rank_hi(tk.my_hi) : __longlong_as_double(0x7ff8000000000000ll);
  }
  const unsigned okm = __ballot_sync(FULL_MASK, ok);
  if (lane == 0) counts[qi] = __popc(okm);
}

// ---- (4) exact scan of the queries listed in fb (count, then query indices) ----------------------------------------
template <int TPR, int U, bool COSINE, class TK>
__global__ void __launch_bounds__(SCAN_THREADS)
batch_fallback_kernel(const float* __restrict__ X, int64_t n, int d, const float* __restrict__ Q,
                      const uint8_t* __restrict__ skip, int k, int kp, const int32_t* __restrict__ fb, int direct, TopkOut out) {
  extern __shared__ __align__(128) ulonglong2 smem[];
  constexpr int L = TPR * 4;
  constexpr int G = 32 / TPR;
  // direct (after scan_half_kernel): fb[y] says whether query y needs the scan, grid.y = queries; else fb holds a
  // count and a list of queries, walked by grid.y slots
  const int count = direct ? (fb[blockIdx.y] != 0 ? (int)blockIdx.y + 1 : 0) : fb[0];
  if ((int)blockIdx.y >= count) return;
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int nw = blockDim.x >> 5;
  const int stride_keys = kp + TOPK_BUF;
  float* qs = reinterpret_cast<float*>(reinterpret_cast<unsigned char*>(smem) + ((topk_block_smem(k, kp, nw) + 15) & ~size_t(15)));
  __shared__ double s_qq;
  const int ub = d - (d % L);
  const int nv = ub / L;
  const int64_t rows_per_batch = (int64_t)G * U;
  const int64_t nbatches = (n + rows_per_batch - 1) / rows_per_batch;
  const int64_t total_warps = (int64_t)gridDim.x * nw;
  for (int slot = blockIdx.y; slot < count; slot += gridDim.y) {
    const int qi = direct ? slot : fb[1 + slot];
    const float* __restrict__ q = Q + (size_t)qi * d;
    __syncthreads();  // previous query's epilogue is done with shared memory
    TK tk;
    tk.init(smem + (size_t)warp * stride_keys, kp, k, lane);
    for (int i = threadIdx.x; i < d; i += blockDim.x) qs[i] = q[i];
    if (COSINE && threadIdx.x == 0) s_qq = ref_sum_thread_L<L, REF_DOT>(q, q, d);
    __syncthreads();
    const double qq = COSINE ? s_qq : 0.0;
    const float qn = COSINE ? (float)sqrt(qq) : 0.0f;
    float fthr = __int_as_float(0x7f800000);
    uint64_t seen_hi = KEY_EMPTY64, seen_lo = KEY_EMPTY64;
    for (int64_t b = (int64_t)blockIdx.x * nw + warp; b < nbatches; b += total_warps) {
      scan_batch_ldg<TPR, U, COSINE, TK>(X, n, d, q, qs, skip, b * rows_per_batch, nv, ub, qq, qn, fthr, tk, lane);
      if (tk.thr.hi != seen_hi || tk.thr.lo != seen_lo) {
        seen_hi = tk.thr.hi;
        seen_lo = tk.thr.lo;
        fthr = scan_filter_threshold<COSINE>(tk.thr);
      }
    }
    topk_epilogue(tk, smem, kp, k, out, qi);
  }
}

// ---- host ---------------------------------------------------------------------------------------------------------
// tensor map over the segment's rows: the fp32 rows themselves (half = false) or their fp16 copy with pitch dp
bool batch_encode_segment_map(void* tm128, const void* X, int64_t n, int d, int64_t pitch, bool half, int box_rows) {
  static_assert(sizeof(CUtensorMap) == 128, "tensor map size");
  return encode_rows_map(reinterpret_cast<CUtensorMap*>(tm128), X, n, d, pitch, half, box_rows);
}

// host copy of the device-side bound, for diagnostics
double batch_slack_host(bool cosine, bool half, int d, double xmax, double qn) { return batch_slack(cosine, half, d, xmax, qn); }

bool batch_supported(int d, int lanes, bool cosine, int64_t n) {
  if (!scan_is_streaming(d, lanes, cosine)) return false;  // the re-score uses the streaming row scorer
  if (d < 32 || d > 65536) return false;
  if (n < 1 || n > (int64_t(1) << 31) - 2 * BQ_N) return false;  // 32-bit TMA coordinates, int group ids
  return encode_fn() != nullptr;
}

// The attribute is per kernel, not per launch shape: plans are cached, so every kernel is opted in to the
// largest dynamic shared memory any plan may ask for (a smaller, later request must not lower it).
template <typename K>
static cudaError_t set_smem_attr(K kern, size_t smem_max) {
  return cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_max);
}
constexpr size_t BQ_SELECT_SMEM_MAX = 220 * 1024, BQ_SELW_SMEM_MAX = 96 * 1024;

typedef void (*SelectKern)(const float*, int64_t, int, const float*, const uint8_t*, const float*, int64_t, int64_t, int,
                           const SegStats*, int, HalfOperands, int, int, int, int32_t*, int, TopkOut);
typedef void (*SelectWarpKern)(const float*, int64_t, int, const float*, const uint8_t*, const float*, int64_t, int64_t, int,
                               const SegStats*, int, HalfOperands, int, int, int32_t*, int, int64_t*, double*, int32_t*, int64_t, int64_t);
static SelectWarpKern pick_select_warp(int TPR, bool cosine) {
  if (cosine) return TPR == 4 ? batch_select_warp_kernel<4, 1, true> : (TPR == 2 ? batch_select_warp_kernel<2, 1, true> : batch_select_warp_kernel<1, 1, true>);
  return TPR == 4 ? batch_select_warp_kernel<4, 2, false> : (TPR == 2 ? batch_select_warp_kernel<2, 2, false> : batch_select_warp_kernel<1, 2, false>);
}
typedef void (*FallbackKern)(const float*, int64_t, int, const float*, const uint8_t*, int, int, const int32_t*, int, TopkOut);

template <class TK>
static SelectKern select_kernel(int TPR, bool cosine) {
  if (cosine) return TPR == 4 ? batch_select_kernel<4, 1, true, TK> : (TPR == 2 ? batch_select_kernel<2, 1, true, TK> : batch_select_kernel<1, 1, true, TK>);
  return TPR == 4 ? batch_select_kernel<4, 2, false, TK> : (TPR == 2 ? batch_select_kernel<2, 2, false, TK> : batch_select_kernel<1, 2, false, TK>);
}
template <class TK>
static FallbackKern fallback_kernel(int TPR, bool cosine) {
  if (cosine) return TPR == 4 ? batch_fallback_kernel<4, 1, true, TK> : (TPR == 2 ? batch_fallback_kernel<2, 1, true, TK> : batch_fallback_kernel<1, 1, true, TK>);
  return TPR == 4 ? batch_fallback_kernel<4, 2, false, TK> : (TPR == 2 ? batch_fallback_kernel<2, 2, false, TK> : batch_fallback_kernel<1, 2, false, TK>);
}
static SelectKern pick_select(int TPR, bool cosine, int k) {
  return k <= TOPK_REG_MAX_K ? select_kernel<WarpTopKReg>(TPR, cosine) : select_kernel<WarpTopK>(TPR, cosine);
}
static FallbackKern pick_fallback(int TPR, bool cosine, int k) {
  return k <= TOPK_REG_MAX_K ? fallback_kernel<WarpTopKReg>(TPR, cosine) : fallback_kernel<WarpTopK>(TPR, cosine);
}

typedef void (*ScanHalfKern)(const __half*, int64_t, int, const float*, float, const SegStats*, const float*, int,
                             const float*, const uint8_t*, int, int, int, int, int32_t*, int, unsigned long long*, TopkOut);
static ScanHalfKern pick_scan_half(bool cosine, int kk, int cpl) {
#define VS_SH(C, K) (cpl == 2 ? scan_half_kernel<C, K, 2> : (cpl == 4 ? scan_half_kernel<C, K, 4> : scan_half_kernel<C, K, 0>))
  if (cosine) return kk == 16 ? VS_SH(true, 16) : VS_SH(true, 32);
  return kk == 16 ? VS_SH(false, 16) : VS_SH(false, 32);
#undef VS_SH
}

typedef void (*GemmKern)(const CUtensorMap, const CUtensorMap, const float*, const float*, int64_t, int, int, int, int64_t,
                         float*, int64_t);
template <int GROUP>
static GemmKern pick_gemm_g(bool stat, bool half, bool cosine) {
  if (stat) {
    if (half) return cosine ? batch_gemm_kernel<true, true, true, GROUP> : batch_gemm_kernel<true, true, false, GROUP>;
    return cosine ? batch_gemm_kernel<true, false, true, GROUP> : batch_gemm_kernel<true, false, false, GROUP>;
  }
  if (half) return cosine ? batch_gemm_kernel<false, true, true, GROUP> : batch_gemm_kernel<false, true, false, GROUP>;
  return cosine ? batch_gemm_kernel<false, false, true, GROUP> : batch_gemm_kernel<false, false, false, GROUP>;
}
static GemmKern pick_gemm(bool stat, bool half, bool cosine, int group) {
  return group == 16 ? pick_gemm_g<16>(stat, half, cosine) : (group == 32 ? pick_gemm_g<32>(stat, half, cosine) : pick_gemm_g<64>(stat, half, cosine));
}

template <int GROUP>
static GemmKern pick_gemm_pair_g(bool stat, bool half, bool cosine) {
  if (stat) {
    if (half) return cosine ? batch_gemm_pair_kernel<true, true, true, GROUP> : batch_gemm_pair_kernel<true, true, false, GROUP>;
    return cosine ? batch_gemm_pair_kernel<true, false, true, GROUP> : batch_gemm_pair_kernel<true, false, false, GROUP>;
  }
  if (half) return cosine ? batch_gemm_pair_kernel<false, true, true, GROUP> : batch_gemm_pair_kernel<false, true, false, GROUP>;
  return cosine ? batch_gemm_pair_kernel<false, false, true, GROUP> : batch_gemm_pair_kernel<false, false, false, GROUP>;
}
static GemmKern pick_gemm_pair(bool stat, bool half, bool cosine, int group) {
  return group == 16 ? pick_gemm_pair_g<16>(stat, half, cosine)
                     : (group == 32 ? pick_gemm_pair_g<32>(stat, half, cosine) : pick_gemm_pair_g<64>(stat, half, cosine));
}

// dynamic shared memory of batch_select_kernel with nw warps: collectors, group list, query
static size_t batch_select_smem(const BatchLaunch& L, int nw) {
  size_t coll = (size_t)nw * topk_warp_smem(L.kp);
  const size_t b3 = topk_block_smem(L.k, L.kp, nw);
  if (b3 > coll) coll = b3;
  coll = (coll + 15) & ~size_t(15);
  // collectors, group list, query (d floats, or the dp floats of the scaled fp16 query), survivor lists of the pre-filter
  return coll + (size_t)((L.cap + 3) & ~3) * 4 + (size_t)((L.dp + 3) & ~3) * 4 + (size_t)nw * BQ_PLIST * 4;
}

bool batch_configure(BatchLaunch& L, int sms) {
  L.kp = topk_pad(L.k);
  L.tiles = (L.n + BQ_N - 1) / BQ_N;
  // Rows per nomination group: small groups cost group-minima traffic (2 * 4 n / G bytes per query, written then
  // read), large ones cost re-score traffic (about 1.3 k + 3 candidate groups of G * d * 4 bytes per query).
  {
    const double kc = 1.3 * L.k + 3.0;
    // (measured: the minima cost about 1.5x their bytes -- short scattered stores in the epilogue)
    const double gopt = 1.5 * sqrt(2.0 * (double)L.n / (kc * (double)L.d));
    int g = gopt < 23.0 ? 16 : (gopt < 46.0 ? 32 : 64);
    const int rpb = (32 / (L.lanes / 4)) * (L.cosine ? 1 : 2);  // rows per scoring batch of the re-score
    if (g < rpb) g = rpb;
    if (L.group_override == 16 || L.group_override == 32 || L.group_override == 64) g = L.group_override < rpb ? rpb : L.group_override;
    L.group = g;
  }
  L.ngroups = (L.n + L.group - 1) / L.group;
  L.gm_stride = L.tiles * (BQ_N / L.group);
  // resident query block when it leaves room for at least 3 row-block stages, else both operands stream
  {
    const int kbe = L.half ? 64 : 32;
    const int num_kb = (L.d + kbe - 1) / kbe;
    const size_t fixed = 1024 + BQ_AB_BYTES + 256;
    const size_t a_res = (size_t)num_kb * BQ_A_BYTES;
    L.gemm_stat = fixed + a_res + 3 * (size_t)BQ_B_BYTES <= BQ_GEMM_SMEM_BUDGET;
    const size_t per = L.gemm_stat ? BQ_B_BYTES : BQ_STAGE_BYTES;
    int ns = (int)((BQ_GEMM_SMEM_BUDGET - fixed - (L.gemm_stat ? a_res : 0)) / per);
    if (ns > BQ_MAX_STAGES) ns = BQ_MAX_STAGES;
    L.gemm_stages = ns;
    L.gemm_smem = fixed + (L.gemm_stat ? a_res : 0) + (size_t)ns * per;
    // CTA pairs (batches > 128 queries): resident query block + at least 4 stages of half a row tile, else both stream
    L.pair_stat = fixed + a_res + 4 * (size_t)BQ_BH_BYTES + 128 <= BQ_GEMM_SMEM_BUDGET;
    {
      const size_t per2 = L.pair_stat ? BQ_BH_BYTES : BQ_A_BYTES + BQ_BH_BYTES;
      int ns2 = (int)((BQ_GEMM_SMEM_BUDGET - fixed - (L.pair_stat ? a_res : 0) - 128) / per2);
      if (ns2 > BQ_PAIR_MAX_STAGES) ns2 = BQ_PAIR_MAX_STAGES;
      L.pair_stages = ns2;
      L.pair_smem = fixed + (L.pair_stat ? a_res : 0) + 128 + (size_t)ns2 * per2;
    }
  }
  // one or two queries: the CUDA-core scan of the fp16 copy (scan_half_kernel)
  L.sh_ok = false;
  if (L.half && L.k <= TOPK_REG_MAX_K && L.n < (int64_t(1) << 32)) {
    const int kk = L.k <= 16 ? 16 : 32;
    // Shapes tried in order.  (1) TWO 256-thread CTAs per SM (short lists only: the last CTA takes twice the lists): a
    // CTA's set-up (query scaling, ~2.5 us) and epilogue (compaction, publish, ticket, ~3 us) stream nothing, and with
    // one CTA per SM that is dead time of the SM in a stream of queries -- with two resident CTAs the neighbour (of this
    // query or, on another stream, of the next) keeps the SM's loads in flight meanwhile.  (2) one 512-thread CTA.
    // (3) one 256-thread CTA with larger tiles (long vectors: sixteen rings do not fit).
    struct Shape { int threads, per_sm; size_t tile_bytes, budget; };
    const Shape shapes[3] = {{256, 2, 4096, 112 * 1024}, {512, 1, 4096, 216 * 1024}, {256, 1, 8192, 216 * 1024}};
    for (int si = 0; si < 3 && !L.sh_ok; si++) {
      const Shape& S = shapes[si];
      if (S.per_sm == 2 && (kk != 16 || L.sh_override == 1)) continue;
      const int threads = S.threads, nw = threads / 32;
      int TR = (int)(S.tile_bytes / ((size_t)L.dp * 2)) & ~7;
      if (TR < 8) TR = 8;
      const size_t stage = ((size_t)TR * L.dp * 2 + (size_t)TR * 4 + 127) & ~size_t(127);
      const size_t fixed = (size_t)((L.dp + 3) & ~3) * 4 + (size_t)nw * 4 * 8 + 256;
      if (fixed + (S.per_sm == 2 ? 3 : 2) * (size_t)nw * stage > S.budget) continue;
      int NS = (int)((S.budget - fixed) / ((size_t)nw * stage));  // bytes in flight per SM are what the stream rate hangs on
      if (NS > 4) NS = 4;
      if (NS < 2) NS = 2;
      const size_t epi = (size_t)(nw * kk + kk) * 8 + (size_t)sh_max_lists(kk) * 4 + (size_t)SH_CAND * 4 + (size_t)(nw * L.k + L.k) * 16 + 64;
      if ((size_t)nw * NS * stage < epi) continue;
      L.sh_ok = true;
      L.sh_threads = threads;
      L.sh_per_sm = S.per_sm;
      L.sh_TR = TR;
      L.sh_NS = NS;
      L.sh_kk = kk;
      L.sh_cpl = L.dp <= 128 ? 2 : (L.dp <= 256 ? 4 : 0);
      L.sh_smem = (size_t)nw * NS * stage + fixed;
      const int64_t ntiles = (L.n + TR - 1) / TR;
      int64_t grid = (int64_t)sms * S.per_sm;
      if (grid > sh_max_lists(kk)) grid = sh_max_lists(kk);
      if (grid > (ntiles + nw - 1) / nw) grid = (ntiles + nw - 1) / nw;
      L.sh_grid = (int)(grid < 1 ? 1 : grid);
      if (set_smem_attr(pick_scan_half(L.cosine, kk, L.sh_cpl), 224 * 1024) != cudaSuccess) return false;
    }
  }
  // candidate groups per select CTA: the k-th smallest group minimum admits about k groups, the slack a few more
  L.cap = 4 * L.k + 256;
  for (;;) {
    L.select_smem = batch_select_smem(L, BQ_SELECT_THREADS / 32);
    if (L.select_smem <= 200 * 1024 || L.cap <= L.k + 64) break;
    L.cap = L.cap / 2 > L.k + 64 ? L.cap / 2 : L.k + 64;
  }
  if (L.select_smem > BQ_SELECT_SMEM_MAX) return false;
  L.selw_smem = (size_t)(BQ_SELECT_THREADS / 32) * (size_t)(L.cap + ((L.dp + 3) & ~3) + BQ_PLIST) * 4;
  if (L.k <= TOPK_REG_MAX_K && L.selw_smem <= BQ_SELW_SMEM_MAX) {
    if (set_smem_attr(pick_select_warp(L.lanes / 4, L.cosine), BQ_SELW_SMEM_MAX) != cudaSuccess) return false;
  } else {
    L.selw_smem = 0;  // warp-per-query selection not available for this shape
  }
  L.fb_threads = SCAN_THREADS;
  L.fb_smem = ((topk_block_smem(L.k, L.kp, L.fb_threads / 32) + 15) & ~size_t(15)) + (((size_t)L.d * 4 + 15) & ~size_t(15));
  if (L.fb_smem > BQ_SELECT_SMEM_MAX) return false;
  L.fb_gx = sms > TOPK_MAX_LISTS ? TOPK_MAX_LISTS : sms;
  // After scan_half_kernel the fallback check runs on every query and almost never has work: 64-thread CTAs (160
  // registers a thread: 10 K registers a CTA) fit beside the two resident scan CTAs of a neighbouring stream's query.
  // 256-thread CTAs (41 K registers) had to wait for an SM with at most one scan CTA, and the queries of a stream
  // queued up behind their own fallback checks (47 instead of 42 us per query with three queries in flight).
  L.fbd_threads = 64;
  L.fbd_gx = 2 * sms > TOPK_MAX_LISTS ? TOPK_MAX_LISTS : 2 * sms;
  L.fbd_smem = ((topk_block_smem(L.k, L.kp, L.fbd_threads / 32) + 15) & ~size_t(15)) + (((size_t)L.d * 4 + 15) & ~size_t(15));
  L.sms = sms;
  const int TPR = L.lanes / 4;
  {
    cudaError_t e;
    e = set_smem_attr(pick_gemm(L.gemm_stat, L.half, L.cosine, L.group), BQ_GEMM_SMEM_BUDGET);
    if (e != cudaSuccess) return false;
    if (L.pair_stages > 0 && set_smem_attr(pick_gemm_pair(L.pair_stat, L.half, L.cosine, L.group), BQ_GEMM_SMEM_BUDGET) != cudaSuccess) return false;
  }
  if (set_smem_attr(pick_select(TPR, L.cosine, L.k), BQ_SELECT_SMEM_MAX) != cudaSuccess) return false;
  if (set_smem_attr(pick_fallback(TPR, L.cosine, L.k), BQ_SELECT_SMEM_MAX) != cudaSuccess) return false;
  return true;
}

int batch_select_ctas(const BatchLaunch& L, int nq) {
  // enough select CTAs to fill the GPU when the batch is small; slices no shorter than 2048 groups
  // (every CTA nominates about k groups of its slice whatever the slice's length, so the re-score work grows with
  // P: rounded down -- 256 queries get one CTA each on 148 SMs, not two)
  int64_t P = 2 * (int64_t)L.sms / nq;
  if (L.select_ctas_override > 0) P = L.select_ctas_override;
  const int64_t by_len = (L.ngroups + 2047) / 2048;
  if (P > by_len) P = by_len;
  if (P > 64) P = 64;
  return (int)(P < 1 ? 1 : P);
}

int64_t batch_partial_keys(const BatchLaunch& L, int nq) {
  const int P = batch_select_ctas(L, nq);
  const int lists = L.fb_gx > L.fbd_gx ? L.fb_gx : L.fbd_gx;
  return (int64_t)(P > lists ? P : lists) * L.k;
}

// One chunk of nq queries (q, outputs and scratch already offset to the chunk).
cudaError_t launch_batch(const BatchLaunch& L, cudaStream_t st) {
  cudaError_t e;
  CUtensorMap tmQ;
  const int nqb = (L.nq + BQ_M - 1) / BQ_M;
  const int kbe = L.half ? 64 : 32;
  const int num_kb = (L.d + kbe - 1) / kbe;
  if (L.direct) {  // one or two queries: no tensor-core tile to amortise -- CUDA-core scan of the fp16 copy
    // (no memset of fb here: the scan writes the fallback flag of every query of the launch)
    count_launch();
    TopkOut o{L.partial, L.ctrl, L.partial_keys, L.ids_out, L.scores_out, L.counts_out, L.id_base, 0,
              L.out_stride > 0 ? L.out_stride : L.k};
    int grid = L.sh_grid - (L.reserve_sms > 0 ? L.reserve_sms * L.sh_per_sm : 0);
    if (grid < 1) grid = 1;
    e = launch_pdl(pick_scan_half(L.cosine, L.sh_kk, L.sh_cpl), dim3(grid, L.nq), dim3(L.sh_threads), L.sh_smem, st,
                   static_cast<const __half*>(L.xh), L.n, L.dp, L.coef, L.x_scale, L.stats, L.X, L.d,
                   L.q, L.skip, L.lanes, L.k, L.sh_TR, L.sh_NS, L.fb, L.nq, reinterpret_cast<unsigned long long*>(L.gm), o);
    if (e != cudaSuccess) return e;
    count_launch();
    const int gy = L.nq < BQ_FB_SLOTS ? L.nq : BQ_FB_SLOTS;
    pick_fallback(L.lanes / 4, L.cosine, L.k)<<<dim3(L.fbd_gx, gy), L.fbd_threads, L.fbd_smem, st>>>(L.X, L.n, L.d, L.q, L.skip, L.k, L.kp, L.fb, 1, o);
    return cudaGetLastError();
  }
  if ((e = cudaMemsetAsync(L.fb, 0, sizeof(int32_t) * (1 + 2 * (size_t)L.nq), st)) != cudaSuccess) return e;
  if (L.half) {
    count_launch();
    const int nq_pad = (L.nq + 2 * BQ_M - 1) / (2 * BQ_M) * (2 * BQ_M);  // whole pairs of query blocks
    query_convert_kernel<<<(nq_pad + 7) / 8, 256, 0, st>>>(L.q, L.nq, nq_pad, L.d, L.dp, L.x_scale, static_cast<__half*>(L.qh), L.qinv);
    if ((e = cudaGetLastError()) != cudaSuccess) return e;
    if (!encode_rows_map(&tmQ, L.qh, L.nq, L.d, L.dp, true, BQ_M)) return cudaErrorInvalidValue;
  } else {
    if (!encode_rows_map(&tmQ, L.q, L.nq, L.d, L.d, false, BQ_M)) return cudaErrorInvalidValue;
  }
  count_launch();
  if (L.pair_stages > 0 && L.pairs && L.nq > BQ_M && L.tmX128 != nullptr && L.sms >= 2) {
    const CUtensorMap tmX = *reinterpret_cast<const CUtensorMap*>(L.tmX128);
    const int nqp = (nqb + 1) / 2;
    int nsplit = (L.sms / 2) / nqp;
    if (nsplit < 1) nsplit = 1;
    if (nsplit > L.tiles) nsplit = (int)L.tiles;
    e = launch_pdl(pick_gemm_pair(L.pair_stat, L.half, L.cosine, L.group), dim3(2 * nqp * nsplit), dim3(BQ_THREADS), L.pair_smem, st,
                   tmQ, tmX, L.coef, L.qinv, L.n, num_kb, L.pair_stages, nqp, L.tiles, L.gm, L.gm_stride);
    if (e != cudaSuccess) return e;
  } else {
    const CUtensorMap tmX = *reinterpret_cast<const CUtensorMap*>(L.tmX);
    int nsplit = L.sms / nqb;
    if (nsplit < 1) nsplit = 1;
    if (nsplit > L.tiles) nsplit = (int)L.tiles;
    e = launch_pdl(pick_gemm(L.gemm_stat, L.half, L.cosine, L.group), dim3(nqb * nsplit), dim3(BQ_THREADS), L.gemm_smem, st,
                   tmQ, tmX, L.coef, L.qinv, L.n, num_kb, L.gemm_stages, nqb, L.tiles, L.gm, L.gm_stride);
    if (e != cudaSuccess) return e;
  }
  if ((e = cudaGetLastError()) != cudaSuccess) return e;
  if (L.gemm_only) return cudaSuccess;  // diagnostics: group minima only
  const int TPR = L.lanes / 4;
  const int P = batch_select_ctas(L, L.nq);
  TopkOut o{L.partial, L.ctrl, L.partial_keys, L.ids_out, L.scores_out, L.counts_out, L.id_base, 0,
            L.out_stride > 0 ? L.out_stride : L.k};
  count_launch();
  // Selection is latency-bound per query, so the batch decides how many warps a query gets: about 16 resident
  // warps per SM in total -- 8 per query for small batches (and P CTAs per query), down to one warp per query.
  int W = 8;
  while (W > 1 && (int64_t)L.nq * W > 16 * (int64_t)L.sms) W >>= 1;
  const bool force_warp = L.warp_min_q > 0 && L.nq >= L.warp_min_q;
  HalfOperands H{};
  if (L.half && L.prefilter) H = HalfOperands{static_cast<const __half*>(L.xh), static_cast<const __half*>(L.qh), L.qinv, L.coef, L.dp};
  // (the one-warp-per-query kernel measured slower than two-warp CTAs of the kernel above at every batch size once
  // the pre-filter batched its survivors -- 800 vs 705 us at 2048 queries, 1M x 128 -- so it only runs when asked for)
  if (L.selw_smem != 0 && force_warp) {
    const int wpb = BQ_SELECT_THREADS / 32;
    e = launch_pdl(pick_select_warp(TPR, L.cosine), dim3((L.nq + wpb - 1) / wpb), dim3(BQ_SELECT_THREADS), L.selw_smem, st,
                   L.X, L.n, L.d, L.q, L.skip, L.gm, L.gm_stride, L.ngroups, L.group, L.stats, L.half ? 1 : 0, H, L.k, L.cap, L.fb,
                   L.nq, L.ids_out, L.scores_out, L.counts_out, L.id_base, L.out_stride > 0 ? L.out_stride : (int64_t)L.k);
    if (e != cudaSuccess) return e;
  } else {
    if (W < 2) W = 2;
    e = launch_pdl(pick_select(TPR, L.cosine, L.k), dim3(P, L.nq), dim3(32 * W), batch_select_smem(L, W), st,
                   L.X, L.n, L.d, L.q, L.skip, L.gm, L.gm_stride, L.ngroups, L.group, L.stats, L.half ? 1 : 0, H, L.k, L.kp, L.cap,
                   L.fb, L.nq, o);
    if (e != cudaSuccess) return e;
  }
  if ((e = cudaGetLastError()) != cudaSuccess) return e;
  count_launch();
  const int gy = L.nq < BQ_FB_SLOTS ? L.nq : BQ_FB_SLOTS;
  // (launched normally: this kernel re-reads, through plain loads, what the selection kernel wrote)
  pick_fallback(TPR, L.cosine, L.k)<<<dim3(L.fb_gx, gy), L.fb_threads, L.fb_smem, st>>>(L.X, L.n, L.d, L.q, L.skip, L.k, L.kp, L.fb, 0, o);
  return cudaGetLastError();
}

}  // namespace vs
