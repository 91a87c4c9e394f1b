// batch.cu -- K2: exact brute-force top-k for a BATCH of queries over a resident segment.
//
// Same contract as scan.cu (searchBruteForceSegment, J/fdb/FdbVectorIndex.java:676-721, once per
// query): every returned id and score is produced by the reference arithmetic of scan_rows.cuh.
// What changes is how rows are NOMINATED.  One query streams the segment from HBM (scan.cu); nq
// queries would stream it nq times, so here the segment is read once per 128 queries and the
// dense part -- the nq x n inner products -- runs on the 5th-generation tensor cores:
//
//   (1) row_prep_kernel   one pass at first use: per row (alpha, beta) so that
//                         a(q, x) = alpha + beta * <q, x>  orders rows like the metric does
//                         (L2: |x|^2 - 2<q,x>;  cosine: -<q,x>/|x|), plus max |x|^2.
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
#include <cuda.h>  // CUtensorMap and enums only; the encoder is resolved through the runtime (no libcuda link)

#include <mutex>

#include "kernels.h"
#include "scan_rows.cuh"

namespace vs {

constexpr int BQ_M = 128;       // queries per tile = TMEM lanes
constexpr int BQ_N = 256;       // rows per tile = TMEM columns of one accumulator
constexpr int BQ_KB = 32;       // floats per k-block: one 128-byte swizzle atom
constexpr int BQ_STAGES = 4;
constexpr int BQ_THREADS = 256;
constexpr uint32_t BQ_A_BYTES = BQ_M * 128;
constexpr uint32_t BQ_B_BYTES = BQ_N * 128;
constexpr uint32_t BQ_STAGE_BYTES = BQ_A_BYTES + BQ_B_BYTES;
constexpr uint32_t BQ_AB_BYTES = 4 * 2 * BQ_N * 8;  // (alpha, beta) of the tile's rows: per epilogue warp, per accumulator
constexpr size_t BQ_GEMM_SMEM = 1024 + (size_t)BQ_STAGES * BQ_STAGE_BYTES + BQ_AB_BYTES + 256;
constexpr int BQ_SELECT_THREADS = 256;
constexpr int BQ_FB_SLOTS = 8;  // grid.y of the fallback scan

// ---- PTX wrappers -----------------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* tm, int c0, int c1, uint64_t* bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
               ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(tm)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// D[tmem] (+)= A[smem] * B[smem]^T, both operands K-major, fp32 bits consumed as tf32
__device__ __forceinline__ void tc_mma_tf32(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}\n"
      ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}
// 32 lanes x 32 consecutive columns: thread t of the warp receives lane (base + t), columns c..c+31
__device__ __forceinline__ void tc_ld32(uint32_t taddr, uint32_t* v) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
        "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
        "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
        "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr));
}
__device__ __forceinline__ void tc_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
// ties the loaded registers to a point after the wait, so no consumer can be scheduled ahead of it
__device__ __forceinline__ void tc_pin32(uint32_t* v) {
  asm volatile(""
               : "+r"(v[0]), "+r"(v[1]), "+r"(v[2]), "+r"(v[3]), "+r"(v[4]), "+r"(v[5]), "+r"(v[6]), "+r"(v[7]),
                 "+r"(v[8]), "+r"(v[9]), "+r"(v[10]), "+r"(v[11]), "+r"(v[12]), "+r"(v[13]), "+r"(v[14]), "+r"(v[15]),
                 "+r"(v[16]), "+r"(v[17]), "+r"(v[18]), "+r"(v[19]), "+r"(v[20]), "+r"(v[21]), "+r"(v[22]), "+r"(v[23]),
                 "+r"(v[24]), "+r"(v[25]), "+r"(v[26]), "+r"(v[27]), "+r"(v[28]), "+r"(v[29]), "+r"(v[30]), "+r"(v[31])
               :
               : "memory");
}

// shared-memory matrix descriptor: K-major, 128-byte swizzle, rows 128 bytes apart, 8-row groups 1024 bytes apart
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t saddr) {
  uint64_t d = (uint64_t)((saddr & 0x3ffffu) >> 4);
  d |= (uint64_t)1 << 16;            // leading byte offset (unused for swizzled K-major): 1
  d |= (uint64_t)(1024 >> 4) << 32;  // stride byte offset
  d |= (uint64_t)1 << 46;            // descriptor version (sm_100)
  d |= (uint64_t)2 << 61;            // SWIZZLE_128B
  return d;
}
// instruction descriptor: D fp32, A and B tf32, both K-major, N = 256, M = 128
constexpr uint32_t BQ_IDESC = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(BQ_N >> 3) << 17) | ((uint32_t)(BQ_M >> 4) << 24);

// ---- (1) per-row nomination coefficients --------------------------------------------------------------------
template <bool COSINE>
__global__ void __launch_bounds__(256)
row_prep_kernel(const float* __restrict__ X, int64_t n, int d, const uint8_t* __restrict__ skip,
                float2* __restrict__ ab, SegStats* __restrict__ stats) {
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
      float2 v;
      if (skip != nullptr && skip[r]) {
        v = make_float2(inf, 0.0f);  // never nominated
      } else {
        if (!(ss < 1e30f)) bad = 1;  // NaN, inf or out of the range the slack bound was derived for
        else vmax = max(vmax, __float_as_uint(ss));
        if (COSINE) v = make_float2(0.0f, ss > 0.0f ? -(1.0f / sqrtf(ss)) : 0.0f);
        else v = make_float2(ss, -2.0f);
      }
      ab[r] = v;
    }
  }
  if (lane == 0) {
    if (vmax) atomicMax(&stats->xmax2_bits, vmax);
    if (bad) atomicOr(&stats->nonfinite, 1);
  }
}

cudaError_t launch_row_prep(const float* X, int64_t n, int d, const uint8_t* skip, bool cosine, float2* ab,
                            SegStats* stats, int sms, cudaStream_t st) {
  cudaError_t e = cudaMemsetAsync(stats, 0, sizeof(SegStats), st);
  if (e != cudaSuccess) return e;
  int64_t grid = (n + 7) / 8;
  if (grid > (int64_t)sms * 8) grid = (int64_t)sms * 8;
  count_launch();
  if (cosine) row_prep_kernel<true><<<(int)grid, 256, 0, st>>>(X, n, d, skip, ab, stats);
  else row_prep_kernel<false><<<(int)grid, 256, 0, st>>>(X, n, d, skip, ab, stats);
  return cudaGetLastError();
}

// ---- (2) tensor-core nomination ---------------------------------------------------------------------------------
// grid.x = nqb * nsplit; CTA (qb, split) owns query block qb and row tiles split, split + nsplit, ...
// (CTAs of the same split run side by side, so a row tile is fetched from HBM once and re-read from L2).
__global__ void __launch_bounds__(BQ_THREADS, 1)
batch_gemm_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmX,
                  const float2* __restrict__ ab, int64_t n, int num_kb, int nqb, int64_t tiles,
                  float* __restrict__ gm, int64_t gm_stride) {
  extern __shared__ uint8_t bq_smem_raw[];
  uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(bq_smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* stages = base;
  float2* sm_ab = reinterpret_cast<float2*>(base + (size_t)BQ_STAGES * BQ_STAGE_BYTES);
  uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(sm_ab) + BQ_AB_BYTES);
  uint64_t* full = bars;                      // [BQ_STAGES] TMA -> MMA
  uint64_t* empty = bars + BQ_STAGES;         // [BQ_STAGES] MMA -> TMA
  uint64_t* tfull = bars + 2 * BQ_STAGES;     // [2] MMA -> epilogue
  uint64_t* tempty = bars + 2 * BQ_STAGES + 2;  // [2] epilogue -> MMA
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(bars + 2 * BQ_STAGES + 4);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int qb = blockIdx.x % nqb;
  const int split = blockIdx.x / nqb;
  const int nsplit = gridDim.x / nqb;

  if (threadIdx.x == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmQ)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmX)) : "memory");
    for (int s = 0; s < BQ_STAGES; s++) {
      mbar_init(full + s, 1);
      mbar_init(empty + s, 1);
    }
    for (int a = 0; a < 2; a++) {
      mbar_init(tfull + a, 1);
      mbar_init(tempty + a, 4);
    }
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

  if (warp == 0) {
    if (lane == 0) {  // ===== TMA producer =====
      int s = 0;
      uint32_t ph = 0;
      for (int64_t tile = split; tile < tiles; tile += nsplit) {
        for (int kb = 0; kb < num_kb; kb++) {
          mbar_wait(empty + s, ph ^ 1);
          mbar_expect_tx(full + s, BQ_STAGE_BYTES);
          uint8_t* st = stages + (size_t)s * BQ_STAGE_BYTES;
          tma_load_2d(st, &tmQ, kb * BQ_KB, qb * BQ_M, full + s);
          tma_load_2d(st + BQ_A_BYTES, &tmX, kb * BQ_KB, (int)(tile * BQ_N), full + s);
          if (++s == BQ_STAGES) {
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
      for (int64_t tile = split; tile < tiles; tile += nsplit, it++) {
        const int a = it & 1;
        mbar_wait(tempty + a, ((it >> 1) & 1) ^ 1);  // the epilogue has drained this accumulator
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)a * BQ_N;
        for (int kb = 0; kb < num_kb; kb++) {
          mbar_wait(full + s, ph);
          tc_fence_after();
          const uint32_t sa = smem_u32(stages + (size_t)s * BQ_STAGE_BYTES);
          const uint64_t adesc = umma_desc_sw128(sa), bdesc = umma_desc_sw128(sa + BQ_A_BYTES);
#pragma unroll
          for (int kk = 0; kk < BQ_KB / 8; kk++)  // 8 tf32 = 32 bytes per instruction: +2 in descriptor units
            tc_mma_tf32(d_tmem, adesc + 2 * kk, bdesc + 2 * kk, BQ_IDESC, (kb | kk) != 0 ? 1u : 0u);
          tc_commit(empty + s);  // frees the stage once these MMAs have read it
          if (++s == BQ_STAGES) {
            s = 0;
            ph ^= 1;
          }
        }
        tc_commit(tfull + a);
      }
    }
  } else if (warp >= 4) {  // ===== epilogue: TMEM -> per-group minima =====
    const int ew = warp - 4;  // TMEM lane quarter this warp may read
    const float inf = __int_as_float(0x7f800000);
    int it = 0;
    for (int64_t tile = split; tile < tiles; tile += nsplit, it++) {
      const int a = it & 1;
      float2* myab = sm_ab + (size_t)(ew * 2 + a) * BQ_N;
      for (int i = lane; i < BQ_N; i += 32) {
        const int64_t r = tile * BQ_N + i;
        myab[i] = r < n ? __ldg(ab + r) : make_float2(inf, 0.0f);
      }
      __syncwarp();
      mbar_wait(tfull + a, (it >> 1) & 1);
      tc_fence_after();
      float mins[4];
#pragma unroll
      for (int g = 0; g < 4; g++) {
        uint32_t v[64];
        const uint32_t taddr = tmem_base + ((uint32_t)(ew * 32) << 16) + (uint32_t)(a * BQ_N + g * 64);
        tc_ld32(taddr, v);
        tc_ld32(taddr + 32, v + 32);
        tc_wait_ld();
        tc_pin32(v);
        tc_pin32(v + 32);
        const float4* p4 = reinterpret_cast<const float4*>(myab + g * 64);
        float m0 = inf, m1 = inf;
#pragma unroll
        for (int j = 0; j < 32; j++) {
          const float4 p = p4[j];  // (alpha, beta) of two rows; same address in every lane: broadcast
          m0 = fminf(m0, fmaf(__uint_as_float(v[2 * j]), p.y, p.x));
          m1 = fminf(m1, fmaf(__uint_as_float(v[2 * j + 1]), p.w, p.z));
        }
        mins[g] = fminf(m0, m1);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty + a);
      float4* dst = reinterpret_cast<float4*>(gm + (size_t)(qb * BQ_M + ew * 32 + lane) * gm_stride + tile * 4);
      *dst = make_float4(mins[0], mins[1], mins[2], mins[3]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
  }
}

// ---- (3) threshold, candidate groups, exact ranking --------------------------------------------------------------
__device__ __forceinline__ double batch_slack(bool cosine, int d, double xmax, double qn) {
  const double c1 = (1.0 / 512.0) * 1.02 + (double)d * (1.0 / 1048576.0);  // tf32 operand truncation + accumulation
  const double c2 = (double)(d + 64) * (1.0 / 8388608.0);                  // fp32 rounding of alpha, of a, of the reference sums
  if (cosine) return (c1 + c2) * qn;
  return 2.0 * c1 * xmax * qn + c2 * (xmax + qn) * (xmax + qn);
}

// fb = [count][nq query indices][nq flags]: list query qi once, whichever CTA asks first
__device__ __forceinline__ void batch_list_fallback(int32_t* fb, int nq_total, int qi) {
  if (atomicExch(fb + 1 + nq_total + qi, 1) == 0) fb[1 + atomicAdd(fb, 1)] = qi;
}

// grid (P, nq): CTA (p, qi) owns the p-th slice of the 64-row groups of query qi.
template <int TPR, int U, bool COSINE, class TK>
__global__ void __launch_bounds__(BQ_SELECT_THREADS)
batch_select_kernel(const float* __restrict__ X, int64_t n, int d, const float* __restrict__ Q,
                    const uint8_t* __restrict__ skip, const float* __restrict__ gm, int64_t gm_stride,
                    int64_t ngroups, const SegStats* __restrict__ stats, int k, int kp, int cap,
                    int32_t* __restrict__ fb, int nq_total, TopkOut out) {
  extern __shared__ __align__(128) ulonglong2 smem[];
  constexpr int L = TPR * 4;
  constexpr int RPB = (32 / TPR) * U;  // rows per scoring batch
  constexpr int BPG = 64 / RPB;        // batches per group
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int nw = blockDim.x >> 5;
  const int qi = blockIdx.y;
  const int stride1 = kp + TOPK_BUF;
  const float* __restrict__ q = Q + (size_t)qi * d;
  const float* __restrict__ gmq = gm + (size_t)qi * gm_stride;

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
  for (int i = threadIdx.x; i < d; i += blockDim.x) qs[i] = q[i];
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
  const double slack = batch_slack(COSINE, d, sqrt((double)__uint_as_float(stats->xmax2_bits)), sqrt((double)q2));

  // phase 1: k-th smallest group minimum of this slice
  const int64_t g0 = ngroups * blockIdx.x / gridDim.x, g1 = ngroups * (blockIdx.x + 1) / gridDim.x;
  {
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
  }
  const Key kth = ld_key(smem + (k - 1));
  __syncthreads();  // everyone has read the k-th key: the collectors may be reused
  float tau = __int_as_float(0x7f800000);
  if (!key_is_empty(kth)) {
    const double T = f64_from_ordered(kth.hi);
    tau = f32_next_up(__double2float_ru(T + 2.0 * slack + 1e-37));
    if (!(tau == tau)) tau = __int_as_float(0x7f800000);
  }
  // phase 2: groups that can hold a top-k row
  for (int64_t g = g0 + threadIdx.x; g < g1; g += blockDim.x) {
    if (gmq[g] <= tau) {
      const int idx = atomicAdd(&s_cnt, 1);
      if (idx < cap) list[idx] = (int)g;
    }
  }
  __syncthreads();
  int cnt = s_cnt;
  if (cnt > cap) {  // adversarial data (e.g. thousands of duplicates): full exact scan for this query
    if (threadIdx.x == 0) batch_list_fallback(fb, nq_total, qi);
    cnt = cap;      // keep the ticket protocol of the epilogue intact; the fallback overwrites the result
  }
  // phase 3: exact scores of the candidate groups' rows, ranked like scan.cu
  TK tk;
  tk.init(smem + (size_t)warp * stride1, kp, k, lane);
  const int ub = d - (d % L);
  const int nv = ub / L;
  float fthr = __int_as_float(0x7f800000);
  uint64_t seen_hi = KEY_EMPTY64, seen_lo = KEY_EMPTY64;
  const int total = cnt * BPG;
  for (int b = warp; b < total; b += nw) {
    const int64_t row_base = (int64_t)list[b / BPG] * 64 + (int64_t)(b % BPG) * RPB;
    if (row_base >= n) continue;
    scan_batch_ldg<TPR, U, COSINE, TK>(X, n, d, q, qs, skip, row_base, nv, ub, qq, qn, fthr, tk, lane);
    if (tk.thr.hi != seen_hi || tk.thr.lo != seen_lo) {
      seen_hi = tk.thr.hi;
      seen_lo = tk.thr.lo;
      fthr = scan_filter_threshold<COSINE>(tk.thr);
    }
  }
  topk_epilogue(tk, smem, kp, k, out);
}

// ---- (4) exact scan of the queries listed in fb (count, then query indices) ----------------------------------------
template <int TPR, int U, bool COSINE, class TK>
__global__ void __launch_bounds__(SCAN_THREADS)
batch_fallback_kernel(const float* __restrict__ X, int64_t n, int d, const float* __restrict__ Q,
                      const uint8_t* __restrict__ skip, int k, int kp, const int32_t* __restrict__ fb, TopkOut out) {
  extern __shared__ __align__(128) ulonglong2 smem[];
  constexpr int L = TPR * 4;
  constexpr int G = 32 / TPR;
  const int count = fb[0];
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
    const int qi = fb[1 + slot];
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
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qr;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qr) != cudaSuccess ||
        qr != cudaDriverEntryPointSuccess)
      p = nullptr;
    return reinterpret_cast<EncodeTiledFn>(p);
  }();
  return fn;
}

// row-major fp32 matrix [rows][d] -> tensor map with a (32 floats x box_rows) box, 128-byte swizzle, zero fill
static bool encode_rows_map(CUtensorMap* tm, const float* base, int64_t rows, int d, int box_rows) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) return false;
  const cuuint64_t dims[2] = {(cuuint64_t)d, (cuuint64_t)rows};
  const cuuint64_t strides[1] = {(cuuint64_t)d * 4};
  const cuuint32_t box[2] = {(cuuint32_t)BQ_KB, (cuuint32_t)box_rows};
  const cuuint32_t estr[2] = {1, 1};
  return fn(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), dims, strides, box, estr,
            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

bool batch_encode_segment_map(void* tm128, const float* X, int64_t n, int d) {
  static_assert(sizeof(CUtensorMap) == 128, "tensor map size");
  return encode_rows_map(reinterpret_cast<CUtensorMap*>(tm128), X, n, d, BQ_N);
}

bool batch_supported(int d, int lanes, bool cosine, int64_t n) {
  if (!scan_is_streaming(d, lanes, cosine)) return false;  // the re-score uses the streaming row scorer
  if (d < BQ_KB || d > 65536) return false;
  if (n < 1 || n > (int64_t(1) << 31) - 2 * BQ_N) return false;  // 32-bit TMA coordinates, int group ids
  return encode_fn() != nullptr;
}

template <typename K>
static cudaError_t set_smem_attr(K kern, size_t smem) {
  if (smem > 48 * 1024) return cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  return cudaSuccess;
}

typedef void (*SelectKern)(const float*, int64_t, int, const float*, const uint8_t*, const float*, int64_t, int64_t,
                           const SegStats*, int, int, int, int32_t*, int, TopkOut);
typedef void (*FallbackKern)(const float*, int64_t, int, const float*, const uint8_t*, int, int, const int32_t*, TopkOut);

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

bool batch_configure(BatchLaunch& L, int sms) {
  L.kp = topk_pad(L.k);
  L.tiles = (L.n + BQ_N - 1) / BQ_N;
  L.ngroups = (L.n + 63) / 64;
  L.gm_stride = L.tiles * 4;
  L.gemm_smem = BQ_GEMM_SMEM;
  // candidate groups per select CTA: the k-th smallest group minimum admits about k groups, the slack a few more
  L.cap = 4 * L.k + 256;
  const int nw = BQ_SELECT_THREADS / 32;
  size_t coll = (size_t)nw * topk_warp_smem(L.kp);
  const size_t b3 = topk_block_smem(L.k, L.kp, nw);
  if (b3 > coll) coll = b3;
  coll = (coll + 15) & ~size_t(15);
  for (;;) {
    L.select_smem = coll + (size_t)((L.cap + 3) & ~3) * 4 + (((size_t)L.d * 4 + 15) & ~size_t(15));
    if (L.select_smem <= 200 * 1024 || L.cap <= L.k + 64) break;
    L.cap = L.cap / 2 > L.k + 64 ? L.cap / 2 : L.k + 64;
  }
  if (L.select_smem > 220 * 1024) return false;
  L.fb_threads = SCAN_THREADS;
  L.fb_smem = ((topk_block_smem(L.k, L.kp, L.fb_threads / 32) + 15) & ~size_t(15)) + (((size_t)L.d * 4 + 15) & ~size_t(15));
  if (L.fb_smem > 220 * 1024) return false;
  L.fb_gx = sms > TOPK_MAX_LISTS ? TOPK_MAX_LISTS : sms;
  L.sms = sms;
  const int TPR = L.lanes / 4;
  if (set_smem_attr(batch_gemm_kernel, L.gemm_smem) != cudaSuccess) return false;
  if (set_smem_attr(pick_select(TPR, L.cosine, L.k), L.select_smem) != cudaSuccess) return false;
  if (set_smem_attr(pick_fallback(TPR, L.cosine, L.k), L.fb_smem) != cudaSuccess) return false;
  return true;
}

int batch_select_ctas(const BatchLaunch& L, int nq) {
  // enough select CTAs to fill the GPU when the batch is small; slices no shorter than 2048 groups
  int64_t P = (2 * (int64_t)L.sms + nq - 1) / nq;
  const int64_t by_len = (L.ngroups + 2047) / 2048;
  if (P > by_len) P = by_len;
  if (P > 64) P = 64;
  return (int)(P < 1 ? 1 : P);
}

int64_t batch_partial_keys(const BatchLaunch& L, int nq) {
  const int P = batch_select_ctas(L, nq);
  return (int64_t)(P > L.fb_gx ? P : L.fb_gx) * L.k;
}

// One chunk of nq queries (q, outputs and scratch already offset to the chunk).
cudaError_t launch_batch(const BatchLaunch& L, cudaStream_t st) {
  cudaError_t e;
  CUtensorMap tmQ;
  if (!encode_rows_map(&tmQ, L.q, L.nq, L.d, BQ_M)) return cudaErrorInvalidValue;
  const CUtensorMap tmX = *reinterpret_cast<const CUtensorMap*>(L.tmX);
  const int nqb = (L.nq + BQ_M - 1) / BQ_M;
  int nsplit = L.sms / nqb;
  if (nsplit < 1) nsplit = 1;
  if (nsplit > L.tiles) nsplit = (int)L.tiles;
  if ((e = cudaMemsetAsync(L.fb, 0, sizeof(int32_t) * (1 + 2 * (size_t)L.nq), st)) != cudaSuccess) return e;
  count_launch();
  batch_gemm_kernel<<<nqb * nsplit, BQ_THREADS, L.gemm_smem, st>>>(tmQ, tmX, L.ab, L.n, (L.d + BQ_KB - 1) / BQ_KB, nqb,
                                                                   L.tiles, L.gm, L.gm_stride);
  if ((e = cudaGetLastError()) != cudaSuccess) return e;
  const int TPR = L.lanes / 4;
  const int P = batch_select_ctas(L, L.nq);
  TopkOut o{L.partial, L.ctrl, L.partial_keys, L.ids_out, L.scores_out, L.counts_out, L.id_base, 0,
            L.out_stride > 0 ? L.out_stride : L.k};
  count_launch();
  pick_select(TPR, L.cosine, L.k)<<<dim3(P, L.nq), BQ_SELECT_THREADS, L.select_smem, st>>>(
      L.X, L.n, L.d, L.q, L.skip, L.gm, L.gm_stride, L.ngroups, L.stats, L.k, L.kp, L.cap, L.fb, L.nq, o);
  if ((e = cudaGetLastError()) != cudaSuccess) return e;
  count_launch();
  const int gy = L.nq < BQ_FB_SLOTS ? L.nq : BQ_FB_SLOTS;
  pick_fallback(TPR, L.cosine, L.k)<<<dim3(L.fb_gx, gy), L.fb_threads, L.fb_smem, st>>>(
      L.X, L.n, L.d, L.q, L.skip, L.k, L.kp, L.fb, o);
  return cudaGetLastError();
}

}  // namespace vs
