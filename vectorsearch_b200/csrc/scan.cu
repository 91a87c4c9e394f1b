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
//  * Rows reach the SM through TMA: every warp owns a private ring of NS shared-memory stages,
//    each filled by ONE bulk asynchronous copy (cp.async.bulk, mbarrier complete_tx) of TR
//    consecutive rows; the warp waits on a stage, scores its rows from shared memory, and lane 0
//    re-arms the stage with the tile NS steps ahead.  Bytes in flight are decoupled from
//    registers and from the (serial) reduction chain, which is what keeps HBM busy.
//    (Rows too long for the ring use the register-staged variant below: 128-bit streaming loads,
//    32/TPR rows x U in flight per warp.)
//  * Top-k: one fp32 compare per row against a warp-uniform threshold; only rows that can
//    still enter the top-k get their fp64 score (sqrt / divide) and go through topk.cuh.
//  * grid = SMs x resident CTAs; warps stride over row batches.  The last CTA to finish merges
//    the per-CTA lists and writes ids/scores (no second launch).
#include "kernels.h"
#include "topk.cuh"
#include "scan_rows.cuh"

namespace vs {

template <int TPR, int U, bool COSINE, class TK>
__global__ void __launch_bounds__(SCAN_THREADS)
scan_kernel(const float* __restrict__ X, int64_t n, int d, const float* __restrict__ Q,
            const uint8_t* __restrict__ skip, int k, int kp, TopkOut out) {
  extern __shared__ __align__(128) ulonglong2 smem[];
  constexpr int L = TPR * 4;
  constexpr int G = 32 / TPR;  // rows per warp per unroll slot
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int nw = blockDim.x >> 5;
  const int stride_keys = kp + TOPK_BUF;
  const float* __restrict__ q = Q + (size_t)blockIdx.y * d;

  TK tk;
  tk.init(smem + (size_t)warp * stride_keys, kp, k, lane);
  float* qs = reinterpret_cast<float*>(reinterpret_cast<unsigned char*>(smem) + topk_block_smem(k, kp, nw));
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
    scan_batch_ldg<TPR, U, COSINE, TK>(X, n, d, q, qs, skip, b * rows_per_batch, nv, ub, qq, qn, fthr, tk, lane);
    if (tk.thr.hi != seen_hi || tk.thr.lo != seen_lo) {
      seen_hi = tk.thr.hi;
      seen_lo = tk.thr.lo;
      fthr = scan_filter_threshold<COSINE>(tk.thr);
    }
  }
  topk_epilogue(tk, smem, kp, k, out);
}

// ---- TMA-fed variant ---------------------------------------------------------------------------------
// TR rows per stage (multiple of U * 32/TPR), NS stages per warp.
template <int TPR, int U, bool COSINE, class TK>
__global__ void __launch_bounds__(SCAN_THREADS, 1)
scan_tma_kernel(const float* __restrict__ X, int64_t n, int d, const float* __restrict__ Q,
                const uint8_t* __restrict__ skip, int k, int kp, int TR, int NS, TopkOut out) {
  extern __shared__ __align__(128) ulonglong2 smem[];
  phase_stamp(0);
  constexpr int L = TPR * 4;
  constexpr int G = 32 / TPR;
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int nw = blockDim.x >> 5;
  const int stride_keys = kp + TOPK_BUF;
  const int t = lane & (TPR - 1);
  const int g = lane / TPR;
  const float* __restrict__ q = Q + (size_t)blockIdx.y * d;
  const size_t row_bytes = (size_t)d * 4;
  const uint32_t stage_bytes = (uint32_t)(TR * row_bytes);

  // shared memory: [rings: nw * NS * stage][collectors][q][barriers]
  unsigned char* base = reinterpret_cast<unsigned char*>(smem);
  unsigned char* ring = base + (size_t)warp * NS * stage_bytes;
  ulonglong2* coll = reinterpret_cast<ulonglong2*>(base + (size_t)nw * NS * stage_bytes);
  float* qs = reinterpret_cast<float*>(reinterpret_cast<unsigned char*>(coll) + topk_block_smem(k, kp, nw));
  uint64_t* bars = reinterpret_cast<uint64_t*>(qs + ((d + 3) & ~3)) + warp * NS;

  TK tk;
  tk.init(coll + (size_t)warp * stride_keys, kp, k, lane);
  for (int i = threadIdx.x; i < d; i += blockDim.x) qs[i] = q[i];
  __shared__ double s_qq;
  if (COSINE && threadIdx.x == 0) s_qq = ref_sum_thread_L<L, REF_DOT>(q, q, d);
  if (lane == 0) {
    for (int s = 0; s < NS; s++) mbar_init(bars + s, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const double qq = COSINE ? s_qq : 0.0;
  const float qn = COSINE ? (float)sqrt(qq) : 0.0f;

  const int ub = d - (d % L);
  const int nv = ub / L;
  const int64_t ntiles = (n + TR - 1) / TR;
  const int64_t first = (int64_t)blockIdx.x * nw + warp, step = (int64_t)gridDim.x * nw;
  auto tile_bytes = [&](int64_t tile) -> uint32_t {
    const int64_t rows = (tile + 1) * TR <= n ? TR : n - tile * TR;
    return (uint32_t)(rows * row_bytes);
  };
  if (lane == 0) {
    for (int s = 0; s < NS; s++) {
      const int64_t tile = first + (int64_t)s * step;
      if (tile < ntiles) {
        const uint32_t b = tile_bytes(tile);
        mbar_expect_tx(bars + s, b);
        bulk_g2s(ring + (size_t)s * stage_bytes, X + (size_t)tile * TR * d, b, bars + s);
      }
    }
  }
  float fthr = __int_as_float(0x7f800000);
  uint64_t seen_hi = KEY_EMPTY64, seen_lo = KEY_EMPTY64;
  int stage = 0;
  uint32_t parity = 0;
  phase_stamp(1);
  for (int64_t tile = first; tile < ntiles; tile += step) {
    mbar_wait(bars + stage, parity);
    const unsigned char* sp = ring + (size_t)stage * stage_bytes;
    for (int r0 = 0; r0 < TR; r0 += G * U) {
      if (tk.thr.hi != seen_hi || tk.thr.lo != seen_lo) {
        seen_hi = tk.thr.hi;
        seen_lo = tk.thr.lo;
        fthr = scan_filter_threshold<COSINE>(tk.thr);
      }
      float acc[U][4];
      float accn[COSINE ? U : 1][4];
      const float4* rp[U];
#pragma unroll
      for (int u = 0; u < U; u++) {
        rp[u] = reinterpret_cast<const float4*>(sp + (size_t)(r0 + u * G + g) * row_bytes) + t;
#pragma unroll
        for (int c = 0; c < 4; c++) acc[u][c] = 0.0f;
        if (COSINE) {
#pragma unroll
          for (int c = 0; c < 4; c++) accn[u][c] = 0.0f;
        }
      }
#pragma unroll 4
      for (int i = 0; i < nv; i++) {
        const float4 q4 = reinterpret_cast<const float4*>(qs)[i * TPR + t];
#pragma unroll
        for (int u = 0; u < U; u++) {
          const float4 x = rp[u][(size_t)i * TPR];
          if (COSINE) {
            acc[u][0] = __fmaf_rn(x.x, q4.x, acc[u][0]);
            acc[u][1] = __fmaf_rn(x.y, q4.y, acc[u][1]);
            acc[u][2] = __fmaf_rn(x.z, q4.z, acc[u][2]);
            acc[u][3] = __fmaf_rn(x.w, q4.w, acc[u][3]);
            accn[u][0] = __fmaf_rn(x.x, x.x, accn[u][0]);
            accn[u][1] = __fmaf_rn(x.y, x.y, accn[u][1]);
            accn[u][2] = __fmaf_rn(x.z, x.z, accn[u][2]);
            accn[u][3] = __fmaf_rn(x.w, x.w, accn[u][3]);
          } else {
            const float dx = __fsub_rn(q4.x, x.x), dy = __fsub_rn(q4.y, x.y);
            const float dz = __fsub_rn(q4.z, x.z), dw = __fsub_rn(q4.w, x.w);
            acc[u][0] = __fmaf_rn(dx, dx, acc[u][0]);
            acc[u][1] = __fmaf_rn(dy, dy, acc[u][1]);
            acc[u][2] = __fmaf_rn(dz, dz, acc[u][2]);
            acc[u][3] = __fmaf_rn(dw, dw, acc[u][3]);
          }
        }
      }
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
#pragma unroll
      for (int u = 0; u < U; u++) {
        const int64_t row = tile * TR + r0 + u * G + g;
        float est;
        if (COSINE) {
          est = -(s[u] / (qn * sqrtf(sn[u])));
        } else {
          est = s[u];
        }
        bool cand = (t == TPR - 1) && row < n && !(est > fthr);
        if (!COSINE && ub < d && (t == TPR - 1) && row < n && !cand)  // a NaN in the fp64 tail only (scan_rows.cuh)
          cand = scan_tail_has_nan(reinterpret_cast<const float*>(sp + (size_t)(r0 + u * G + g) * row_bytes), ub, d);
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
    __syncwarp();
    if (lane == 0) {
      const int64_t nt = tile + (int64_t)NS * step;
      if (nt < ntiles) {
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        const uint32_t b = tile_bytes(nt);
        mbar_expect_tx(bars + stage, b);
        bulk_g2s(ring + (size_t)stage * stage_bytes, X + (size_t)nt * TR * d, b, bars + stage);
      }
    }
    if (++stage == NS) {
      stage = 0;
      parity ^= 1;
    }
#ifdef VS_PHASE_STAMPS
    if (tile == first + 4 * step) phase_stamp(2);
#endif
  }
  phase_stamp(3);
  topk_epilogue(tk, coll, kp, k, out);
  phase_stamp(4);
}

// Any d, any lane count: one thread per row, reference arithmetic straight from common.cuh.
// Uncoalesced; used only for shapes the streaming kernel does not take (d % 4 != 0 or d < L).
template <bool COSINE, class TK>
__global__ void __launch_bounds__(SCAN_THREADS)
scan_rowthread_kernel(const float* __restrict__ X, int64_t n, int d, const float* __restrict__ Q,
                      const uint8_t* __restrict__ skip, int lanes, int k, int kp, TopkOut out) {
  extern __shared__ __align__(128) ulonglong2 smem[];
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int nw = blockDim.x >> 5;
  const int stride_keys = kp + TOPK_BUF;
  const float* __restrict__ q = Q + (size_t)blockIdx.y * d;
  TK tk;
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
// host: configuration + launch
// ------------------------------------------------------------------------------------------------
template <typename K>
static cudaError_t set_smem(K kern, size_t smem) {
  if (smem > 48 * 1024)
    return cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  return cudaSuccess;
}

// the cosine pre-filter needs the whole dot product in the fp32 lanes (no fp64 tail)
bool scan_is_streaming(int d, int lanes, bool cosine) {
  if ((d % 4) != 0 || d < lanes || lanes < 4) return false;
  return !cosine || (d % lanes) == 0;
}

template <typename K>
static int occ_of(K kern, int threads, size_t smem) {
  if (set_smem(kern, smem) != cudaSuccess) return 0;
  int nb = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kern, threads, smem) != cudaSuccess) nb = 0;
  return nb;
}

constexpr size_t SCAN_SMEM_BUDGET = 200 * 1024;

// kernel pointer for (variant, TPR, cosine, collector)
typedef void (*LdgKern)(const float*, int64_t, int, const float*, const uint8_t*, int, int, TopkOut);
typedef void (*TmaKern)(const float*, int64_t, int, const float*, const uint8_t*, int, int, int, int, TopkOut);
typedef void (*RowKern)(const float*, int64_t, int, const float*, const uint8_t*, int, int, int, TopkOut);

template <class TK>
static LdgKern ldg_kernel(int TPR, bool cosine) {
  if (cosine) return TPR == 4 ? scan_kernel<4, 1, true, TK> : (TPR == 2 ? scan_kernel<2, 1, true, TK> : scan_kernel<1, 1, true, TK>);
  return TPR == 4 ? scan_kernel<4, 2, false, TK> : (TPR == 2 ? scan_kernel<2, 2, false, TK> : scan_kernel<1, 2, false, TK>);
}
template <class TK>
static TmaKern tma_kernel(int TPR, bool cosine) {
  if (cosine) return TPR == 4 ? scan_tma_kernel<4, 1, true, TK> : (TPR == 2 ? scan_tma_kernel<2, 1, true, TK> : scan_tma_kernel<1, 1, true, TK>);
  return TPR == 4 ? scan_tma_kernel<4, 2, false, TK> : (TPR == 2 ? scan_tma_kernel<2, 2, false, TK> : scan_tma_kernel<1, 2, false, TK>);
}
template <class TK>
static RowKern row_kernel(bool cosine) {
  return cosine ? scan_rowthread_kernel<true, TK> : scan_rowthread_kernel<false, TK>;
}
static LdgKern pick_ldg(int TPR, bool cosine, int k) {
  return k <= TOPK_REG_MAX_K ? ldg_kernel<WarpTopKReg>(TPR, cosine) : ldg_kernel<WarpTopK>(TPR, cosine);
}
static TmaKern pick_tma(int TPR, bool cosine, int k) {
  return k <= TOPK_REG_MAX_K ? tma_kernel<WarpTopKReg>(TPR, cosine) : tma_kernel<WarpTopK>(TPR, cosine);
}
static RowKern pick_row(bool cosine, int k) {
  return k <= TOPK_REG_MAX_K ? row_kernel<WarpTopKReg>(cosine) : row_kernel<WarpTopK>(cosine);
}

static int clamp_grid(int64_t grid, int64_t cap, int k) {
  if (k <= TOPK_REG_MAX_K && grid > TOPK_MAX_LISTS) grid = TOPK_MAX_LISTS;
  if (grid > cap) grid = cap;
  return (int)(grid < 1 ? 1 : grid);
}

// Fills variant / threads / grid / smem_bytes / TR / NS of L from (d, lanes, cosine, k, n, nq).
// Returns false when no variant can be resident.
bool scan_configure(ScanLaunch& L, int sms) {
  L.kp = topk_pad(L.k);
  const size_t qbytes = (((size_t)L.d * 4 + 15) / 16) * 16;
  const int TPR = L.lanes / 4, G = 32 / (TPR > 0 ? TPR : 1);
  if (!scan_is_streaming(L.d, L.lanes, L.cosine)) {
    L.variant = SCAN_ROWTHREAD;
    L.threads = SCAN_THREADS;
    while (topk_block_smem(L.k, L.kp, L.threads / 32) > SCAN_SMEM_BUDGET && L.threads > 32) L.threads /= 2;
    L.smem_bytes = topk_block_smem(L.k, L.kp, L.threads / 32);
    const int occ = occ_of(pick_row(L.cosine, L.k), L.threads, L.smem_bytes);
    if (occ < 1) return false;
    L.grid = clamp_grid((int64_t)sms * (L.nq > 1 ? 1 : (occ > 2 ? 2 : occ)), (L.n + 255) / 256, L.k);
    L.partial_keys = topk_partial_keys(L.k, L.grid, L.threads / 32);
    return true;
  }
  // TMA ring: nw warps x NS stages x TR rows
  for (int nw = 8; nw >= 2; nw /= 2) {
    const int U = L.cosine ? 1 : 2;
    const int TR = G * U;
    const size_t stage = (size_t)TR * L.d * 4;
    const size_t fixed = ((topk_block_smem(L.k, L.kp, nw) + 15) / 16) * 16 + qbytes + (size_t)nw * 4 * 8 + 256;
    if (fixed >= SCAN_SMEM_BUDGET) continue;
    int NS = (int)((SCAN_SMEM_BUDGET - fixed) / ((size_t)nw * stage));
    if (NS > 4) NS = 4;
    if (NS < 2) continue;
    if ((size_t)nw * NS * stage > 160 * 1024 && NS > 2) NS = 2;
    L.variant = SCAN_TMA;
    L.threads = nw * 32;
    L.TR = TR;
    L.NS = NS;
    L.smem_bytes = (size_t)nw * NS * stage + fixed;
    const int64_t ntiles = (L.n + TR - 1) / TR;
    L.grid = clamp_grid(sms, (ntiles + nw - 1) / nw, L.k);
    L.partial_keys = topk_partial_keys(L.k, L.grid, nw);
    return true;
  }
  // rows too long for the ring: register-staged streaming loads
  L.variant = SCAN_LDG;
  L.threads = SCAN_THREADS;
  while (topk_block_smem(L.k, L.kp, L.threads / 32) + qbytes > SCAN_SMEM_BUDGET && L.threads > 32) L.threads /= 2;
  L.smem_bytes = ((topk_block_smem(L.k, L.kp, L.threads / 32) + 15) / 16) * 16 + qbytes;
  const int occ = occ_of(pick_ldg(TPR, L.cosine, L.k), L.threads, L.smem_bytes);
  if (occ < 1) return false;
  L.grid = clamp_grid((int64_t)sms * (L.nq > 1 ? 1 : (occ > 2 ? 2 : occ)), (L.n + 15) / 16, L.k);
  L.partial_keys = topk_partial_keys(L.k, L.grid, L.threads / 32);
  return true;
}

#ifdef VS_PHASE_STAMPS
int debug_read_stamps(void* dst, size_t bytes) {
  cudaDeviceSynchronize();
  return (int)cudaMemcpyFromSymbol(dst, g_phase_stamps, bytes);
}
#endif

cudaError_t launch_scan(const ScanLaunch& L, cudaStream_t st) {
  TopkOut o{L.partial, L.ctrl, L.partial_keys, L.ids_out, L.scores_out, L.counts_out, L.id_base, 0, L.out_stride > 0 ? L.out_stride : L.k};
  count_launch();
  const int TPR = L.lanes / 4;
  const dim3 grid(L.grid, L.nq);
  cudaError_t e;
  if (L.variant == SCAN_ROWTHREAD) {
    RowKern kern = pick_row(L.cosine, L.k);
    if ((e = set_smem(kern, L.smem_bytes)) != cudaSuccess) return e;
    kern<<<grid, L.threads, L.smem_bytes, st>>>(L.X, L.n, L.d, L.q, L.skip, L.lanes, L.k, L.kp, o);
  } else if (L.variant == SCAN_TMA) {
    TmaKern kern = pick_tma(TPR, L.cosine, L.k);
    if ((e = set_smem(kern, L.smem_bytes)) != cudaSuccess) return e;
    kern<<<grid, L.threads, L.smem_bytes, st>>>(L.X, L.n, L.d, L.q, L.skip, L.k, L.kp, L.TR, L.NS, o);
  } else {
    LdgKern kern = pick_ldg(TPR, L.cosine, L.k);
    if ((e = set_smem(kern, L.smem_bytes)) != cudaSuccess) return e;
    kern<<<grid, L.threads, L.smem_bytes, st>>>(L.X, L.n, L.d, L.q, L.skip, L.k, L.kp, o);
  }
  return cudaGetLastError();
}

}  // namespace vs
