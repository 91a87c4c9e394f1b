// adc_fast.cu -- K6: the full-segment ADC scan as a conflict-free shared-memory gather.
//
// Replaces the sealed-segment scan loop + sort of searchSealedSegment
// (J/fdb/FdbVectorIndex.java:754-769, first n_cand of the ascending stable sort :820-822) for
// M <= 16 subspaces of K <= 256 centroids.  Algorithmic traffic: M bytes per distance evaluation.
//
// Why not the obvious kernel (adc.cu): M random 4-byte shared-memory lookups per row serialise on
// bank conflicts (~3.5-way for 32 random addresses) and every survivor of a per-warp threshold
// walks a divergent fp64 path; measured 390 GB/s of code bytes (6% of HBM).  This kernel:
//  * quantises the query's LUT (the reference's doubles, buildLut :1067-1079; build_lut_mm_kernel)
//    to ONE BYTE per entry, q = floor((lut[s][c] - min_s) / delta), with one global step delta, so that
//        sum_s min_s + delta * S  <=  pqApproxDistance  <=  sum_s min_s + delta * (S + M),
//    S = sum_s q[s][code_s] an integer <= 255*M;
//  * keeps 32 copies of the M x 256 byte table in shared memory, copy l entirely inside bank l
//    (the table sits at the FIXED shared-memory window address 0x4000; byte (s, c) of lane l at
//    0x4000 + (s>>2)*32768 + ((c&0x7c) + (s&3))*256 + (c&0x80) + l*4 + (c&3): four subspaces interleave their
//    256-byte lines, so bits 2..6 of a code are the high address byte where they stand, without a shift; the table
//    address and the subspace's offset are the immediate of the load, and ONE byte permute per code assembles the rest):
//    lane l only ever touches bank l, so every LDS.U8 of a warp is conflict-free for ANY combination of codes;
//  * filters rows on the integer S against a device-wide threshold T that is maintained from a
//    global histogram of S: if b* is the k-th smallest S seen so far, every row of the final
//    top-k has S <= b* + M + 2 (the bound above plus rounding slack), so T = b* + M + 2 never
//    rejects a row the reference would return.  Rows with S <= T are appended to a candidate
//    list (warp-aggregated atomics; a few thousand rows out of 1e8 after the warm-up);
//    list (per-CTA lists and a per-CTA shared-memory histogram that is drained into the global one,
//    so the hot loop issues no same-address global atomics);
//  * the last CTA to finish evaluates the candidates with S <= T_final exactly like
//    pqApproxDistance (:1057-1065: fp64 adds in subspace order, codes >= K skipped), sorts them by
//    (distance, row) in shared memory and writes the first k: ids and distances are the
//    reference's, ties go to the lowest row.  One launch per query batch.
// Degenerate tables (non-finite entries, zero range), candidate-list overflow and more survivors
// than the sort can hold raise a flag instead; adc_fallback_kernel (a no-op unless the flag is
// set) then evaluates every row exactly: slower, never wrong.
#include "kernels.h"
#include "topk.cuh"

namespace vs {

__device__ __forceinline__ unsigned int ld_cg_u32(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.global.cg.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

__device__ __forceinline__ double fs_exact(const double* __restrict__ lut, const uint8_t* __restrict__ cr,
                                           int M, int K) {
  double ad = 0.0;
  for (int s = 0; s < M; s++) {
    const int ci = cr[s];
    if (ci >= K) continue;
    ad = __dadd_rn(ad, lut[(size_t)s * K + ci]);
  }
  return ad;
}

// development counters: [0] candidates listed, [1] T_final, [2] survivors ranked exactly, [3] fallback runs
__device__ unsigned int g_adc_dbg[8];

// k-th smallest bin (1-based k) of the histogram, looking at bins [0, nb) rounded up to a multiple
// of 128, by one warp; FS_T_INF when fewer than k entries are counted there.  Lane l reads bins
// 128*i + 4*l .. + 3 of group i (coalesced 16-byte loads, four groups in flight).  Counts only grow
// while this runs, which can only move the answer down.
__device__ __forceinline__ unsigned int fs_warp_kth(const unsigned int* hist, unsigned int nb, unsigned int k,
                                                    int lane) {
  const uint4* h4 = reinterpret_cast<const uint4*>(hist);
  const unsigned int ngroups = (nb + 127) / 128;  // <= FS_BINS / 128
  unsigned int run = 0;
  for (unsigned int g0 = 0; g0 < ngroups; g0 += 4) {
    uint4 v[4];
#pragma unroll
    for (int j = 0; j < 4; j++) {
      const unsigned int g = g0 + j < (unsigned int)(FS_BINS / 128) ? g0 + j : (unsigned int)(FS_BINS / 128) - 1;
      v[j] = __ldcg(h4 + g * 32 + lane);
    }
#pragma unroll
    for (int j = 0; j < 4; j++) {
      if (g0 + j >= ngroups) break;
      const unsigned int s4 = v[j].x + v[j].y + v[j].z + v[j].w;
      unsigned int incl = s4;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const unsigned int t = __shfl_up_sync(FULL_MASK, incl, o);
        if (lane >= o) incl += t;
      }
      const unsigned int total = __shfl_sync(FULL_MASK, incl, 31);
      if (run + total >= k) {
        unsigned int found = FS_T_INF;
        const unsigned int excl = run + incl - s4;
        if (excl < k && excl + s4 >= k) {
          const unsigned int b = (g0 + j) * 128 + lane * 4;
          unsigned int acc = excl + v[j].x;
          if (acc >= k) found = b;
          else if ((acc += v[j].y) >= k) found = b + 1;
          else if ((acc += v[j].z) >= k) found = b + 2;
          else found = b + 3;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) found = min(found, __shfl_xor_sync(FULL_MASK, found, o));
        return found;
      }
      run += total;
    }
  }
  return FS_T_INF;
}

// {0, 0, hi.byte[B], lo.byte[B]}: selector nibble 8|n replicates the (clear) sign bit of byte n.
// (__byte_perm drops bit 3 of the selector nibbles, hence prmt directly.)
template <int B>
__device__ __forceinline__ unsigned int fs_addr(unsigned int lo, unsigned int hi) {
  unsigned int off;
  asm("prmt.b32 %0, %1, %2, %3;" : "=r"(off) : "r"(lo), "r"(hi), "n"(((0xc + B) << 12) | ((0xc + B) << 8) | ((4 + B) << 4) | B));
  return off;
}
// One table byte: LDS.U8 [R + imm] with R the ABSOLUTE shared-memory address PRMT assembled.  (Through a C++ pointer
// the compiler kept the table base in a register and spent an IMAD per lookup adding it -- a quarter of the loop's
// instructions, in a loop bound by instruction issue.)
constexpr unsigned int FS_TABLE_ADDR = 0x4000u;  // of the table inside the CTA's shared-memory window (a constant, so that
                                                 // it can ride in the immediate of every lookup)
template <int IMM>
__device__ __forceinline__ unsigned int fs_lds(unsigned int addr) {
  unsigned int v;
  asm("ld.shared.u8 %0, [%1+%2];" : "=r"(v) : "r"(addr), "n"(IMM));
  return v;
}
// immediate part of the address of subspace S's entries: the table's address + the subspace's place in it
template <int S>
struct FsImm {
  static constexpr int value = (int)FS_TABLE_ADDR + (S >> 2) * 32768 + (S & 3) * 256;
};
// high / low address bytes of the four codes of a word: TWO word-wide operations, then ONE byte permute per code
// (both bytes of {hi, lo} are offsets from the immediate; hi <= 124 keeps the sign bit the permute replicates clear)
__device__ __forceinline__ unsigned int fs_hi(unsigned int w) { return w & 0x7c7c7c7cu; }
__device__ __forceinline__ unsigned int fs_lo(unsigned int w, unsigned int lane4x4) { return (w & 0x83838383u) | lane4x4; }
// sum of the four table bytes selected by the codes of word J of a row (subspaces 4J .. 4J+3);
// lane4x4 = lane*4 in every byte: the lane's bank enters through the low address byte
template <int J>
__device__ __forceinline__ unsigned int fs_word_sum(unsigned int w, unsigned int lane4x4) {
  const unsigned int hi = fs_hi(w), lo = fs_lo(w, lane4x4);
  unsigned int acc = fs_lds<FsImm<4 * J + 0>::value>(fs_addr<0>(lo, hi));
  acc += fs_lds<FsImm<4 * J + 1>::value>(fs_addr<1>(lo, hi));
  acc += fs_lds<FsImm<4 * J + 2>::value>(fs_addr<2>(lo, hi));
  acc += fs_lds<FsImm<4 * J + 3>::value>(fs_addr<3>(lo, hi));
  return acc;
}
// the same for word J split in two: codes 0-1 of the word (HALF = 0) or codes 2-3 (HALF = 1)
template <int J, int HALF>
__device__ __forceinline__ unsigned int fs_half_word_sum(unsigned int lo, unsigned int hi) {
  unsigned int acc = fs_lds<FsImm<4 * J + 2 * HALF + 0>::value>(fs_addr<2 * HALF + 0>(lo, hi));
  acc += fs_lds<FsImm<4 * J + 2 * HALF + 1>::value>(fs_addr<2 * HALF + 1>(lo, hi));
  return acc;
}
// byte offset of entry c of subspace s inside lane 0's copy, from the table's start (add lane*4)
__host__ __device__ inline unsigned int fs_entry_offset(unsigned int s, unsigned int c) {
  return (s >> 2) * 32768u + ((c & 0x7cu) + (s & 3u)) * 256u + (c & 0x83u);
}

// ---- LUT build: one CTA per (subspace, query) ----------------------------------------------------------
// lut64[q][s][c] = Distances.l2Squared(query, s*subDim, centroids[s][c], 0, subDim) (buildLut,
// J/fdb/FdbVectorIndex.java:1067-1079) in reference arithmetic, plus what the scan needs to derive
// the byte image: mm[q][s] = {min, max} of the subspace's entries as order-preserving uint64 images
// (max = ~0 when an entry is NaN or infinite).
__global__ void __launch_bounds__(256)
build_lut_mm_kernel(const float* __restrict__ centroids, int M, int K, int subDim, const float* __restrict__ Q,
                    int lanes, double* __restrict__ LUT64, unsigned long long* __restrict__ MM) {
  pdl_trigger();  // the scan's CTAs may be placed (they wait for this grid before reading the table)
  const int s = blockIdx.x, qi = blockIdx.y;
  const int lane = threadIdx.x & 31;
  const float* q = Q + (size_t)qi * M * subDim + (size_t)s * subDim;
  __shared__ unsigned long long s_mn, s_mx;
  if (threadIdx.x == 0) {
    s_mn = ~0ull;
    s_mx = 0ull;
  }
  __syncthreads();
  for (int c0 = 0; c0 < K; c0 += blockDim.x) {
    const int c = c0 + threadIdx.x;
    unsigned long long lo = ~0ull, hi = 0ull;
    if (c < K) {
      const double v = ref_sum_thread<REF_L2SQ>(q, centroids + ((size_t)s * K + c) * subDim, subDim, lanes);
      LUT64[((size_t)qi * M + s) * K + c] = v;
      lo = hi = f64_ordered(v);
      if (!(fabs(v) <= 1.7976931348623157e308)) hi = ~0ull;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      lo = min(lo, __shfl_xor_sync(FULL_MASK, lo, o));
      hi = max(hi, __shfl_xor_sync(FULL_MASK, hi, o));
    }
    if (lane == 0) {
      atomicMin(&s_mn, lo);
      atomicMax(&s_mx, hi);
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    MM[((size_t)qi * M + s) * 2 + 0] = s_mn;
    MM[((size_t)qi * M + s) * 2 + 1] = s_mx;
  }
}

// MW = M/4 words of codes per row (M = 8 or 16).
template <int MW>
__global__ void __launch_bounds__(FS_THREADS, 1)
adc_fastscan_kernel(const uint8_t* __restrict__ codes, int64_t n, int K,
                    const double* LUT64 /* written by the LUT kernel this one may overlap (launch_pdl): no __restrict__ */,
                    const unsigned long long* MM, unsigned int k, unsigned int* __restrict__ fs, unsigned long long* __restrict__ cand_all,
                    unsigned int cap, int64_t* __restrict__ ids_out, double* __restrict__ approx_out,
                    int32_t* __restrict__ counts_out, int64_t id_base, int64_t out_stride) {
  extern __shared__ __align__(128) unsigned char fsm[];
  pdl_trigger();  // the fallback check may be set up and placed while this kernel runs
  pdl_wait();     // the LUT kernel's output
  constexpr int M = MW * 4;
  constexpr int BPT = FS_BINS / FS_THREADS;  // histogram bins owned by a thread
  const int tid = threadIdx.x;
  const int lane = tid & 31;
  const int warp = tid >> 5;
  const int qi = blockIdx.y;
  const double* lut64 = LUT64 + (size_t)qi * M * K;
  unsigned int* hist_g = fs + (size_t)qi * FS_WORDS;
  unsigned int* ctrl = hist_g + FS_BINS;
  unsigned int* counts_g = ctrl + FS_CTRL;
  unsigned long long* cand_q = cand_all + (size_t)qi * gridDim.x * cap;
  unsigned long long* cand = cand_q + (size_t)blockIdx.x * cap;

  // the table starts at window address FS_TABLE_ADDR (what lies below it, ~14 KB, is not used)
  const unsigned int fsm_addr = smem_u32(fsm);
  unsigned char* table = fsm + (FS_TABLE_ADDR - fsm_addr);                           // [M/4][128 lines][256 B]
  unsigned int* s_hist = reinterpret_cast<unsigned int*>(table + (size_t)M * 8192);  // [FS_BINS]
  __shared__ unsigned int s_wsum[32];
  __shared__ unsigned int s_T, s_cnt, s_cnt2, s_last, s_nsurv;
  __shared__ long long s_next[2];
  phase_stamp(0);

  // ---- byte image of the LUT: q = floor((v - min_s) / delta), delta = largest subspace range / 255 ------
  // (entries c >= K stand for the 0 that pqApproxDistance adds for codes outside the codebook, :1061)
  __shared__ double s_mn[16];
  __shared__ double s_delta;
  if (tid == 0) {
    double range = 0.0, amax = 0.0;
    bool bad = false;
    for (int s = 0; s < M; s++) {
      const unsigned long long lo = MM[((size_t)qi * M + s) * 2], hi = MM[((size_t)qi * M + s) * 2 + 1];
      bad |= hi == ~0ull;
      double mn = f64_from_ordered(lo), mx = f64_from_ordered(hi);
      if (K < 256) {
        mn = fmin(mn, 0.0);
        mx = fmax(mx, 0.0);
      }
      s_mn[s] = mn;
      range = fmax(range, mx - mn);
      amax = fmax(amax, fmax(fabs(mn), fabs(mx)));
    }
    const double delta = range / 255.0;
    s_delta = (bad || !(delta > 0.0) || !(delta > amax * 1e-9)) ? 0.0 : delta;
  }
  __syncthreads();
  if (!(s_delta > 0.0) || fsm_addr > FS_TABLE_ADDR) {
    // no usable byte image of this table (or, never seen, static shared memory reaching beyond the table's fixed address):
    // adc_fallback_kernel evaluates every row exactly
    if (blockIdx.x == 0 && tid == 0) atomicExch(ctrl + FS_FLAG, 1u);
    return;
  }
  // ---- replicate the byte table: lane l's copy lives entirely in bank l ------------------------------
  {
    // word (s, hbit, line) of the byte table holds the four codes c = hbit*128 + line*4 + {0..3}; it goes to
    // fs_entry_offset(s, c) + l*4 = (s>>2)*32768 + (line*4 + (s&3))*256 + hbit*128 + l*4 for every lane l
    unsigned char* qtab = reinterpret_cast<unsigned char*>(s_hist);  // staging; the histogram is zeroed below
    const double inv_delta = 1.0 / s_delta;
    double v[(M * 256 + FS_THREADS - 1) / FS_THREADS];
#pragma unroll
    for (int i = 0; i < (M * 256 + FS_THREADS - 1) / FS_THREADS; i++) {
      const int e = tid + i * FS_THREADS, s = e >> 8, c = e & 255;
      v[i] = (e < M * 256 && c < K) ? lut64[(size_t)s * K + c] : 0.0;
    }
#pragma unroll
    for (int i = 0; i < (M * 256 + FS_THREADS - 1) / FS_THREADS; i++) {
      const int e = tid + i * FS_THREADS;
      if (e < M * 256) {
        int qv = (int)((v[i] - s_mn[e >> 8]) * inv_delta);
        qtab[e] = (unsigned char)(qv < 0 ? 0 : (qv > 255 ? 255 : qv));
      }
    }
    __syncthreads();
    const unsigned int* q32 = reinterpret_cast<const unsigned int*>(qtab);
    unsigned int* t32 = reinterpret_cast<unsigned int*>(table);
    for (int wi = tid; wi < M * 2048; wi += FS_THREADS) {
      const int l = wi & 31, rest = wi >> 5, s = rest >> 6, hbit = (rest >> 5) & 1, line = rest & 31;
      t32[(s >> 2) * 8192 + (line * 4 + (s & 3)) * 64 + hbit * 32 + l] = q32[rest];
    }
    __syncthreads();
  }
  if (tid == 0) {
    s_T = FS_T_INF;
    s_cnt = 0u;
    s_cnt2 = 0u;
    s_nsurv = 0u;
  }
  for (int i = tid; i < FS_BINS; i += FS_THREADS) s_hist[i] = 0u;
  __syncthreads();
  phase_stamp(1);

  // The ALU pipe (LOP3 / SHF / PRMT / IADD3) is the scarce resource: three word-wide operations
  // split the four codes of a word into their high and low address bytes, ONE byte permute per
  // code assembles the offset.
  const unsigned int lane4x4 = (unsigned int)(lane * 4) * 0x01010101u;
  auto row_sum = [&](const uint32_t* w) -> unsigned int {
    unsigned int acc = fs_word_sum<0>(w[0], lane4x4) + fs_word_sum<1>(w[1], lane4x4);
    if (MW == 4) acc += fs_word_sum<2>(w[MW - 2], lane4x4) + fs_word_sum<3>(w[MW - 1], lane4x4);
    return acc;
  };
  auto load_row = [&](const uint8_t* p, uint32_t* w) {
    if (MW == 4) {
      const uint4 v = ld_stream_u4(reinterpret_cast<const uint4*>(p));
      w[0] = v.x; w[1] = v.y; w[2] = v.z; w[3] = v.w;
    } else {
      const uint2 v = __ldg(reinterpret_cast<const uint2*>(p));
      w[0] = v.x; w[1] = v.y;
    }
  };
  // drain this thread's bins of the CTA histogram into the global one (no barrier needed: the
  // exchange takes whatever has been counted so far, later counts go with the next drain)
  auto drain = [&]() {
    const uint4 v = *reinterpret_cast<const uint4*>(s_hist + tid * BPT);
    if (v.x | v.y | v.z | v.w) {
#pragma unroll
      for (int i = 0; i < BPT; i++) {
        const unsigned int c = atomicExch(&s_hist[tid * BPT + i], 0u);
        if (c) atomicAdd(hist_g + tid * BPT + i, c);
      }
    }
  };
  static_assert(BPT == 4, "drain reads one uint4 per thread");

  constexpr int U = FS_U;
  constexpr int RPB = FS_THREADS * U;  // rows per batch
  const int64_t nbatches = (n + RPB - 1) / RPB;

  // ---- warm-up: the first FS_WU * FS_THREADS rows of the CTA's first batch give a local threshold ---------
  {
#pragma unroll
    for (int u = 0; u < FS_WU; u++) {
      const int64_t r = (int64_t)blockIdx.x * RPB + u * FS_THREADS + tid;
      if (r < n) {
        uint32_t w[MW];
        load_row(codes + (size_t)r * M, w);
        atomicAdd(&s_hist[row_sum(w)], 1u);
      }
    }
    __syncthreads();
    unsigned int c4 = 0;
#pragma unroll
    for (int i = 0; i < BPT; i++) c4 += s_hist[tid * BPT + i];
    unsigned int incl = c4;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const unsigned int v = __shfl_up_sync(FULL_MASK, incl, o);
      if (lane >= o) incl += v;
    }
    if (lane == 31) s_wsum[warp] = incl;
    __syncthreads();
    unsigned int before = 0;
    for (int wj = 0; wj < warp; wj++) before += s_wsum[wj];
    const unsigned int excl = before + incl - c4;
    if (excl < k && excl + c4 >= k) {
      unsigned int acc = excl;
#pragma unroll
      for (int i = 0; i < BPT; i++) {
        acc += s_hist[tid * BPT + i];
        if (acc >= k) {
          const unsigned int T0 = tid * BPT + i + M + 2;
          s_T = T0;
          atomicMax(ctrl + FS_TINV, ~T0);
          break;
        }
      }
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < BPT; i++) s_hist[tid * BPT + i] = 0u;  // the loop below counts its own candidates
    __syncthreads();
  }
  unsigned int Tloc = s_T;
  phase_stamp(2);

  // ---- streaming loop --------------------------------------------------------------------------------
  // A thread holds U rows (64 bytes) in registers per batch: 64 KB per SM, which at the loaded HBM
  // latency (~2 us) caps the stream near 4.4 TB/s.  So every batch is requested into L2 FS_PF batches
  // ahead with prefetch.global.L2 (no register, no scoreboard), and the loads that follow are L2 hits.
  // The first 7/8 of the batches are split statically (grid-stride); the rest are handed out through
  // a global counter so that SMs that stream faster take more of the tail.
  int it = 0;
  unsigned int turn = 0;  // it mod gridDim.x
  auto prefetch = [&](const int64_t batch) {
    const int64_t row0 = batch * RPB + tid;
    if (row0 + (U - 1) * FS_THREADS < n) {
      const uint8_t* p0 = codes + (size_t)row0 * M;
#pragma unroll
      for (int u = 0; u < U; u++) asm volatile("prefetch.global.L2 [%0];" ::"l"(p0 + (size_t)u * FS_THREADS * M));
    }
  };
  auto process = [&](const int64_t batch) {
    const unsigned int tinv = ld_cg_u32(ctrl + FS_TINV);
    const int64_t row0 = batch * RPB + tid;
    const uint8_t* p0 = codes + (size_t)row0 * M;
    uint32_t w[U][MW];
    unsigned int S[U];
    bool skip_section = false;
    if (row0 + (U - 1) * FS_THREADS < n) {  // all of this thread's rows exist (every batch but the last)
#pragma unroll
      for (int u = 0; u < U; u++) load_row(p0 + (size_t)u * FS_THREADS * M, w[u]);
      drain();
      if (MW == 4) {
        // Table bytes are >= 0, so the sum over the first fourteen subspaces is a lower bound of S: when none of the
        // warp's 256 rows has it at or below the threshold (the threshold sits near the 1e-6 quantile of S at C4; twelve
        // subspaces were measured too weak a bound: four warp-batches in five had a row below it), the last two lookups
        // of every row and the candidate section are skipped.
        unsigned int pmin = 0xffffffffu;
        unsigned int lo3[U], hi3[U];
#pragma unroll
        for (int u = 0; u < U; u++) {
          hi3[u] = fs_hi(w[u][3]);
          lo3[u] = fs_lo(w[u][3], lane4x4);
          S[u] = fs_word_sum<0>(w[u][0], lane4x4) + fs_word_sum<1>(w[u][1], lane4x4) +
                 fs_word_sum<2>(w[u][2], lane4x4) + fs_half_word_sum<3, 0>(lo3[u], hi3[u]);
          pmin = min(pmin, S[u]);
        }
        Tloc = min(Tloc, ~tinv);
        if (!__any_sync(FULL_MASK, pmin <= Tloc)) {
          skip_section = true;
        } else {
#pragma unroll
          for (int u = 0; u < U; u++) S[u] += fs_half_word_sum<3, 1>(lo3[u], hi3[u]);
        }
      } else {
#pragma unroll
        for (int u = 0; u < U; u++) S[u] = row_sum(w[u]);
      }
    } else {
      drain();
#pragma unroll
      for (int u = 0; u < U; u++) {
        S[u] = 0xffffffffu;  // above every threshold (FS_T_INF is 0xfffffffe)
        if (row0 + u * FS_THREADS < n) {
          load_row(p0 + (size_t)u * FS_THREADS * M, w[u]);
          S[u] = row_sum(w[u]);
        }
      }
    }
    Tloc = min(Tloc, ~tinv);
#pragma unroll
    for (int u = 0; u < U; u++) {
      if (skip_section) break;  // (warp-uniform)
      const bool pred = S[u] <= Tloc;
      const unsigned m = __ballot_sync(FULL_MASK, pred);
      if (m) {
        const int leader = __ffs(m) - 1;
        unsigned int base = 0;
        if (lane == leader) base = atomicAdd(&s_cnt, (unsigned int)__popc(m));
        base = __shfl_sync(FULL_MASK, base, leader);
        if (pred) {
          const unsigned int idx = base + __popc(m & ((1u << lane) - 1u));
          if (idx < cap)
            cand[idx] = ((unsigned long long)S[u] << 48) | (unsigned long long)(row0 + u * FS_THREADS);
          else
            atomicExch(ctrl + FS_FLAG, 1u);
          atomicAdd(&s_hist[S[u]], 1u);
        }
      }
    }
    // re-derive the device-wide threshold from the global histogram: every CTA after its batches
    // 1, 2, 4, 8, ... (rotating the warp), and one CTA in turn after every other batch
    const bool pow2 = it > 0 && (it & (it - 1)) == 0;
    const bool my_turn = pow2 ? warp == (31 - __clz(it)) : (turn == blockIdx.x && warp == (it & 31));
    if (++turn == gridDim.x) turn = 0;
    if (my_turn) {
      const unsigned int Tcur = min(Tloc, ~ld_cg_u32(ctrl + FS_TINV));
      const unsigned int nb = min(Tcur, (unsigned int)(FS_BINS - 1)) + 1u;
      const unsigned int b = fs_warp_kth(hist_g, nb, k, lane);
      if (b != FS_T_INF) {
        const unsigned int Tn = min(Tcur, b + M + 2);
        if (lane == 0) atomicMax(ctrl + FS_TINV, ~Tn);
        Tloc = min(Tloc, Tn);
      }
    }
    ++it;
  };
  const int64_t nstatic = (nbatches * 7 / 8) / gridDim.x * gridDim.x;
  for (int p = 1; p < FS_PF; p++) prefetch((int64_t)blockIdx.x + (int64_t)p * gridDim.x);
  for (int64_t batch = blockIdx.x; batch < nstatic; batch += gridDim.x) {
    if (batch + (int64_t)FS_PF * gridDim.x < nstatic) prefetch(batch + (int64_t)FS_PF * gridDim.x);
    process(batch);
  }
  if (tid == 0) s_next[0] = nstatic + (int64_t)atomicAdd(ctrl + FS_NEXT, 1u);
  for (int j = 0;; j++) {
    __syncthreads();  // s_next[j & 1] is visible; nobody still reads the other slot
    const int64_t batch = s_next[j & 1];
    if (tid == 0) s_next[(j + 1) & 1] = nstatic + (int64_t)atomicAdd(ctrl + FS_NEXT, 1u);  // latency hides behind the batch
    if (batch >= nbatches) break;
    process(batch);
  }

  // ---- compact and publish this CTA's list; the last CTA to arrive selects -----------------------------
  __syncthreads();
  phase_stamp(3);
  drain();
  {
    const unsigned int Tg = min(Tloc, ~ld_cg_u32(ctrl + FS_TINV));
    const unsigned int cnt = min(s_cnt, cap);
    constexpr int EPT = 8;  // entries per thread and round
    for (unsigned int base = 0; base < cnt; base += FS_THREADS * EPT) {  // one round unless cap > 8192
      unsigned long long e[EPT];
#pragma unroll
      for (int j = 0; j < EPT; j++) {
        const unsigned int i = base + j * FS_THREADS + tid;
        e[j] = i < cnt ? __ldcg(cand + i) : ~0ull;  // S = 0xffff > 255 * 16 marks "no entry"
      }
      __syncthreads();  // every entry of this round is in registers before slots are overwritten
#pragma unroll
      for (int j = 0; j < EPT; j++)
        if ((unsigned int)(e[j] >> 48) <= min(Tg, (unsigned int)(FS_BINS - 1))) cand[atomicAdd(&s_cnt2, 1u)] = e[j];
    }
  }
  __threadfence();
  __syncthreads();
  if (tid == 0) {
    counts_g[blockIdx.x] = s_cnt2;
    __threadfence();
    s_last = (atomicAdd(ctrl + FS_TICKET, 1u) == gridDim.x - 1) ? 1u : 0u;
  }
  __syncthreads();
  phase_stamp(4);
  if (!s_last) return;
  __threadfence();

  const bool flagged = ld_cg_u32(ctrl + FS_FLAG) != 0u;
  ulonglong2* keys = reinterpret_cast<ulonglong2*>(table);  // the table is dead: M*512 sortable keys
  unsigned int* s_off = s_hist;                            // [gridDim.x + 1] list offsets
  const unsigned int sort_cap = M * 512;
  if (!flagged) {
    if (warp == 0) {
      const unsigned int Tg = ~ld_cg_u32(ctrl + FS_TINV);
      const unsigned int b = fs_warp_kth(hist_g, min(Tg, (unsigned int)(FS_BINS - 1)) + 1u, k, lane);
      if (lane == 0) s_T = (b == FS_T_INF) ? FS_T_INF : min(Tg, b + (unsigned int)M + 2u);
    } else if (warp == 1) {
      // exclusive prefix of the per-CTA list lengths (gridDim.x <= FS_MAX_GRID = 8 * 32)
      unsigned int c[FS_MAX_GRID / 32], tot = 0;
#pragma unroll
      for (int j = 0; j < FS_MAX_GRID / 32; j++) {
        const unsigned int ci = lane * (FS_MAX_GRID / 32) + j;
        c[j] = ci < gridDim.x ? ld_cg_u32(counts_g + ci) : 0u;
        tot += c[j];
      }
      unsigned int incl = tot;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const unsigned int v = __shfl_up_sync(FULL_MASK, incl, o);
        if (lane >= o) incl += v;
      }
      unsigned int run = incl - tot;
#pragma unroll
      for (int j = 0; j < FS_MAX_GRID / 32; j++) {
        const unsigned int ci = lane * (FS_MAX_GRID / 32) + j;
        if (ci < gridDim.x) s_off[ci] = run;
        run += c[j];
      }
      if (lane == 31) s_off[gridDim.x] = incl;
    }
    __syncthreads();
    const unsigned int Tf = s_T;
    const unsigned int total = s_off[gridDim.x];
    phase_stamp(5);
    for (unsigned int i = tid; i < total; i += FS_THREADS) {
      unsigned int lo = 0, hi = gridDim.x;  // the list with s_off[c] <= i < s_off[c + 1]
      while (hi - lo > 1) {
        const unsigned int mid = (lo + hi) >> 1;
        if (s_off[mid] <= i) lo = mid; else hi = mid;
      }
      const unsigned long long e = __ldcg(cand_q + (size_t)lo * cap + (i - s_off[lo]));
      if ((unsigned int)(e >> 48) <= Tf) {
        const int64_t row = (int64_t)(e & 0xffffffffffffull);
        const unsigned int slot = atomicAdd(&s_nsurv, 1u);
        if (slot < sort_cap)
          st_key(keys + slot, Key{rank_hi_from_dist(fs_exact(lut64, codes + (size_t)row * M, M, K)), (uint64_t)row});
      }
    }
    __syncthreads();
    const unsigned int ns = s_nsurv;
    phase_stamp(6);
    if (qi == 0 && tid == 0) {
      g_adc_dbg[0] += total;
      g_adc_dbg[1] = Tf;
      g_adc_dbg[2] = ns;
    }
    if (ns > sort_cap) {
      if (tid == 0) atomicExch(ctrl + FS_FLAG, 1u);
    } else {
      if (ns <= FS_THREADS) {
        // few survivors (the usual case): rank by counting, no sort.  Keys are unique (row ids).
        if (tid < ns) {
          const Key me = ld_key(keys + tid);
          unsigned int rank = 0;
          for (unsigned int j = 0; j < ns; j++) rank += key_lt(ld_key(keys + j), me) ? 1u : 0u;
          if (rank < k) {
            ids_out[(size_t)qi * out_stride + rank] = id_base + (int64_t)me.lo;
            approx_out[(size_t)qi * out_stride + rank] = dist_from_rank_hi(me.hi);
          }
        }
        for (unsigned int i = ns + tid; i < k; i += FS_THREADS) {
          ids_out[(size_t)qi * out_stride + i] = -1;
          approx_out[(size_t)qi * out_stride + i] = __longlong_as_double(0x7ff8000000000000ll);
        }
      } else {
        unsigned int np = 2;
        while (np < ns) np <<= 1;
        for (unsigned int i = ns + tid; i < np; i += FS_THREADS) st_key(keys + i, key_empty());
        __syncthreads();
        for (unsigned int size = 2; size <= np; size <<= 1) {
          for (unsigned int stride = size >> 1; stride > 0; stride >>= 1) {
            for (unsigned int t = tid; t < (np >> 1); t += FS_THREADS) {
              const unsigned int i = ((t & ~(stride - 1)) << 1) | (t & (stride - 1));
              cswap(keys, (int)i, (int)(i | stride), (i & size) == 0);
            }
            __syncthreads();
          }
        }
        for (unsigned int i = tid; i < k; i += FS_THREADS) {
          const bool ok = i < ns;
          const Key e = ok ? ld_key(keys + i) : key_empty();
          ids_out[(size_t)qi * out_stride + i] = ok ? id_base + (int64_t)e.lo : -1;
          approx_out[(size_t)qi * out_stride + i] = ok ? dist_from_rank_hi(e.hi) : __longlong_as_double(0x7ff8000000000000ll);
        }
      }
      if (tid == 0) counts_out[qi] = (int32_t)min(ns, k);
    }
  }
  // leave the scratch zeroed for the next launch (the fallback flag is cleared by its consumer)
  __syncthreads();
  for (int i = tid; i < FS_WORDS; i += FS_THREADS)
    if (i != FS_BINS + FS_FLAG) hist_g[i] = 0u;
  phase_stamp(7);
}

// Exact evaluation of every row (pqApproxDistance :1057-1065) + top-k; runs only when the fast scan
// raised its flag (degenerate table, list overflow), otherwise exits at once.
template <class TK>
__global__ void __launch_bounds__(SCAN_THREADS)
adc_fallback_kernel(const uint8_t* __restrict__ codes, int64_t n, int M, int K, const double* __restrict__ LUT64,
                    unsigned int* __restrict__ fs, int k, int kp, TopkOut out) {
  extern __shared__ __align__(128) ulonglong2 smem[];
  const int qi = blockIdx.y;
  unsigned int* ctrl = fs + (size_t)qi * FS_WORDS + FS_BINS;
  // launched with programmatic stream serialisation: the launch itself overlaps the scan kernel's tail, this
  // waits for the scan (and its flag) to be complete and visible
  pdl_trigger();  // the re-rank kernel may be placed
  pdl_wait();
  if (ld_cg_u32(ctrl + FS_FLAG) == 0u) return;
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int nw = blockDim.x >> 5;
  const int stride_keys = kp + TOPK_BUF;
  const double* __restrict__ lut64 = LUT64 + (size_t)qi * M * K;
  TK tk;
  tk.init(smem + (size_t)warp * stride_keys, kp, k, lane);
  __syncthreads();
  const int64_t nb = (n + 31) / 32;
  for (int64_t b = (int64_t)blockIdx.x * nw + warp; b < nb; b += (int64_t)gridDim.x * nw) {
    const int64_t row = b * 32 + lane;
    Key key = key_empty();
    if (row < n) key = Key{rank_hi_from_dist(fs_exact(lut64, codes + (size_t)row * M, M, K)), (uint64_t)row};
    tk.push(key, row < n, lane);
  }
  if (qi == 0 && blockIdx.x == 0 && threadIdx.x == 0) g_adc_dbg[3] += 1u;
  const bool last = topk_epilogue(tk, smem, kp, k, out);
  if (last && threadIdx.x == 0) ctrl[FS_FLAG] = 0u;
}

#ifdef VS_PHASE_STAMPS
int debug_read_stamps_adc(void* dst, size_t bytes) {
  cudaDeviceSynchronize();
  return (int)cudaMemcpyFromSymbol(dst, g_phase_stamps, bytes);
}
#endif

int debug_adc_stats(unsigned int* out) {
  cudaDeviceSynchronize();
  cudaError_t e = cudaMemcpyFromSymbol(out, g_adc_dbg, sizeof(unsigned int) * 8);
  unsigned int zero[8] = {0};
  if (e == cudaSuccess) e = cudaMemcpyToSymbol(g_adc_dbg, zero, sizeof zero);
  return (int)e;
}

// ---- host -------------------------------------------------------------------------------------------
bool adc_fast_supported(int M, int K) { return (M == 8 || M == 16) && K >= 1 && K <= 256; }

// (the table sits at window address FS_TABLE_ADDR whatever lies before the dynamic region: this bound holds for any)
static size_t fastscan_smem(int M) { return (size_t)FS_TABLE_ADDR + (size_t)M * 8192 + FS_BINS * 4; }

typedef void (*FallbackKern)(const uint8_t*, int64_t, int, int, const double*, unsigned int*, int, int, TopkOut);

bool adc_fast_configure(AdcFastLaunch& L, int sms) {
  L.kp = topk_pad(L.k);
  const int64_t rows_per_batch = (int64_t)FS_THREADS * FS_U;
  const int64_t nbatches = (L.n + rows_per_batch - 1) / rows_per_batch;
  int64_t grid = nbatches < sms ? nbatches : sms;
  if (grid > FS_MAX_GRID) grid = FS_MAX_GRID;
  L.grid = (int)(grid < 1 ? 1 : grid);
  L.smem_bytes = fastscan_smem(L.M);
  cudaError_t e = L.M == 16
                      ? cudaFuncSetAttribute(adc_fastscan_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L.smem_bytes)
                      : cudaFuncSetAttribute(adc_fastscan_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L.smem_bytes);
  if (e != cudaSuccess) return false;
  L.final_threads = SCAN_THREADS;
  L.final_grid = sms < TOPK_MAX_LISTS ? sms : TOPK_MAX_LISTS;
  L.final_smem = topk_block_smem(L.k, L.kp, L.final_threads / 32);
  FallbackKern fk = L.k <= TOPK_REG_MAX_K ? adc_fallback_kernel<WarpTopKReg> : adc_fallback_kernel<WarpTopK>;
  // the attribute is per kernel and plans are cached: opt in to the largest size any plan can ask for
  // (k = TOPK_MAX_K), so that a later plan with a smaller k cannot lower it under an earlier one
  e = cudaFuncSetAttribute(fk, cudaFuncAttributeMaxDynamicSharedMemorySize,
                           (int)topk_block_smem(TOPK_MAX_K, topk_pad(TOPK_MAX_K), L.final_threads / 32));
  if (e != cudaSuccess) return false;
  L.partial_keys = topk_partial_keys(L.k, L.final_grid, L.final_threads / 32);
  return true;
}

cudaError_t launch_build_lut_mm(const float* centroids, int M, int K, int subDim, const float* q, int nq, int lanes,
                                double* lut64, unsigned long long* mm, cudaStream_t st) {
  build_lut_mm_kernel<<<dim3(M, nq), 256, 0, st>>>(centroids, M, K, subDim, q, lanes, lut64, mm);
  count_launch();
  return cudaGetLastError();
}

cudaError_t launch_adc_fast(const AdcFastLaunch& L, cudaStream_t st) {
  // (the kernel sizes everything from gridDim.x; the lists and counters were sized for L.grid CTAs)
  const dim3 grid(L.reserve_sms > 0 && L.grid - L.reserve_sms >= 1 ? L.grid - L.reserve_sms : L.grid, L.nq);
  const int64_t stride = L.out_stride > 0 ? L.out_stride : L.k;
  count_launch();
  if (L.M == 16)
    return launch_pdl(adc_fastscan_kernel<4>, grid, dim3(FS_THREADS), L.smem_bytes, st, L.codes, L.n, L.K, L.lut64, L.mm,
                      (unsigned int)L.k, L.fs, L.cand, L.cap, L.ids_out, L.approx_out, L.counts_out, L.id_base, stride);
  return launch_pdl(adc_fastscan_kernel<2>, grid, dim3(FS_THREADS), L.smem_bytes, st, L.codes, L.n, L.K, L.lut64, L.mm,
                    (unsigned int)L.k, L.fs, L.cand, L.cap, L.ids_out, L.approx_out, L.counts_out, L.id_base, stride);
}
cudaError_t launch_adc_fallback(const AdcFastLaunch& L, cudaStream_t st) {
  TopkOut o{L.partial, L.ctrl, L.partial_keys, L.ids_out, L.approx_out, L.counts_out, L.id_base, 1,
            L.out_stride > 0 ? L.out_stride : L.k};
  FallbackKern fk = L.k <= TOPK_REG_MAX_K ? adc_fallback_kernel<WarpTopKReg> : adc_fallback_kernel<WarpTopK>;
  count_launch();
  return launch_pdl(fk, dim3(L.final_grid, L.nq), dim3(L.final_threads), L.final_smem, st, L.codes, L.n, L.M, L.K, L.lut64,
                    L.fs, L.k, L.kp, o);
}

}  // namespace vs
