// pq_tc.cu -- K3 on the 5th-generation tensor cores: PQ assignment (PqEncoder.encode, J/pq/PqEncoder.java:18-37, and
// the assignment step of PqTrainer.train, J/pq/PqTrainer.java:56-68) for the production shape: 8-float sub-vectors
// (d = 8 M), K <= 256 centroids per subspace.
//
// Per subspace the assignment is [rows x 8] x [8 x K]: a GEMM with an inner dimension of 8.  The reference decides
// every argmin in fp64 (the sub-vector is shorter than the SIMD register, J/util/Distances.java:77-94), so the
// tensor cores only NOMINATE, with enough precision that a second look is rare:
//   * operands are fp16 pairs: x s = xh + xl, c s = ch + cl (s a power of two that brings both into [-64, 64]), and
//     32 halfs per (row, subspace) / (centroid, subspace) -- two subspaces share a 128-byte operand row -- carry
//         A = [ xh | xl | xh | 1 1 0 .. ]      B = [ ch | ch | cl | nh nl 0 .. ],   nh + nl = -|c s|^2 / 2
//     so two tcgen05.mma.kind::f16 (K = 16 each) leave  D = s^2 (<x,c> - |c|^2 / 2)  in TMEM, good to ~2^-21:
//     argmax_c D = argmin_c |x - c|^2, no coefficient to apply in the epilogue.
//   * pq_tc_assign_kernel: persistent, warp-specialised like batch_gemm_kernel.  A CTA keeps the B blocks of 4
//     subspaces resident (64 KB), TMA streams the A blocks of its row tiles (128 rows x 128 bytes per subspace pair), one
//     thread issues the MMAs into a double-buffered 128 x 256 accumulator, eight epilogue warps (thread = row, two
//     warps share a row's 256 columns) read it back twice: once for the largest key (D with the low 6 mantissa bits
//     replaced by a column slot; ALU pipe), once to count the columns inside the error band of that maximum with
//     saturating multiply-adds (FMA pipe).  If nothing else is inside the band the winner IS the reference argmin; otherwise
//     the warp re-reads the accumulator and the centroids inside the band are decided in the reference's own
//     arithmetic, strict '<' in ascending index -- bit-identical codes, like pq.cu.
//   * tq_prep_*: one pass that writes the operand images (64 bytes per row and subspace; the scale and the
//     norm bounds of the error band stay on the device, nothing synchronises).
// The kernel is bound by its epilogue (3.5 instructions per (row, centroid): 1.5 on the half-rate ALU pipe, 2 on the FMA
// pipe), not by the MMA.
#include <algorithm>
#include <atomic>
#include <mutex>

#include "kernels.h"
#include "tc05.cuh"

namespace vs {

namespace {

constexpr int TQ_M = 128;          // rows per tile = TMEM lanes
constexpr int TQ_N = 256;          // centroids per subspace (padded) = accumulator columns
constexpr int TQ_SD = 8;           // floats per sub-vector
constexpr int TQ_SPC = 4;          // subspaces whose B blocks a CTA keeps resident
constexpr int TQ_STAGES = 4;
constexpr int TQ_THREADS = 384;    // warp 0 TMA, warp 1 MMA, warp 2 TMEM allocator, warps 4..11 epilogue
constexpr uint32_t TQ_A_BYTES = TQ_M * 128;
constexpr uint32_t TQ_B_BYTES = TQ_N * 128;
constexpr int TQ_PPC = TQ_SPC / 2;  // subspace pairs per CTA: a 128-byte operand row carries two subspaces
constexpr size_t TQ_XCH_BYTES = 2 * 4 * 32 * 16;  // hand-over of the second epilogue warp's two best keys, double-buffered
constexpr size_t TQ_SMEM = (size_t)TQ_PPC * TQ_B_BYTES + (size_t)TQ_STAGES * TQ_A_BYTES + TQ_XCH_BYTES + 256;
constexpr uint32_t TQ_IDESC = (1u << 4) | (0u << 7) | (0u << 10) | ((uint32_t)(TQ_N >> 3) << 17) | ((uint32_t)(TQ_M >> 4) << 24);

struct TqStats {            // device-resident, written by the prep kernels
  unsigned int amax_bits;   // max |element| over rows and centroids (fixes the scale)
  unsigned int xxmax_bits;  // max |sub-vector|^2 over the rows
  unsigned int nmax_bits;   // max |centroid|^2
  int pad;
};

__device__ __forceinline__ float tq_scale(const TqStats* st) {
  const float amax = __uint_as_float(st->amax_bits);
  if (!(amax > 0.0f) || !(amax < 1e30f)) return 1.0f;
  int e;
  frexpf(amax, &e);            // amax <= 2^e
  return ldexpf(1.0f, 6 - e);  // |x s| <= 64
}

// max |element| of a float array (finite values only)
__global__ void __launch_bounds__(256) tq_amax_kernel(const float* __restrict__ v, int64_t count, TqStats* __restrict__ st) {
  float m = 0.0f;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += (int64_t)gridDim.x * blockDim.x) {
    const float a = fabsf(v[i]);
    if (a < 1e30f) m = fmaxf(m, a);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(FULL_MASK, m, o));
  if ((threadIdx.x & 31) == 0 && m > 0.0f) atomicMax(&st->amax_bits, __float_as_uint(m));
}

__device__ __forceinline__ void tq_split(float v, float s, __half& hi, __half& lo) {
  const float t = v * s;  // exact: s is a power of two
  hi = __float2half_rn(t);
  lo = __float2half_rn(t - __half2float(hi));
}

// operand image of the rows: Xop[r][p] = 128 bytes for the subspace PAIR p = (2p, 2p + 1):
//   {xh[8], xl[8], xh[8], 1, 1, 0 x 6} of subspace 2p, then the same of subspace 2p + 1.
// Eight threads share a (row, pair) and write one 16-byte chunk each, so a warp stores 512 contiguous bytes.
__global__ void __launch_bounds__(256)
tq_prep_rows_kernel(const float* __restrict__ X, int64_t n, int d, int M, __half* __restrict__ Xop, TqStats* __restrict__ st) {
  const float s = tq_scale(st);
  float xxm = 0.0f;
  const int P = M >> 1;
  const int64_t total = n * P * 8;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
    const int chunk = (int)(t & 7);
    const int64_t i = t >> 3;  // (row, pair)
    const int64_t r = i / P;
    const int p = (int)(i - r * P);
    const int sub = 2 * p + (chunk >> 2);
    const float4* src = reinterpret_cast<const float4*>(X + (size_t)r * d + (size_t)sub * TQ_SD);
    const float4 a = __ldg(src), b = __ldg(src + 1);
    const float x[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
    __align__(16) __half h[8];
    const int kind = chunk & 3;  // 0: xh, 1: xl, 2: xh, 3: ones
    if (kind == 3) {
#pragma unroll
      for (int j = 0; j < 8; j++) h[j] = __float2half_rn(j < 2 ? 1.0f : 0.0f);
      float xx = 0.0f;
#pragma unroll
      for (int j = 0; j < 8; j++) xx = fmaf(x[j], x[j], xx);
      if (xx < 1e30f) xxm = fmaxf(xxm, xx);
    } else {
#pragma unroll
      for (int j = 0; j < 8; j++) {
        __half hi, lo;
        tq_split(x[j], s, hi, lo);
        h[j] = kind == 1 ? lo : hi;
      }
    }
    reinterpret_cast<uint4*>(Xop)[t] = *reinterpret_cast<const uint4*>(h);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) xxm = fmaxf(xxm, __shfl_xor_sync(FULL_MASK, xxm, o));
  if ((threadIdx.x & 31) == 0 && xxm > 0.0f) atomicMax(&st->xxmax_bits, __float_as_uint(xxm));
}

// operand image of the centroids: Cop[p][c] = 128 bytes {ch[8], ch[8], cl[8], nh, nl, 0 x 6} of subspace 2p, then
// the same of subspace 2p + 1; rows c >= K (padding up to 256) get the most negative norm term, so they never win
__global__ void __launch_bounds__(256)
tq_prep_centroids_kernel(const float* __restrict__ C, int M, int K, __half* __restrict__ Cop, TqStats* __restrict__ st) {
  const float s = tq_scale(st);
  const int i = blockIdx.x * blockDim.x + threadIdx.x;  // (subspace, centroid)
  if (i >= M * TQ_N) return;
  const int sub = i / TQ_N, c = i % TQ_N;
  __align__(16) __half h[32];
#pragma unroll
  for (int j = 0; j < 32; j++) h[j] = __float2half_rn(0.0f);
  if (c < K) {
    const float* src = C + ((size_t)sub * K + c) * TQ_SD;
    float nn = 0.0f, ns = 0.0f;
#pragma unroll
    for (int j = 0; j < 8; j++) {
      __half hi, lo;
      tq_split(src[j], s, hi, lo);
      h[j] = hi;
      h[8 + j] = hi;
      h[16 + j] = lo;
      const float t = src[j] * s;
      ns = fmaf(t, t, ns);
      nn = fmaf(src[j], src[j], nn);
    }
    const float nt = -0.5f * ns;
    const __half nh = __float2half_rn(nt);
    h[24] = nh;
    h[25] = __float2half_rn(nt - __half2float(nh));
    if (nn < 1e30f && nn > 0.0f) atomicMax(&st->nmax_bits, __float_as_uint(nn));
  } else {
    h[24] = __float2half_rn(-60000.0f);
  }
  uint4* dst = reinterpret_cast<uint4*>(Cop + ((size_t)(sub >> 1) * TQ_N + c) * 64 + (size_t)(sub & 1) * 32);
#pragma unroll
  for (int j = 0; j < 4; j++) dst[j] = reinterpret_cast<const uint4*>(h)[j];
}

// grid.x = (M / TQ_SPC) * nsplit: CTA (sg, split) owns subspaces sg*4 .. sg*4+3 and row tiles split, split + nsplit, ..
__global__ void __launch_bounds__(TQ_THREADS, 1)
pq_tc_assign_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                    const float* __restrict__ X, int64_t n, int d, int M, int K, const float* __restrict__ centroids, int lanes,
                    const TqStats* stats, int64_t row0 /* first row of this slab */, int64_t nslab, int nsg, int64_t tiles,
                    uint8_t* __restrict__ codes_u8, int32_t* __restrict__ assign_i32,
                    unsigned long long* __restrict__ dlist /* [nsg][dcap]: (row << 8 | subspace) of undecided entries */,
                    unsigned int* __restrict__ dcount /* [nsg] */, unsigned int dcap) {
  extern __shared__ __align__(1024) uint8_t tq_smem[];
  uint8_t* bq = tq_smem;                                          // [TQ_PPC][32 KB] centroid blocks (two subspaces each)
  uint8_t* stages = bq + (size_t)TQ_PPC * TQ_B_BYTES;             // [TQ_STAGES][16 KB] row blocks (two subspaces each)
  float4* sm_x = reinterpret_cast<float4*>(stages + (size_t)TQ_STAGES * TQ_A_BYTES);
  uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(sm_x) + TQ_XCH_BYTES);
  uint64_t* full = bars;                    // [TQ_STAGES]
  uint64_t* empty = bars + TQ_STAGES;       // [TQ_STAGES]
  uint64_t* tfull = bars + 2 * TQ_STAGES;   // [2]
  uint64_t* tempty = bars + 2 * TQ_STAGES + 2;
  uint64_t* bfull = bars + 2 * TQ_STAGES + 4;
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(bars + 2 * TQ_STAGES + 5);
  if ((smem_u32(tq_smem) & 1023u) != 0) __trap();

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int sg = blockIdx.x % nsg;
  const int split = blockIdx.x / nsg;
  const int nsplit = gridDim.x / nsg;
  const int s_first = sg * TQ_SPC;
  const int ns = min(TQ_SPC, M - s_first);  // even: M is even
  const int np = ns >> 1;

  if (threadIdx.x == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmA)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmB)) : "memory");
    for (int s = 0; s < TQ_STAGES; s++) {
      mbar_init(full + s, 1);
      mbar_init(empty + s, 1);
    }
    for (int a = 0; a < 2; a++) {
      mbar_init(tfull + a, 1);
      mbar_init(tempty + a, 8);
    }
    mbar_init(bfull, 1);
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
      mbar_expect_tx(bfull, (uint32_t)np * TQ_B_BYTES);
      for (int pl = 0; pl < np; pl++) tma_load_2d(bq + (size_t)pl * TQ_B_BYTES, &tmB, 0, ((s_first >> 1) + pl) * TQ_N, bfull);
      int s = 0;
      uint32_t ph = 0;
      for (int64_t tile = split; tile < tiles; tile += nsplit) {
        for (int pl = 0; pl < np; pl++) {
          mbar_wait_suspend(empty + s, ph ^ 1);
          mbar_expect_tx(full + s, TQ_A_BYTES);
          tma_load_2d(stages + (size_t)s * TQ_A_BYTES, &tmA, ((s_first >> 1) + pl) * 64, (int)(tile * TQ_M), full + s);
          if (++s == TQ_STAGES) {
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
      mbar_wait_suspend(bfull, 0);
      for (int64_t tile = split; tile < tiles; tile += nsplit) {
        for (int pl = 0; pl < np; pl++) {
          mbar_wait_suspend(full + s, ph);
          const uint64_t adesc = umma_desc_sw128(smem_u32(stages + (size_t)s * TQ_A_BYTES));
          const uint64_t bdesc = umma_desc_sw128(smem_u32(bq + (size_t)pl * TQ_B_BYTES));
          for (int e = 0; e < 2; e++, it++) {  // the two subspaces of the pair: halfs 0..31 and 32..63 of the operand rows
            const int a = it & 1;
            mbar_wait_suspend(tempty + a, ((it >> 1) & 1) ^ 1);
            tc_fence_after();
            const uint32_t d_tmem = tmem_base + (uint32_t)a * TQ_N;
            tc_mma_f16(d_tmem, adesc + 4 * e, bdesc + 4 * e, TQ_IDESC, 0u);          // xh.ch + xl.ch
            tc_mma_f16(d_tmem, adesc + 4 * e + 2, bdesc + 4 * e + 2, TQ_IDESC, 1u);  // xh.cl + 1.nh + 1.nl
            tc_commit(tfull + a);
          }
          tc_commit(empty + s);
          if (++s == TQ_STAGES) {
            s = 0;
            ph ^= 1;
          }
        }
      }
    }
  } else if (warp >= 4) {  // ===== epilogue: thread = row, two warps per TMEM lane quarter =====
    // Warps 4..7 (ch = 0) and 8..11 (ch = 1) reduce columns 0..127 and 128..255 of the same 32 rows; the second
    // warp hands its two best keys over through shared memory and the first one decides.  Two warps per SM
    // sub-partition hide the dependent min / max chains and the TMEM round trips of each other.
    const int ew = warp - 4;
    const int lq = ew & 3;
    const int ch = ew >> 2;
    const float ninf = __int_as_float(0xff800000);
    const float s = tq_scale(stats);
    // |estimate - reference| <= band (in distance units), see pq.cu; in D units: D = s^2 (|x|^2 - dist) / 2
    const float band = (__uint_as_float(stats->xxmax_bits) + __uint_as_float(stats->nmax_bits)) * (1.0f / 32768.0f) + 1e-30f;
    const float bandD = band * s * s * 0.5f;
    int it = 0;
    for (int64_t tile = split; tile < tiles; tile += nsplit) {
      const int64_t lrow = tile * TQ_M + lq * 32 + lane;  // row within the slab
      const bool live = lrow < nslab;
      const int64_t row = row0 + lrow;
      for (int sl = 0; sl < ns; sl++, it++) {
        const int a = it & 1;
        const int sub = s_first + sl;
        mbar_wait_suspend(tfull + a, (it >> 1) & 1);
        tc_fence_after();
        const uint32_t taddr = tmem_base + ((uint32_t)(lq * 32) << 16) + (uint32_t)(a * TQ_N);
        // Pass 1 (ALU pipe, 1.5 operations per column): the LARGEST key of this warp's 128 columns -- D with its low 6
        // mantissa bits replaced by a column slot -- in four independent chains, two columns per 3-input maximum
        // (column j -> chain j & 3, slot j >> 2).
        float m1[4];
#pragma unroll
        for (int c = 0; c < 4; c++) m1[c] = ninf;
        uint32_t v[2][32];
        tc_ld32(taddr + (uint32_t)ch * 128, v[0]);
#pragma unroll
        for (int h = 0; h < 4; h++) {
          tc_wait_ld();
          tc_pin32(v[h & 1]);
          if (h < 3) tc_ld32(taddr + (uint32_t)(ch * 128 + (h + 1) * 32), v[(h + 1) & 1]);
#pragma unroll
          for (int p = 0; p < 4; p++) {
#pragma unroll
            for (int c = 0; c < 4; c++) {
              const int ja = 8 * p + c, jb = ja + 4;  // columns of this load; the chain index ignores ch and h
              const int ca = h * 32 + ja, cb = h * 32 + jb;
              const float ka = __uint_as_float((v[h & 1][ja] & 0xffffffc0u) | (uint32_t)(ch * 32 + (ca >> 2)));
              const float kb = __uint_as_float((v[h & 1][jb] & 0xffffffc0u) | (uint32_t)(ch * 32 + (cb >> 2)));
              m1[c] = fmaxf(fmaxf(m1[c], ka), kb);
            }
          }
        }
        float b1 = m1[0];
        int bc = 0;
#pragma unroll
        for (int c = 1; c < 4; c++) {
          if (m1[c] > b1) {
            b1 = m1[c];
            bc = c;
          }
        }
        // Pass 2 (FMA pipe, which pass 1 leaves idle): is any OTHER column of this half within the error band of that
        // maximum?  The accumulator is read once more and every column adds sat((D - (v - 4 band)) / (2 band)): 1 for
        // the maximum itself and for every column within 2 band of it (the band the decision needs), a fraction for
        // columns between 2 and 4 band, 0 below and for NaN -- so a sum above 1.5 means "doubt" and can only err on
        // the safe side.  The runner-up VALUE is no longer tracked: 3.5 -> 1.5 half-rate operations per column.
        float cnt;
        {
          const float vl = __uint_as_float(__float_as_uint(b1) & 0xffffffc0u);
          // (a band so narrow that 1 / band would overflow the product -- rows scaled into the denormals by one huge
          //  element -- cannot be counted this way: every row of such a call is decided exactly)
          const float big_raw = 1.0f / (2.0f * bandD);
          const bool countable = big_raw <= 1.152921504606846976e18f;  // 2^60: D * big stays finite
          const float big = countable ? big_raw : 0.0f;
          const float off = -(vl - 4.0f * bandD) * big;
          float c4[4] = {0.0f, 0.0f, 0.0f, 0.0f};
          tc_ld32(taddr + (uint32_t)ch * 128, v[0]);
#pragma unroll
          for (int h = 0; h < 4; h++) {
            tc_wait_ld();
            tc_pin32(v[h & 1]);
            if (h < 3) tc_ld32(taddr + (uint32_t)(ch * 128 + (h + 1) * 32), v[(h + 1) & 1]);
#pragma unroll
            for (int j = 0; j < 32; j++) c4[j & 3] += __saturatef(fmaf(__uint_as_float(v[h & 1][j]), big, off));
          }
          cnt = (c4[0] + c4[1]) + (c4[2] + c4[3]);
          if (!(vl > ninf) || !countable) cnt = 2.0f;  // no finite maximum (NaN row, overflow) or no usable band: doubt
        }
        float4* xch = sm_x + (size_t)((it & 1) * 4 + lq) * 32 + lane;  // double-buffered by step parity
        if (ch == 1) {
          *xch = make_float4(b1, cnt, __int_as_float(bc), 0.0f);
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(tempty + a);                        // this warp is done with the accumulator
          // hand-over to the deciding warp; one named barrier per (step parity, quarter): this warp may run one
          // step ahead (the other accumulator), never two (tempty needs the deciding warp's arrival)
          asm volatile("bar.arrive %0, 64;" ::"r"(8 + (it & 1) * 4 + lq) : "memory");
          continue;
        }
        asm volatile("bar.sync %0, 64;" ::"r"(8 + (it & 1) * 4 + lq) : "memory");
        float b2;  // the best key of the half that lost: the only column of that half that can matter
        {
          const float4 o = *xch;
          if (o.x > b1) {
            b2 = b1;
            b1 = o.x;
            cnt = o.y;
            bc = __float_as_int(o.z);
          } else {
            b2 = o.x;
          }
        }
        int best = (int)((__float_as_uint(b1) & 63u) * 4u) + bc;
        const float v1 = __uint_as_float(__float_as_uint(b1) & 0xffffffc0u), v2 = __uint_as_float(__float_as_uint(b2) & 0xffffffc0u);
        const bool finite = v1 > ninf && v1 < __int_as_float(0x7f800000);
        // truncating 6 mantissa bits moves a key by at most 2^-17 |D|: part of the band (|D| <= s^2 (|x|^2 + |c|^2))
        // doubt: a second column of the winning half inside the band (cnt), or the other half's best inside it (v2)
        const bool doubt = live && (!finite || cnt > 1.5f || !(v1 - v2 > 2.0f * bandD) || best >= K);
        // Rows in doubt are NOT decided here.  Deciding them inline (re-reading the accumulator, gathering the row and the
        // in-band centroids from global memory) made such a step ~8x as long as a normal one, and because the accumulator
        // is released only when all four lane quarters are done, one doubtful quarter (29 % of the steps on uniform data)
        // stalled the whole CTA pipeline: 37 % of the epilogue's stall samples sat on the wait for the next accumulator
        // (profiles/r2_pq_tc_full.txt).  They are appended to a per-subspace-group list instead and decided by
        // pq_tc_resolve_kernel right after this kernel; the code written below is provisional for them.
        const unsigned dm = __ballot_sync(FULL_MASK, doubt);
        if (dm) {
          unsigned int base = 0;
          if (lane == 0) base = atomicAdd(dcount + sg, (unsigned int)__popc(dm));
          base = __shfl_sync(FULL_MASK, base, 0);
          if (doubt) {
            const unsigned int pos = base + (unsigned int)__popc(dm & ((1u << lane) - 1u));
            if (pos < dcap) dlist[(size_t)sg * dcap + pos] = ((unsigned long long)row << 8) | (unsigned long long)sub;
            // (beyond the capacity the count alone tells the resolver to decide every entry of this group)
          }
        }
        if (best >= K) best = 0;
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(tempty + a);
        if (live) {
          if (codes_u8) codes_u8[(size_t)row * M + sub] = (uint8_t)(best & 0xFF);
          if (assign_i32) assign_i32[(size_t)sub * n + row] = best;
        }
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

// Decides the entries pq_tc_assign_kernel left undecided, in the reference's arithmetic: fp32 distances to all K
// centroids nominate, everything inside the rounding band of the best is evaluated like Distances.l2Squared and compared
// with strict '<' in ascending index (the logic of pq_assign_kernel, pq.cu) -- independent of the tensor-core estimate.
// grid.y = subspace group; the group's centroids sit in shared memory.  A list that overflowed its capacity (degenerate
// data: everything in doubt) makes the group decide EVERY (row, subspace) of the slab.
constexpr int TQ_RESOLVE_THREADS = 256;
__global__ void __launch_bounds__(TQ_RESOLVE_THREADS)
pq_tc_resolve_kernel(const float* __restrict__ X, int64_t n, int d, int M, int K, const float* __restrict__ centroids, int lanes,
                     int64_t row0, int64_t nslab, const unsigned long long* __restrict__ dlist, const unsigned int* __restrict__ dcount,
                     unsigned int dcap, uint8_t* __restrict__ codes_u8, int32_t* __restrict__ assign_i32) {
  extern __shared__ __align__(16) float rcs[];  // [ns][K][8]
  const int sg = blockIdx.y;
  const unsigned int cnt = dcount[sg];
  if (cnt == 0) return;
  const int s_first = sg * TQ_SPC;
  const int ns = min(TQ_SPC, M - s_first);
  {
    const float4* src = reinterpret_cast<const float4*>(centroids + (size_t)s_first * K * TQ_SD);
    float4* dst = reinterpret_cast<float4*>(rcs);
    for (int i = threadIdx.x; i < ns * K * TQ_SD / 4; i += blockDim.x) dst[i] = src[i];
  }
  __syncthreads();
  const bool all = cnt > dcap;
  const int64_t items = all ? nslab * ns : (int64_t)cnt;
  for (int64_t it = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; it < items; it += (int64_t)gridDim.x * blockDim.x) {
    int64_t row;
    int sub;
    if (all) {
      row = row0 + it / ns;
      sub = s_first + (int)(it % ns);
    } else {
      const unsigned long long e = dlist[(size_t)sg * dcap + it];
      row = (int64_t)(e >> 8);
      sub = (int)(e & 255u);
    }
    const float* xr = X + (size_t)row * d + (size_t)sub * TQ_SD;
    float x[TQ_SD];
    {
      const float4 a = __ldg(reinterpret_cast<const float4*>(xr)), b = __ldg(reinterpret_cast<const float4*>(xr) + 1);
      x[0] = a.x; x[1] = a.y; x[2] = a.z; x[3] = a.w; x[4] = b.x; x[5] = b.y; x[6] = b.z; x[7] = b.w;
    }
    const float* c0 = rcs + (size_t)(sub - s_first) * K * TQ_SD;
    auto est = [&](int ci) {
      const float4* c = reinterpret_cast<const float4*>(c0 + (size_t)ci * TQ_SD);
      const float4 u = c[0], w = c[1];
      float e0 = 0.0f, e1 = 0.0f;
      const float t0 = x[0] - u.x, t1 = x[1] - u.y, t2 = x[2] - u.z, t3 = x[3] - u.w;
      const float t4 = x[4] - w.x, t5 = x[5] - w.y, t6 = x[6] - w.z, t7 = x[7] - w.w;
      e0 = fmaf(t0, t0, e0); e1 = fmaf(t1, t1, e1); e0 = fmaf(t2, t2, e0); e1 = fmaf(t3, t3, e1);
      e0 = fmaf(t4, t4, e0); e1 = fmaf(t5, t5, e1); e0 = fmaf(t6, t6, e0); e1 = fmaf(t7, t7, e1);
      return e0 + e1;
    };
    float m1 = __int_as_float(0x7f800000);
    for (int ci = 0; ci < K; ci++) m1 = fminf(m1, est(ci));  // NaN estimates are ignored by fminf
    const bool finite = m1 < __int_as_float(0x7f800000);
    // |estimate - reference| <= (2 SD + 24) 2^-24 relative on both sides; 3x slack (pq_band in pq.cu)
    const float lim = finite ? m1 * (1.0f + 3.0f * (float)(2 * TQ_SD + 24) * (1.0f / 16777216.0f)) + 1e-30f : __int_as_float(0x7f800000);
    const float* cg = centroids + (size_t)sub * K * TQ_SD;
    double bestDist = __longlong_as_double(0x7ff0000000000000ll);
    int best = 0;
    for (int ci = 0; ci < K; ci++) {
      const float e = est(ci);
      if (finite && !(e <= lim)) continue;  // outside the band (a NaN estimate with a finite best cannot be the argmin:
                                            //  its exact distance is NaN as well, and NaN never wins)
      const double dd = ref_sum_thread<REF_L2SQ>(xr, cg + (size_t)ci * TQ_SD, TQ_SD, lanes);
      if (dd < bestDist) {  // strict <: lowest ci wins ties, NaN never wins (PqEncoder.java:29)
        bestDist = dd;
        best = ci;
      }
    }
    if (codes_u8) codes_u8[(size_t)row * M + sub] = (uint8_t)(best & 0xFF);
    if (assign_i32) assign_i32[(size_t)sub * n + row] = best;
  }
}

// Scratch of the operand images comes from a private stream-ordered pool (one per device): allocation and release
// are ordered on the caller's stream -- no device-wide synchronisation -- and up to `pq_tc_keep_bytes` stay cached
// in the pool between calls, so the many small encode calls of an ingest path never reach the driver's allocator.
std::mutex g_pool_mu;
cudaMemPool_t g_pools[64] = {};
std::atomic<unsigned long long> g_pool_keep{16ull << 30};

cudaError_t tq_pool(cudaMemPool_t* out) {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  if (dev < 0 || dev >= 64) return cudaErrorInvalidDevice;
  std::lock_guard<std::mutex> g(g_pool_mu);
  if (!g_pools[dev]) {
    cudaMemPoolProps props = {};
    props.allocType = cudaMemAllocationTypePinned;
    props.handleTypes = cudaMemHandleTypeNone;
    props.location.type = cudaMemLocationTypeDevice;
    props.location.id = dev;
    if ((e = cudaMemPoolCreate(&g_pools[dev], &props)) != cudaSuccess) return e;
  }
  unsigned long long keep = g_pool_keep.load();
  if ((e = cudaMemPoolSetAttribute(g_pools[dev], cudaMemPoolAttrReleaseThreshold, &keep)) != cudaSuccess) return e;
  *out = g_pools[dev];
  return cudaSuccess;
}

struct PoolMem {  // released on the stream it was allocated on unless handed over
  void* p = nullptr;
  cudaStream_t st = nullptr;
  cudaError_t alloc(size_t bytes, cudaMemPool_t pool, cudaStream_t s) {
    st = s;
    return cudaMallocFromPoolAsync(&p, bytes, pool, s);
  }
  ~PoolMem() {
    if (p) cudaFreeAsync(p, st);
  }
};

}  // namespace

bool pq_tc_supported(const PqAssignLaunch& L) {
  return L.subDim == TQ_SD && L.K >= 1 && L.K <= TQ_N && L.d == L.M * TQ_SD && (L.M % 2) == 0 && L.s_begin == 0 &&
         L.s_end == L.M && encode_fn() != nullptr;
}

// Operand image kept between calls: PqTrainer.train assigns the SAME rows once per Lloyd iteration, so the trainer
// opens a scope (pq_tc_scope_begin / _end) within which the rows' image -- when it fits in one slab -- is built once
// and only the centroid image is refreshed.  Outside a scope the scratch goes back to the pool when the call's kernels
// have run.
namespace {
struct TcScope {
  bool open = false, valid = false;
  const float* X = nullptr;
  int64_t n = 0;
  int d = 0, M = 0;
  void *xop = nullptr, *cop = nullptr, *stats = nullptr;
  cudaStream_t st = nullptr;
  void release() {
    if (xop) cudaFreeAsync(xop, st);
    if (cop) cudaFreeAsync(cop, st);
    if (stats) cudaFreeAsync(stats, st);
    xop = cop = stats = nullptr;
    valid = false;
  }
};
thread_local TcScope t_scope;
}  // namespace

void pq_tc_scope_begin() {
  t_scope.release();
  t_scope.open = true;
}
void pq_tc_scope_end() {
  t_scope.release();
  t_scope.open = false;
}

// Tensor-core assignment over all subspaces.  The operand image of the rows (64 bytes per row and subspace) is built in
// scratch memory:
//   * inside a training scope, when the whole image fits the budget, it is built ONCE and kept for the Lloyd iterations;
//   * otherwise (every encode call) in two small slabs (TQ_SLAB_BYTES each) that alternate: the image of slab i+1 is
//     written on a helper stream while the tensor-core kernel of slab i runs, so the conversion pass hides behind the
//     assignment, and the 2 x 2 GB stay cached in the pool between calls -- an encode never reaches the driver's
//     allocator after the first one (a 24 GB slab per call, as round 1 had it, cost 0.1 s of mapping per 4 GB above the
//     pool's keep threshold: 100M rows encoded cold in 4.5 s instead of 0.6 s).
namespace {
constexpr size_t TQ_SLAB_BYTES = size_t(2) << 30;
struct HelperStream {
  cudaStream_t st[64] = {};
  ~HelperStream() {}  // process teardown: the context may be gone
};
thread_local HelperStream t_helper;
cudaError_t helper_stream(cudaStream_t* out) {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  if (dev < 0 || dev >= 64) return cudaErrorInvalidDevice;
  if (!t_helper.st[dev] && (e = cudaStreamCreateWithFlags(&t_helper.st[dev], cudaStreamNonBlocking)) != cudaSuccess) return e;
  *out = t_helper.st[dev];
  return cudaSuccess;
}
struct Ev {
  cudaEvent_t e = nullptr;
  cudaError_t make() { return cudaEventCreateWithFlags(&e, cudaEventDisableTiming); }
  ~Ev() {
    if (e) cudaEventDestroy(e);
  }
};
}  // namespace

cudaError_t launch_pq_assign_tc(const PqAssignLaunch& L, int sms, cudaStream_t st) {
  cudaError_t e;
  const int M = L.M;
  const int nsg = (M + TQ_SPC - 1) / TQ_SPC;
  const size_t per_row = (size_t)(M / 2) * 128;
  TcScope& sc = t_scope;
  const bool reuse = sc.open && sc.valid && sc.X == L.X && sc.n == L.n && sc.d == L.d && sc.M == M && sc.st == st;
  const int64_t n_pad = (L.n + TQ_M - 1) / TQ_M * TQ_M;
  int64_t slab = 0;
  bool whole = false;         // one image of all rows (kept by an open scope)
  PoolMem xop, xop2, cop, stats;  // owners when nothing is kept
  void *p_xop[2] = {nullptr, nullptr}, *p_cop, *p_stats;
  if (reuse) {
    slab = n_pad;
    whole = true;
    p_xop[0] = sc.xop; p_cop = sc.cop; p_stats = sc.stats;
  } else {
    cudaMemPool_t pool;
    if ((e = tq_pool(&pool)) != cudaSuccess) return e;
    if (sc.open) {  // a training run: keep the whole image if the device has room for it
      size_t free_b = 0, total_b = 0, cached = 0, used = 0;
      if ((e = cudaMemGetInfo(&free_b, &total_b)) != cudaSuccess) return e;
      cudaMemPoolGetAttribute(pool, cudaMemPoolAttrReservedMemCurrent, &cached);
      cudaMemPoolGetAttribute(pool, cudaMemPoolAttrUsedMemCurrent, &used);
      size_t budget = (free_b + (cached > used ? cached - used : 0)) / 2;  // what the pool holds idle is ours to use
      if (budget > (size_t(24) << 30)) budget = size_t(24) << 30;
      whole = (size_t)n_pad * per_row <= budget;
    }
    if (whole) {
      slab = n_pad;
      if ((e = xop.alloc((size_t)slab * per_row, pool, st)) != cudaSuccess) return e;
    } else {
      slab = (int64_t)(TQ_SLAB_BYTES / per_row) / TQ_M * TQ_M;
      if (slab < TQ_M) slab = TQ_M;
      if (slab > n_pad) slab = n_pad;
      if ((e = xop.alloc((size_t)slab * per_row, pool, st)) != cudaSuccess) return e;
      if (slab < n_pad && (e = xop2.alloc((size_t)slab * per_row, pool, st)) != cudaSuccess) return e;
    }
    if ((e = cop.alloc((size_t)(M / 2) * TQ_N * 128, pool, st)) != cudaSuccess) return e;
    if ((e = stats.alloc(sizeof(TqStats), pool, st)) != cudaSuccess) return e;
    p_xop[0] = xop.p; p_xop[1] = xop2.p ? xop2.p : xop.p; p_cop = cop.p; p_stats = stats.p;
  }
  // undecided entries of one slab: per subspace group a list of (row, subspace); room for one entry in eight
  PoolMem dl, dc;
  const unsigned int dcap = (unsigned int)std::max<int64_t>(4096, std::min<int64_t>(slab, L.n) * TQ_SPC / 8);
  {
    cudaMemPool_t pool;
    if ((e = tq_pool(&pool)) != cudaSuccess) return e;
    if ((e = dl.alloc((size_t)nsg * dcap * 8, pool, st)) != cudaSuccess) return e;
    if ((e = dc.alloc((size_t)nsg * 4, pool, st)) != cudaSuccess) return e;
  }
  unsigned long long* d_list = static_cast<unsigned long long*>(dl.p);
  unsigned int* d_cnt = static_cast<unsigned int*>(dc.p);
  const size_t resolve_smem = (size_t)TQ_SPC * L.K * TQ_SD * 4;
  if (resolve_smem > 48 * 1024 &&
      (e = cudaFuncSetAttribute(pq_tc_resolve_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)resolve_smem)) != cudaSuccess)
    return e;
  TqStats* d_st = static_cast<TqStats*>(p_stats);
  if (reuse) {
    // same rows, new centroids: the scale stays (centroids of a training run are means of the rows), the
    // centroid norm bound is refreshed
    if ((e = cudaMemsetAsync(&d_st->nmax_bits, 0, sizeof(unsigned int), st)) != cudaSuccess) return e;
  } else {
    if ((e = cudaMemsetAsync(d_st, 0, sizeof(TqStats), st)) != cudaSuccess) return e;
    // one scale for rows and centroids: max |element| of both
    count_launch();
    tq_amax_kernel<<<sms * 8, 256, 0, st>>>(L.X, L.n * L.d, d_st);
    count_launch();
    tq_amax_kernel<<<8, 256, 0, st>>>(L.centroids, (int64_t)M * L.K * TQ_SD, d_st);
  }
  count_launch();
  tq_prep_centroids_kernel<<<(M * TQ_N + 255) / 256, 256, 0, st>>>(L.centroids, M, L.K, static_cast<__half*>(p_cop), d_st);
  if ((e = cudaGetLastError()) != cudaSuccess) return e;
  CUtensorMap tmB;
  if (!encode_rows_map(&tmB, p_cop, (int64_t)(M / 2) * TQ_N, 64, 64, true, TQ_N)) return cudaErrorInvalidValue;
  if ((e = cudaFuncSetAttribute(pq_tc_assign_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TQ_SMEM)) != cudaSuccess) return e;
  // Two slabs alternate: the helper stream writes the image of slab i (after the assignment that last read that buffer)
  // while `st` runs the assignment of slab i-1.  NOTE the row-norm bound (xxmax) grows as slabs are converted: the
  // assignment of slab i reads it after the images of slabs 0..i, all of which it covers (a bound over MORE rows only
  // widens the band).
  const bool piped = !reuse && !whole && slab < n_pad;
  cudaStream_t hs = st;
  Ev ready[2], done[2], start;
  if (piped) {
    if ((e = helper_stream(&hs)) != cudaSuccess) return e;
    for (int i = 0; i < 2; i++)
      if ((e = ready[i].make()) != cudaSuccess || (e = done[i].make()) != cudaSuccess) return e;
    if ((e = start.make()) != cudaSuccess) return e;
    if ((e = cudaEventRecord(start.e, st)) != cudaSuccess) return e;       // scale, centroid image, fresh buffers
    if ((e = cudaStreamWaitEvent(hs, start.e, 0)) != cudaSuccess) return e;
  }
  int64_t si = 0;
  for (int64_t r0 = 0; r0 < L.n; r0 += slab, si++) {
    const int64_t cnt = std::min<int64_t>(slab, L.n - r0);
    const int b = (int)(si & 1);
    void* xb = p_xop[piped ? b : 0];
    if (!reuse) {
      if (piped && si >= 2 && (e = cudaStreamWaitEvent(hs, done[b].e, 0)) != cudaSuccess) return e;
      count_launch();
      tq_prep_rows_kernel<<<sms * 16, 256, 0, hs>>>(L.X + (size_t)r0 * L.d, cnt, L.d, M, static_cast<__half*>(xb), d_st);
      if ((e = cudaGetLastError()) != cudaSuccess) return e;
      if (piped) {
        if ((e = cudaEventRecord(ready[b].e, hs)) != cudaSuccess) return e;
        if ((e = cudaStreamWaitEvent(st, ready[b].e, 0)) != cudaSuccess) return e;
      }
    }
    CUtensorMap tmA;  // [cnt rows][M / 2 pairs * 64 halfs], box = 64 halfs x 128 rows
    if (!encode_rows_map(&tmA, xb, cnt, M * 32, (int64_t)M * 32, true, TQ_M)) return cudaErrorInvalidValue;
    const int64_t tiles = (cnt + TQ_M - 1) / TQ_M;
    int nsplit = sms / nsg;
    if (nsplit < 1) nsplit = 1;
    if (nsplit > tiles) nsplit = (int)tiles;
    if ((e = cudaMemsetAsync(d_cnt, 0, (size_t)nsg * 4, st)) != cudaSuccess) return e;
    count_launch();
    pq_tc_assign_kernel<<<nsg * nsplit, TQ_THREADS, TQ_SMEM, st>>>(tmA, tmB, L.X, L.n, L.d, M, L.K, L.centroids, L.lanes, d_st, r0,
                                                                   cnt, nsg, tiles, L.codes_u8, L.assign_i32, d_list, d_cnt, dcap);
    if ((e = cudaGetLastError()) != cudaSuccess) return e;
    if (piped && (e = cudaEventRecord(done[b].e, st)) != cudaSuccess) return e;  // (the resolver does not read the image)
    count_launch();
    pq_tc_resolve_kernel<<<dim3((unsigned)std::max(1, sms / nsg) * 2, nsg), TQ_RESOLVE_THREADS, resolve_smem, st>>>(
        L.X, L.n, L.d, M, L.K, L.centroids, L.lanes, r0, cnt, d_list, d_cnt, dcap, L.codes_u8, L.assign_i32);
    if ((e = cudaGetLastError()) != cudaSuccess) return e;
  }
  if (reuse) return cudaSuccess;  // everything this call touched lives until the scope ends
  if (sc.open && whole) {         // first call of a scope and the whole image fits: keep it
    sc.release();
    sc.X = L.X; sc.n = L.n; sc.d = L.d; sc.M = M; sc.st = st;
    sc.xop = xop.p; sc.cop = cop.p; sc.stats = stats.p;
    xop.p = cop.p = stats.p = nullptr;
    sc.valid = true;
    return cudaSuccess;
  }
  return cudaSuccess;  // the scratch is released in stream order behind the kernels that read it
}

void pq_tc_set_keep_bytes(unsigned long long bytes) { g_pool_keep.store(bytes); }

cudaError_t pq_pool_alloc(void** p, size_t bytes, cudaStream_t st) {
  cudaMemPool_t pool;
  cudaError_t e = tq_pool(&pool);
  if (e != cudaSuccess) return e;
  return cudaMallocFromPoolAsync(p, bytes, pool, st);
}

}  // namespace vs
