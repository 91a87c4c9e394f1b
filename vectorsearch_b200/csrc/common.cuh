// common.cuh -- shared device helpers of libvsgpu (sm_100a only).
//
// Two families live here:
//  (1) "reference arithmetic": device functions that reproduce, operation for operation, the
//      Java arithmetic of J/util/Distances.java (fp32 fused multiply-add per SIMD lane, ordered
//      ascending-lane fp32 reduction, fp64 tail; no contraction anywhere else).  Everything the
//      library RETURNS (scores, LUT entries, PQ argmins) comes from these.
//  (2) sortable keys: 128-bit (value, tiebreak) keys whose unsigned lexicographic order is the
//      reference's stable-sort order, used by the streaming top-k collectors (topk.cuh).
// J/ = /root/reference/src/main/java/io/github/panghy/vectorsearch/
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

namespace vs {

constexpr unsigned FULL_MASK = 0xffffffffu;

// development-only phase timestamps (build with -DVS_PHASE_STAMPS; tools/phase_stamps.py)
#ifdef VS_PHASE_STAMPS
static __device__ unsigned long long g_phase_stamps[8 * 1024];
__device__ __forceinline__ void phase_stamp(int ph) {
  if (threadIdx.x == 0 && blockIdx.y == 0 && blockIdx.x < 1024) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    g_phase_stamps[blockIdx.x * 8 + ph] = t;
  }
}
#else
__device__ __forceinline__ void phase_stamp(int) {}
#endif

// ----------------------------------------------------------------------------------------------
// sortable keys
// ----------------------------------------------------------------------------------------------
// hi: order-preserving image of the IEEE double being ranked ("smaller hi sorts first"),
// lo: tiebreak (row index for scans, candidate position for re-rank).  Both unsigned.
struct Key {
  uint64_t hi, lo;
};
constexpr uint64_t KEY_EMPTY64 = ~0ull;
__device__ __forceinline__ Key key_empty() { return Key{KEY_EMPTY64, KEY_EMPTY64}; }
// branch-free on purpose: short-circuit evaluation turns every compare into a divergent branch
__device__ __forceinline__ bool key_lt(const Key& a, const Key& b) {
  return (a.hi < b.hi) | ((a.hi == b.hi) & (a.lo < b.lo));
}
__device__ __forceinline__ bool key_is_empty(const Key& a) { return a.lo == KEY_EMPTY64; }
__device__ __forceinline__ Key ld_key(const ulonglong2* p) {
  ulonglong2 v = *p;
  return Key{v.x, v.y};
}
__device__ __forceinline__ void st_key(ulonglong2* p, const Key& k) { *p = make_ulonglong2(k.hi, k.lo); }

// Monotone map double -> uint64 in java.lang.Double.compare order: -inf < ... < -0.0 < +0.0 < ...
// < +inf < NaN (NaN canonicalised like Double.doubleToLongBits).
__device__ __forceinline__ uint64_t f64_ordered(double d) {
  uint64_t u = (d != d) ? 0x7ff8000000000000ull : (uint64_t)__double_as_longlong(d);
  return u ^ ((u >> 63) ? ~0ull : 0x8000000000000000ull);
}
__device__ __forceinline__ double f64_from_ordered(uint64_t o) {
  uint64_t u = (o >> 63) ? (o ^ 0x8000000000000000ull) : ~o;
  return __longlong_as_double((long long)u);
}
// Ranking image of a SCORE sorted descending with Double.compare (brute force and re-rank,
// J/fdb/FdbVectorIndex.java:708,1031): NaN is the largest score, so it sorts first (hi = 0);
// otherwise the image of -score.  Unary minus is exact, so the score is recovered exactly.
__device__ __forceinline__ uint64_t rank_hi_from_score(double score) {
  if (score != score) return 0ull;
  return f64_ordered(-score);  // never 0: that would need -score == -NaN
}
__device__ __forceinline__ double score_from_rank_hi(uint64_t hi) {
  if (hi == 0ull) return __longlong_as_double(0x7ff8000000000000ll);
  return -f64_from_ordered(hi);
}
// Ranking image of a DISTANCE sorted ascending with Double.compare (ADC scan, :769): NaN last.
__device__ __forceinline__ uint64_t rank_hi_from_dist(double dist) {
  uint64_t h = f64_ordered(dist);
  return h == KEY_EMPTY64 ? KEY_EMPTY64 - 1 : h;  // cannot happen for canonical NaN; keep EMPTY unique
}
__device__ __forceinline__ double dist_from_rank_hi(uint64_t hi) { return f64_from_ordered(hi); }

// ----------------------------------------------------------------------------------------------
// reference arithmetic, one thread per pair
// ----------------------------------------------------------------------------------------------
enum RefOp { REF_L2SQ = 0, REF_DOT = 1 };

// Sum over i of (a[i]-b[i])^2 or a[i]*b[i] exactly as Distances.l2Squared / dot compute it with
// an L-lane FloatVector: J/util/Distances.java:48-64 (:77-94), :103-118.
template <int L, int OP>
__device__ __forceinline__ double ref_sum_thread_L(const float* __restrict__ a,
                                                   const float* __restrict__ b, int len) {
  const int ub = len - (len % L);  // SPECIES.loopBound
  float acc[L];
#pragma unroll
  for (int l = 0; l < L; l++) acc[l] = 0.0f;
  for (int i = 0; i < ub; i += L) {
#pragma unroll
    for (int l = 0; l < L; l++) {
      if (OP == REF_L2SQ) {
        float diff = __fsub_rn(a[i + l], b[i + l]);
        acc[l] = __fmaf_rn(diff, diff, acc[l]);
      } else {
        acc[l] = __fmaf_rn(a[i + l], b[i + l], acc[l]);
      }
    }
  }
  float s = 0.0f;  // reduceLanes(ADD): ordered, ascending lane
#pragma unroll
  for (int l = 0; l < L; l++) s = __fadd_rn(s, acc[l]);
  double sum = (double)s;
  for (int i = ub; i < len; i++) {  // scalar tail in double, no contraction
    if (OP == REF_L2SQ) {
      double d = __dsub_rn((double)a[i], (double)b[i]);
      sum = __dadd_rn(sum, __dmul_rn(d, d));
    } else {
      sum = __dadd_rn(sum, __dmul_rn((double)a[i], (double)b[i]));
    }
  }
  return sum;
}

template <int OP>
__device__ __forceinline__ double ref_sum_thread(const float* __restrict__ a,
                                                 const float* __restrict__ b, int len, int lanes) {
  switch (lanes) {
    case 16: return ref_sum_thread_L<16, OP>(a, b, len);
    case 8: return ref_sum_thread_L<8, OP>(a, b, len);
    case 4: return ref_sum_thread_L<4, OP>(a, b, len);
    case 2: return ref_sum_thread_L<2, OP>(a, b, len);
    default: return ref_sum_thread_L<1, OP>(a, b, len);
  }
}

// The scalar tail alone (elements [ub, len)), added to an already reduced vector-loop sum.
template <int OP>
__device__ __forceinline__ double ref_add_tail(double sum, const float* __restrict__ a,
                                               const float* __restrict__ b, int ub, int len) {
#pragma unroll 1
  for (int i = ub; i < len; i++) {
    if (OP == REF_L2SQ) {
      double d = __dsub_rn((double)a[i], (double)b[i]);
      sum = __dadd_rn(sum, __dmul_rn(d, d));
    } else {
      sum = __dadd_rn(sum, __dmul_rn((double)a[i], (double)b[i]));
    }
  }
  return sum;
}

// Same value, computed cooperatively by the 16 threads of a half-warp (hl = lane & 15).
// Lane l of the modelled SIMD register lives in thread hl == l; all 16 threads return the sum.
template <int OP>
__device__ __forceinline__ double ref_sum_halfwarp(const float* __restrict__ a,
                                                   const float* __restrict__ b, int len, int lanes,
                                                   int hl, unsigned mask, int base_lane) {
  const int ub = len - (len % lanes);
  float acc = 0.0f;
  if (hl < lanes) {
#pragma unroll 8
    for (int i = hl; i < ub; i += lanes) {
      if (OP == REF_L2SQ) {
        float diff = __fsub_rn(a[i], b[i]);
        acc = __fmaf_rn(diff, diff, acc);
      } else {
        acc = __fmaf_rn(a[i], b[i], acc);
      }
    }
  }
  float s = 0.0f;
  for (int l = 0; l < lanes; l++) s = __fadd_rn(s, __shfl_sync(mask, acc, base_lane + l));
  return ref_add_tail<OP>((double)s, a, b, ub, len);
}

// Distances.cosine(a,b) = dot/(norm(a)*norm(b)), 0 when the product of norms is 0 (:149-153);
// also the normalizeOnRead branch of fetchExactAndScore (FdbVectorIndex.java:1006-1010), which is
// the same expression with norm(q) hoisted.  qq / xx are the squared norms before Math.sqrt.
__device__ __forceinline__ double ref_cosine_from_sums(double dot, double qq, double xx) {
  double n = __dmul_rn(__dsqrt_rn(qq), __dsqrt_rn(xx));
  if (n == 0.0) return 0.0;
  return __ddiv_rn(dot, n);
}

__device__ __forceinline__ float4 ld_stream_f4(const float4* p) {
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.L2::128B.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
               : "l"(p));
  return r;
}
__device__ __forceinline__ uint4 ld_stream_u4(const uint4* p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.L2::128B.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(p));
  return r;
}

// programmatic dependent launch (see launch_pdl in kernels.h)
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

// ---- mbarrier + bulk asynchronous copy (TMA without a tensor map: 1-D, 16-byte granular) ----------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  do {
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
                 : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  } while (!ok);
}
// The same with a suspend-time hint: the thread sleeps in the barrier unit until the phase completes (it is resumed at
// once when it does) instead of coming back every few hundred cycles to issue another try -- the spinning single-thread
// roles of a warp-specialised kernel otherwise steal issue slots from the working warps of their SM sub-partition.
__device__ __forceinline__ void mbar_wait_suspend(uint64_t* bar, uint32_t parity, uint32_t hint_ns = 20000u) {
  uint32_t ok;
  do {
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\nselp.u32 %0, 1, 0, p;\n}\n"
                 : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity), "r"(hint_ns) : "memory");
  } while (!ok);
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

// next float toward +inf (finite, non-NaN input)
__device__ __forceinline__ float f32_next_up(float f) {
  if (f == 0.0f) return __uint_as_float(1u);
  uint32_t u = __float_as_uint(f);
  return __uint_as_float((u >> 31) ? u - 1u : u + 1u);
}

}  // namespace vs
