// topk.cuh -- streaming top-k over 128-bit sortable keys (smaller key = better).
//
// Replaces the reference's "score everything, List.sort, subList(0,k)"
// (J/fdb/FdbVectorIndex.java:708-721, :769, :820-822, :1031-1043).  Keys are (exact value image,
// row) pairs with unique rows, so "the k smallest keys" is exactly the head of the reference's
// stable sort.
//
// Each warp owns a sorted list of KP keys and a 64-entry staging buffer in shared memory.  The
// streaming kernels pre-filter rows with one fp32 compare against a warp-uniform threshold, so
// the steady-state cost per scored row is that compare; only rows that can still enter the top-k
// get their exact fp64 value, are staged, and a full buffer is bitonic-sorted and folded into the
// list with the min(list[i], buf[KP-1-i]) half-cleaner.  No atomics, no block-wide barriers
// inside the streaming loop.
#pragma once

#include "common.cuh"

namespace vs {

constexpr int TOPK_BUF = 64;
constexpr int TOPK_MAX_K = 1024;

__host__ __device__ inline int topk_pad(int k) {
  int p = 32;
  while (p < k) p <<= 1;
  return p;
}
// shared memory bytes one warp needs for a list of capacity kp
__host__ __device__ inline size_t topk_warp_smem(int kp) { return (size_t)(kp + TOPK_BUF) * 16; }

__device__ __forceinline__ void cswap(ulonglong2* a, int i, int j, bool asc) {
  Key x = ld_key(a + i), y = ld_key(a + j);
  if (key_lt(y, x) == asc) {
    st_key(a + i, y);
    st_key(a + j, x);
  }
}

// a[0..n) is bitonic (n power of two >= 2): sort ascending. Warp-cooperative.
__device__ __forceinline__ void warp_bitonic_merge(ulonglong2* a, int n, int lane) {
  for (int stride = n >> 1; stride > 0; stride >>= 1) {
    for (int t = lane; t < (n >> 1); t += 32) {
      int i = ((t & ~(stride - 1)) << 1) | (t & (stride - 1));
      cswap(a, i, i | stride, true);
    }
    __syncwarp();
  }
}

// full ascending bitonic sort of a[0..n), n power of two >= 2. Warp-cooperative.
__device__ __forceinline__ void warp_bitonic_sort(ulonglong2* a, int n, int lane) {
  for (int size = 2; size <= n; size <<= 1) {
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      for (int t = lane; t < (n >> 1); t += 32) {
        int i = ((t & ~(stride - 1)) << 1) | (t & (stride - 1));
        cswap(a, i, i | stride, (i & size) == 0);
      }
      __syncwarp();
    }
  }
}

// dst[0..kp) and src[0..kp) sorted ascending -> dst = the kp smallest of the union, sorted.
__device__ __forceinline__ void warp_merge_lists(ulonglong2* dst, const ulonglong2* src, int kp,
                                                 int lane) {
  for (int i = lane; i < kp; i += 32) {
    Key o = ld_key(src + (kp - 1 - i));
    if (key_lt(o, ld_key(dst + i))) st_key(dst + i, o);
  }
  __syncwarp();
  warp_bitonic_merge(dst, kp, lane);
}

struct WarpTopK {
  ulonglong2* list;  // [kp] ascending, EMPTY padded
  ulonglong2* buf;   // [TOPK_BUF]
  Key thr;           // current k-th best (warp-uniform); EMPTY until k keys were seen
  int kp, k, cnt;

  __device__ __forceinline__ void init(ulonglong2* smem, int kp_, int k_, int lane) {
    list = smem;
    buf = smem + kp_;
    kp = kp_;
    k = k_;
    cnt = 0;
    thr = key_empty();
    for (int i = lane; i < kp_; i += 32) st_key(list + i, key_empty());
    __syncwarp();
  }

  __device__ __forceinline__ void flush(int lane) {
    if (cnt == 0) return;
    for (int i = cnt + lane; i < TOPK_BUF; i += 32) st_key(buf + i, key_empty());
    __syncwarp();
    warp_bitonic_sort(buf, TOPK_BUF, lane);
    for (int i = lane; i < kp; i += 32) {
      int j = kp - 1 - i;
      if (j < TOPK_BUF) {
        Key o = ld_key(buf + j);
        if (key_lt(o, ld_key(list + i))) st_key(list + i, o);
      }
    }
    __syncwarp();
    warp_bitonic_merge(list, kp, lane);
    thr = ld_key(list + (k - 1));
    cnt = 0;
    __syncwarp();
  }

  // All 32 lanes must call; lanes with pred offer `key`.  Keys not below thr are dropped here.
  __device__ __forceinline__ void push(const Key& key, bool pred, int lane) {
    pred = pred && key_lt(key, thr);
    unsigned m = __ballot_sync(FULL_MASK, pred);
    if (m == 0) return;
    if (pred) st_key(buf + (cnt + __popc(m & ((1u << lane) - 1u))), key);
    cnt += __popc(m);
    __syncwarp();
    if (cnt > TOPK_BUF - 32) flush(lane);
  }
};

// Combine the per-warp lists of a block into warp 0's list. lists = base of nw lists laid out
// with stride `stride_keys`. All threads of the block must call (uses __syncthreads).
__device__ __forceinline__ void block_combine_lists(ulonglong2* lists, int stride_keys, int nw, int kp,
                                                    int warp, int lane) {
  __syncthreads();
  for (int step = 1; step < nw; step <<= 1) {
    if ((warp % (2 * step)) == 0 && warp + step < nw)
      warp_merge_lists(lists + (size_t)warp * stride_keys, lists + (size_t)(warp + step) * stride_keys,
                       kp, lane);
    __syncthreads();
  }
}

// Merge `total` keys from global memory (per-CTA partial lists) into the block's warp-0 list.
__device__ __forceinline__ void block_collect_keys(const ulonglong2* __restrict__ keys, int64_t total,
                                                   ulonglong2* smem, int kp, int k) {
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int nw = blockDim.x >> 5;
  const int stride_keys = kp + TOPK_BUF;
  WarpTopK tk;
  tk.init(smem + (size_t)warp * stride_keys, kp, k, lane);
  for (int64_t i0 = (int64_t)warp * 32; i0 < total; i0 += (int64_t)nw * 32) {
    int64_t i = i0 + lane;
    Key key = key_empty();
    if (i < total) key = ld_key(keys + i);
    tk.push(key, !key_is_empty(key), lane);
  }
  tk.flush(lane);
  block_combine_lists(smem, stride_keys, nw, kp, warp, lane);
}


struct TopkOut {
  ulonglong2* partial;   // [nq][grid][k]
  unsigned long long* ctrl;  // [nq][4] zero between launches: ticket, append count, ~global bound, tile counter
  int64_t partial_keys;      // keys of `partial` reserved per query
  int64_t* ids;          // [nq][k]
  double* scores;        // [nq][k]
  int32_t* counts;       // [nq]
  int64_t id_base;
  int score_kind;        // 0: hi is a descending-score image, 1: hi is an ascending-distance image
  int64_t out_stride;    // elements between consecutive queries in ids / scores (>= k)
};

// ---- register-resident collector for k <= 32 ------------------------------------------------------
// One key per lane, sorted ascending across the warp.  An insertion is a ballot (position), one
// shuffle-up and a broadcast of the new k-th key: the threshold tightens after EVERY accepted
// row, so few rows survive the fp32 pre-filter and nothing is ever sorted.
constexpr int TOPK_REG_MAX_K = 32;
constexpr int TOPK_MAX_LISTS = 320;  // per-query CTA lists the fast final merge can take

struct WarpTopKReg {
  uint64_t my_hi, my_lo;
  Key thr;
  int k;

  __device__ __forceinline__ void init(ulonglong2*, int, int k_, int) {
    my_hi = KEY_EMPTY64;
    my_lo = KEY_EMPTY64;
    thr = key_empty();
    k = k_;
  }
  __device__ __forceinline__ void flush(int) {}
  __device__ __forceinline__ void push(const Key& key, bool pred, int lane) {
    pred = pred && key_lt(key, thr);
    unsigned m = __ballot_sync(FULL_MASK, pred);
    while (m) {
      const int src = __ffs(m) - 1;
      m &= m - 1;
      const Key c{__shfl_sync(FULL_MASK, key.hi, src), __shfl_sync(FULL_MASK, key.lo, src)};
      if (!key_lt(c, thr)) continue;  // the threshold may have tightened since the ballot
      const bool lt = key_lt(Key{my_hi, my_lo}, c);
      const int pos = __popc(__ballot_sync(FULL_MASK, lt));  // sorted: the smaller keys are a prefix
      const uint64_t uh = __shfl_up_sync(FULL_MASK, my_hi, 1), ul = __shfl_up_sync(FULL_MASK, my_lo, 1);
      if (lane > pos) {
        my_hi = uh;
        my_lo = ul;
      } else if (lane == pos) {
        my_hi = c.hi;
        my_lo = c.lo;
      }
      thr = Key{__shfl_sync(FULL_MASK, my_hi, k - 1), __shfl_sync(FULL_MASK, my_lo, k - 1)};
    }
  }
};

// shared memory (bytes) the block-level top-k machinery needs, for either collector
__host__ __device__ inline size_t topk_block_smem(int k, int kp, int nw) {
  if (k <= TOPK_REG_MAX_K) {
    const int surv = k * k > TOPK_MAX_LISTS ? k * k : TOPK_MAX_LISTS;
    return (size_t)(nw * k + k + surv) * 16;
  }
  return (size_t)nw * topk_warp_smem(kp);
}
// keys of `partial` scratch one query needs
__host__ __device__ inline int64_t topk_partial_keys(int k, int grid, int nw) {
  (void)nw;
  return (int64_t)grid * k;
}

// rank of every key among `total` keys in shared memory; keys with rank < k land in outk[rank]
__device__ __forceinline__ void block_rank_select(const ulonglong2* keys, int total, int k, ulonglong2* outk) {
  for (int i = threadIdx.x; i < total; i += blockDim.x) {
    const Key me = ld_key(keys + i);
    if (key_is_empty(me)) continue;
    int rank = 0;
#pragma unroll 4
    for (int j = 0; j < total; j++) rank += key_lt(ld_key(keys + j), me) ? 1 : 0;
    if (rank < k) st_key(outk + rank, me);
  }
}

__device__ __forceinline__ void topk_write_out(const ulonglong2* outk, int k, const TopkOut& o, int qi) {
  __shared__ int s_found;
  if (threadIdx.x == 0) s_found = 0;
  __syncthreads();
  int found = 0;
  for (int i = threadIdx.x; i < k; i += blockDim.x) {
    Key e = ld_key(outk + i);
    const bool ok = !key_is_empty(e);
    o.ids[(size_t)qi * o.out_stride + i] = ok ? o.id_base + (int64_t)e.lo : -1;
    o.scores[(size_t)qi * o.out_stride + i] = ok ? (o.score_kind == 0 ? score_from_rank_hi(e.hi) : dist_from_rank_hi(e.hi))
                                                 : __longlong_as_double(0x7ff8000000000000ll);
    found += ok ? 1 : 0;
  }
  if (found) atomicAdd(&s_found, found);
  __syncthreads();
  if (threadIdx.x == 0) {
    o.counts[qi] = s_found;
    o.ctrl[4 * qi + 0] = 0ull;  // ready for the next launch on this scratch
    o.ctrl[4 * qi + 1] = 0ull;
    o.ctrl[4 * qi + 2] = 0ull;
    o.ctrl[4 * qi + 3] = 0ull;
  }
}

// ---- shared epilogue of the streaming kernels ---------------------------------------------------

// flush + block combine + publish + (last CTA) final merge.  smem holds nw collectors.
// Both overloads return true in the CTA that ran the final merge (the last one of its query).
static __device__ __noinline__ bool topk_epilogue(WarpTopK& tk, ulonglong2* smem, int kp, int k,
                                              const TopkOut& o, int qi_or_neg = -1) {
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int nw = blockDim.x >> 5;
  const int stride_keys = kp + TOPK_BUF;
  const int qi = qi_or_neg >= 0 ? qi_or_neg : (int)blockIdx.y;  // query slot of ids / scores / partial / ctrl
  tk.flush(lane);
  block_combine_lists(smem, stride_keys, nw, kp, warp, lane);
  ulonglong2* mine = o.partial + (size_t)qi * o.partial_keys + (size_t)blockIdx.x * k;
  for (int i = threadIdx.x; i < k; i += blockDim.x) __stcg(mine + i, smem[i]);
  __threadfence();
  __syncthreads();
  __shared__ unsigned int s_last;
  if (threadIdx.x == 0) {
    unsigned long long t = atomicAdd(o.ctrl + 4 * qi, 1ull);
    s_last = (t == gridDim.x - 1) ? 1u : 0u;
  }
  __syncthreads();
  if (!s_last) return false;
  __threadfence();
  // final merge by the last CTA of this query
  {
    const ulonglong2* all = o.partial + (size_t)qi * o.partial_keys;
    const int64_t total = (int64_t)gridDim.x * k;
    WarpTopK fk;
    fk.init(smem + (size_t)warp * stride_keys, kp, k, lane);
    for (int64_t i0 = (int64_t)warp * 32; i0 < total; i0 += (int64_t)nw * 32) {
      int64_t i = i0 + lane;
      Key key = key_empty();
      if (i < total) {
        ulonglong2 v = __ldcg(all + i);
        key = Key{v.x, v.y};
      }
      fk.push(key, !key_is_empty(key), lane);
    }
    fk.flush(lane);
    block_combine_lists(smem, stride_keys, nw, kp, warp, lane);
  }
  __shared__ int s_found;
  if (threadIdx.x == 0) s_found = 0;
  __syncthreads();
  int found = 0;
  for (int i = threadIdx.x; i < k; i += blockDim.x) {
    Key e = ld_key(smem + i);
    const bool ok = !key_is_empty(e);
    o.ids[(size_t)qi * o.out_stride + i] = ok ? o.id_base + (int64_t)e.lo : -1;
    o.scores[(size_t)qi * o.out_stride + i] = ok ? (o.score_kind == 0 ? score_from_rank_hi(e.hi) : dist_from_rank_hi(e.hi))
                                      : __longlong_as_double(0x7ff8000000000000ll);
    found += ok ? 1 : 0;
  }
  if (found) atomicAdd(&s_found, found);
  __syncthreads();
  if (threadIdx.x == 0) {
    o.counts[qi] = s_found;
    o.ctrl[4 * qi] = 0ull;  // ready for the next launch on this scratch
  }
  return true;
}

// Epilogue of the register collector: block combine by rank selection, publish the CTA's k keys,
// and (last CTA of the query) the final merge.  The k-th smallest list head bounds the launch-wide
// k-th key, so only keys at or below that bound -- at most k lists x k keys -- are gathered and
// rank-selected; nothing is sorted and nothing scales with grid x k beyond two coalesced reads.
static __device__ __noinline__ bool topk_epilogue(WarpTopKReg& tk, ulonglong2* smem, int kp, int k,
                                              const TopkOut& o, int qi_or_neg = -1) {
  (void)kp;
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int nw = blockDim.x >> 5;
  const int qi = qi_or_neg >= 0 ? qi_or_neg : (int)blockIdx.y;  // query slot of ids / scores / partial / ctrl
  unsigned long long* ctrl = o.ctrl + 4 * qi;
  ulonglong2* keys = smem;            // [nw * k]
  ulonglong2* outk = smem + nw * k;   // [k]
  ulonglong2* surv = outk + k;        // [max(k*k, TOPK_MAX_LISTS)]
  if (lane < k) st_key(keys + warp * k + lane, Key{tk.my_hi, tk.my_lo});
  if (threadIdx.x < k) st_key(outk + threadIdx.x, key_empty());
  __syncthreads();
  block_rank_select(keys, nw * k, k, outk);
  __syncthreads();
  phase_stamp(5);
  ulonglong2* all = o.partial + (size_t)qi * o.partial_keys;
  for (int i = threadIdx.x; i < k; i += blockDim.x) __stcg(all + (size_t)blockIdx.x * k + i, outk[i]);
  __threadfence();
  __syncthreads();
  __shared__ unsigned int s_last;
  __shared__ int s_cnt;
  __shared__ ulonglong2 s_bound;
  if (threadIdx.x == 0) {
    unsigned long long t = atomicAdd(ctrl, 1ull);
    s_last = (t == gridDim.x - 1) ? 1u : 0u;
    s_cnt = 0;
    st_key(&s_bound, key_empty());
  }
  __syncthreads();
  if (!s_last) return false;
  __threadfence();
  phase_stamp(6);
  const int nl = gridDim.x;  // <= TOPK_MAX_LISTS (host guarantees)
  const int total = nl * k;
  // every published key is requested now (up to PRE per thread, in registers) so that the filter below does not
  // start a second round trip to L2 after the bound is known
  constexpr int PRE = 8;
  ulonglong2 pre[PRE];
  const bool cached = total <= PRE * (int)blockDim.x;
  if (cached) {
#pragma unroll
    for (int i = 0; i < PRE; i++) {
      const int idx = threadIdx.x + i * blockDim.x;
      pre[i] = idx < total ? __ldcg(all + idx) : make_ulonglong2(KEY_EMPTY64, KEY_EMPTY64);
    }
  }
  for (int i = threadIdx.x; i < nl; i += blockDim.x) surv[i] = __ldcg(all + (size_t)i * k);
  if (threadIdx.x < k) st_key(outk + threadIdx.x, key_empty());
  __syncthreads();
  phase_stamp(1);
  for (int i = threadIdx.x; i < nl; i += blockDim.x) {
    const Key h = ld_key(surv + i);
    if (key_is_empty(h)) continue;
    int rank = 0;
#pragma unroll 4
    for (int j = 0; j < nl; j++) rank += key_lt(ld_key(surv + j), h) ? 1 : 0;
    if (rank == k - 1) st_key(&s_bound, h);
  }
  __syncthreads();
  const Key B = ld_key(&s_bound);
  phase_stamp(2);
  __syncthreads();  // heads are dead from here on: surv is reused for the survivors
  if (cached) {
#pragma unroll
    for (int i = 0; i < PRE; i++) {
      const Key x{pre[i].x, pre[i].y};
      if (!key_is_empty(x) && !key_lt(B, x)) {
        const int slot = atomicAdd(&s_cnt, 1);
        st_key(surv + slot, x);  // at most k lists reach below the bound: slot < k*k
      }
    }
  } else {
    for (int i = threadIdx.x; i < total; i += blockDim.x) {
      const ulonglong2 v = __ldcg(all + i);
      const Key x{v.x, v.y};
      if (!key_is_empty(x) && !key_lt(B, x)) {
        const int slot = atomicAdd(&s_cnt, 1);
        st_key(surv + slot, x);
      }
    }
  }
  __syncthreads();
  phase_stamp(3);
  block_rank_select(surv, s_cnt, k, outk);
  __syncthreads();
  phase_stamp(7);
  topk_write_out(outk, k, o, qi);
  return true;
}

}  // namespace vs
