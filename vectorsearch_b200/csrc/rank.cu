// rank.cu -- exact scoring + final ordering of a short candidate list, and the cross-segment merge.
//
// Re-rank = fetchExactAndScore (J/fdb/FdbVectorIndex.java:997-1043): candidates are scored IN THE
// GIVEN ORDER with Distances.l2 / cosine (normalizeOnRead evaluates the same expression with
// norm(q) hoisted, :1006-1010), missing / deleted / gid-less records are dropped (:1000,:1022),
// the list is stably sorted by score descending (:1031) and cut to k (:1042): ties keep candidate
// order.  Merge = query() :432-437: concatenated per-segment lists, stable sort by score
// descending, first k.
//
// One CTA per query.  A half-warp scores one candidate with the reference's lane arithmetic
// (thread l of the half-warp is SIMD lane l), the block sorts (score image, position) keys.
#include "kernels.h"
#include "topk.cuh"

namespace vs {

constexpr int RANK_THREADS = 256;
constexpr int RANK_MAX_CAND = 8192;

__device__ __forceinline__ void block_bitonic_sort_keys(ulonglong2* a, int n) {
  for (int size = 2; size <= n; size <<= 1) {
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      for (int t = threadIdx.x; t < (n >> 1); t += blockDim.x) {
        int i = ((t & ~(stride - 1)) << 1) | (t & (stride - 1));
        cswap(a, i, i | stride, (i & size) == 0);
      }
      __syncthreads();
    }
  }
}

__global__ void __launch_bounds__(1024)
rank_kernel(const float* __restrict__ X, int64_t n, int d, const uint8_t* __restrict__ skip, int lanes,
            const float* __restrict__ Q, const int64_t* cand_ids /* predecessor's output (launch_pdl): no __restrict__ */, int nc, int np, int k,
            int cosine, int64_t id_base, int64_t* __restrict__ ids_out, double* __restrict__ scores_out,
            int32_t* __restrict__ counts_out) {
  extern __shared__ __align__(16) ulonglong2 skey[];  // [np]
  pdl_wait();  // the candidate ids come from the kernel before (ADC scan / its fallback check)
  const int qi = blockIdx.x;
  const float* q = Q + (size_t)qi * d;
  const int64_t* cand = cand_ids + (size_t)qi * nc;
  const int lane = threadIdx.x & 31;
  const int hl = lane & 15;
  const int hw = threadIdx.x >> 4;
  const int nhw = blockDim.x >> 4;
  const unsigned hmask = (lane < 16) ? 0x0000ffffu : 0xffff0000u;
  const int base_lane = lane & 16;

  for (int i = threadIdx.x; i < np; i += blockDim.x) st_key(skey + i, key_empty());
  __syncthreads();
  double qq = 0.0;
  if (cosine) qq = ref_sum_halfwarp<REF_DOT>(q, q, d, lanes, hl, hmask, base_lane);
  for (int c = hw; c < nc; c += nhw) {
    const int64_t g = cand[c];
    const int64_t row = g - id_base;
    bool ok = g >= 0 && row >= 0 && row < n;      // rec == null otherwise
    if (ok && skip != nullptr && skip[row]) ok = false;  // deleted / gid missing
    if (!ok) continue;                             // uniform across the half-warp
    const float* x = X + (size_t)row * d;
    double score;
    if (cosine) {
      const double dot = ref_sum_halfwarp<REF_DOT>(q, x, d, lanes, hl, hmask, base_lane);
      const double xx = ref_sum_halfwarp<REF_DOT>(x, x, d, lanes, hl, hmask, base_lane);
      score = ref_cosine_from_sums(dot, qq, xx);
    } else {
      score = -__dsqrt_rn(ref_sum_halfwarp<REF_L2SQ>(q, x, d, lanes, hl, hmask, base_lane));
    }
    if (hl == 0) st_key(skey + c, Key{rank_hi_from_score(score), (uint64_t)c});
  }
  __syncthreads();
  block_bitonic_sort_keys(skey, np);
  __shared__ int s_found;
  if (threadIdx.x == 0) s_found = 0;
  __syncthreads();
  int found = 0;
  for (int i = threadIdx.x; i < k; i += blockDim.x) {
    Key e = i < np ? ld_key(skey + i) : key_empty();
    const bool ok = !key_is_empty(e);
    ids_out[(size_t)qi * k + i] = ok ? cand[e.lo] : -1;
    scores_out[(size_t)qi * k + i] = ok ? score_from_rank_hi(e.hi) : __longlong_as_double(0x7ff8000000000000ll);
    found += ok ? 1 : 0;
  }
  if (found) atomicAdd(&s_found, found);
  __syncthreads();
  if (threadIdx.x == 0) counts_out[qi] = s_found;
}

__global__ void __launch_bounds__(RANK_THREADS)
merge_kernel(const int64_t* __restrict__ ids, const double* __restrict__ scores, int64_t total, int k,
             int kp, int64_t* __restrict__ ids_out, double* __restrict__ scores_out,
             int32_t* __restrict__ count_out) {
  extern __shared__ __align__(128) ulonglong2 smem[];
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int nw = blockDim.x >> 5;
  const int stride_keys = kp + TOPK_BUF;
  WarpTopK tk;
  tk.init(smem + (size_t)warp * stride_keys, kp, k, lane);
  for (int64_t i0 = (int64_t)warp * 32; i0 < total; i0 += (int64_t)nw * 32) {
    const int64_t i = i0 + lane;
    Key key = key_empty();
    if (i < total) key = Key{rank_hi_from_score(scores[i]), (uint64_t)i};
    tk.push(key, i < total, lane);
  }
  tk.flush(lane);
  block_combine_lists(smem, stride_keys, nw, kp, warp, lane);
  __shared__ int s_found;
  if (threadIdx.x == 0) s_found = 0;
  __syncthreads();
  int found = 0;
  for (int i = threadIdx.x; i < k; i += blockDim.x) {
    Key e = ld_key(smem + i);
    const bool ok = !key_is_empty(e);
    ids_out[i] = ok ? ids[e.lo] : -1;
    scores_out[i] = ok ? score_from_rank_hi(e.hi) : __longlong_as_double(0x7ff8000000000000ll);
    found += ok ? 1 : 0;
  }
  if (found) atomicAdd(&s_found, found);
  __syncthreads();
  if (threadIdx.x == 0) *count_out = s_found;
}

// ---- cross-shard exchange over NVLink peer memory ----------------------------------------------------
// Every rank owns a communication buffer that all peers have mapped (cudaIpc): `depth` slots of [world][slot_bytes]
// payload plus one arrival flag per (slot, source rank).  A rank PUSHES its packed result into slot s of every
// peer with plain stores (posted writes over NVLink), fences, and raises its flag there to the exchange's sequence
// number; the merge kernel of each rank waits for the `world` flags of ITS OWN buffer and reads local memory only.
// peer_publish_kernel never waits, so a multi-query exchange depends on nothing but stream order.  The fused
// one-query path (peer_publish_inline below) is different: that kernel publishes and THEN spins on its peers, so
// every rank's merge kernel must be able to run at the same time.  That holds with one GPU per rank; ranks that share
// a device (vs_peer_connect_ptrs with buffers on one GPU) or serialised launches (CUDA_LAUNCH_BLOCKING, a profiler's
// kernel replay) would stall until the time-out trap, so api.cu turns the fused path off for communicators whose
// ranks share a device.
struct PeerBases {
  unsigned char* base[VS_PEER_MAX_WORLD];
};

__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long global_timer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

// Block-wide: returns once flags[0..w) have all reached seq.  A peer that never arrives (a dead process) traps
// after VS_PEER_TIMEOUT_NS instead of hanging the GPU: the error surfaces at the caller's next synchronisation.
__device__ __forceinline__ void peer_wait(const unsigned long long* flags, int w, unsigned long long seq) {
  if (flags == nullptr) return;
  if ((int)threadIdx.x < w) {
    const unsigned long long t0 = global_timer_ns();
    while (ld_acquire_sys(flags + threadIdx.x) < seq) {
      __nanosleep(64);
      if (global_timer_ns() - t0 > VS_PEER_TIMEOUT_NS) __trap();
    }
  }
  __syncthreads();
}

// One query (a single CTA): the merge kernel publishes this rank's list itself before it waits -- one launch less
// on the latency path.  With more CTAs the publishing stays a kernel of its own (nobody may wait before every CTA of
// every rank has published, which only a kernel boundary guarantees without assumptions about CTA scheduling).
struct PeerPublishDev {
  PeerBases peers;
  const uint4* payload;
  size_t n16, data_off, flag_off;  // n16 == 0: nothing to publish
  int rank;
};
__device__ __forceinline__ void peer_publish_inline(const PeerPublishDev& pp, int w, unsigned long long seq) {
  if (pp.n16 == 0) return;
  for (size_t i = threadIdx.x; i < pp.n16; i += blockDim.x) {
    const uint4 v = pp.payload[i];
    for (int p = 0; p < w; p++) reinterpret_cast<uint4*>(pp.peers.base[p] + pp.data_off)[i] = v;
  }
  __threadfence_system();
  __syncthreads();
  if ((int)threadIdx.x < w)
    st_release_sys(reinterpret_cast<unsigned long long*>(pp.peers.base[threadIdx.x] + pp.flag_off) + pp.rank, seq);
}

__global__ void __launch_bounds__(1024)
peer_publish_kernel(PeerBases peers, int w, int rank, const uint4* payload, size_t n16, size_t data_off, size_t flag_off,
                    unsigned long long seq, unsigned int* ticket) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += stride) {
    const uint4 v = payload[i];
    for (int p = 0; p < w; p++) reinterpret_cast<uint4*>(peers.base[p] + data_off)[i] = v;
  }
  __threadfence_system();
  __shared__ int s_last;
  __syncthreads();
  if (threadIdx.x == 0) s_last = (atomicAdd(ticket, 1u) == gridDim.x - 1) ? 1 : 0;
  __syncthreads();
  if (s_last) {  // every CTA's stores are fenced: raise this rank's flag at every peer
    __threadfence_system();
    if ((int)threadIdx.x < w)
      st_release_sys(reinterpret_cast<unsigned long long*>(peers.base[threadIdx.x] + flag_off) + rank, seq);
    if (threadIdx.x == 0) *ticket = 0u;
  }
}

__global__ void __launch_bounds__(RANK_THREADS)
merge_packed_kernel(const int64_t* gath, int w, int nq, int k, int kp, int descending,
                    int64_t* __restrict__ ids_out, double* __restrict__ scores_out,
                    int32_t* __restrict__ counts_out, const unsigned long long* wait_flags, unsigned long long wait_seq,
                    PeerPublishDev pub) {
  extern __shared__ __align__(128) ulonglong2 smem[];
  peer_publish_inline(pub, w, wait_seq);
  peer_wait(wait_flags, w, wait_seq);  // peer exchange: the gathered lists arrive in this rank's own buffer
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int nw = blockDim.x >> 5;
  const int stride_keys = kp + TOPK_BUF;
  const int qi = blockIdx.x;
  const int total = w * k;
  WarpTopK tk;
  tk.init(smem + (size_t)warp * stride_keys, kp, k, lane);
  for (int i0 = warp * 32; i0 < total; i0 += nw * 32) {
    const int i = i0 + lane;
    Key key = key_empty();
    bool ok = false;
    if (i < total) {
      const int r = i / k, j = i % k;
      const int64_t* row = gath + ((size_t)r * nq + qi) * 2 * k;
      const int64_t id = row[j];
      if (id >= 0) {
        const double sc = __longlong_as_double(row[k + j]);
        key = Key{descending ? rank_hi_from_score(sc) : rank_hi_from_dist(sc), (uint64_t)i};
        ok = true;
      }
    }
    tk.push(key, ok, lane);
  }
  tk.flush(lane);
  block_combine_lists(smem, stride_keys, nw, kp, warp, lane);
  __shared__ int s_found;
  if (threadIdx.x == 0) s_found = 0;
  __syncthreads();
  int found = 0;
  for (int i = threadIdx.x; i < k; i += blockDim.x) {
    Key e = ld_key(smem + i);
    const bool ok = !key_is_empty(e);
    int64_t id = -1;
    double sc = __longlong_as_double(0x7ff8000000000000ll);
    if (ok) {
      const int r = (int)(e.lo / k), j = (int)(e.lo % k);
      const int64_t* row = gath + ((size_t)r * nq + qi) * 2 * k;
      id = row[j];
      sc = __longlong_as_double(row[k + j]);
    }
    ids_out[(size_t)qi * k + i] = id;
    scores_out[(size_t)qi * k + i] = sc;
    found += ok ? 1 : 0;
  }
  if (found) atomicAdd(&s_found, found);
  __syncthreads();
  if (threadIdx.x == 0) counts_out[qi] = s_found;
}

// ---- cross-shard ADC + re-rank (one collective per query batch) -------------------------------------------
// The reference re-ranks the GLOBAL first n_cand rows by approximate distance (searchSealedSegment,
// J/fdb/FdbVectorIndex.java:769,820-828); re-ranking each shard's own candidates and merging the winners
// would look at world x n_cand rows and can return different results.  So every shard ships its ADC
// candidates together with their exact scores, packed per query as int64[4][nc]:
//   ids | approximate distance bits | exact score bits | state (1 scored, 0 dropped at re-rank, -1 empty slot)
// and after the all-gather every rank selects the global first nc by (approx, rank, position) and ranks
// those that were scored by (exact score descending, approximate order).
__global__ void __launch_bounds__(1024)
score_pack_kernel(const float* __restrict__ X, int64_t n, int d, const uint8_t* __restrict__ skip, int lanes,
                  const float* __restrict__ Q, const int64_t* __restrict__ cand_ids, const double* __restrict__ cand_approx,
                  const int32_t* __restrict__ cand_counts, int nc, int cosine, int64_t id_base, int foreign_empty,
                  int64_t* __restrict__ pack) {
  const int qi = blockIdx.x;
  const float* q = Q + (size_t)qi * d;
  const int64_t* cand = cand_ids + (size_t)qi * nc;
  const double* appr = cand_approx != nullptr ? cand_approx + (size_t)qi * nc : nullptr;  // null: the position is the key
  int64_t* out = pack + (size_t)qi * 4 * nc;
  const int cnt = cand_counts != nullptr ? cand_counts[qi] : nc;
  const int lane = threadIdx.x & 31;
  const int hl = lane & 15;
  const int hw = threadIdx.x >> 4;
  const int nhw = blockDim.x >> 4;
  const unsigned hmask = (lane < 16) ? 0x0000ffffu : 0xffff0000u;
  const int base_lane = lane & 16;
  double qq = 0.0;
  if (cosine) qq = ref_sum_halfwarp<REF_DOT>(q, q, d, lanes, hl, hmask, base_lane);
  for (int c = hw; c < nc; c += nhw) {
    int64_t g = c < cnt ? cand[c] : -1;
    const int64_t row = g - id_base;
    // caller-supplied candidates re-ranked across shards: an id another shard owns is not this shard's entry
    if (foreign_empty && (row < 0 || row >= n)) g = -1;
    int64_t state = g < 0 ? -1 : 0;
    double score = 0.0;
    bool ok = g >= 0 && row >= 0 && row < n;
    if (ok && skip != nullptr && skip[row]) ok = false;  // deleted / gid missing: takes a slot, is not scored
    if (ok) {                                            // uniform across the half-warp
      const float* x = X + (size_t)row * d;
      if (cosine) {
        const double dot = ref_sum_halfwarp<REF_DOT>(q, x, d, lanes, hl, hmask, base_lane);
        const double xx = ref_sum_halfwarp<REF_DOT>(x, x, d, lanes, hl, hmask, base_lane);
        score = ref_cosine_from_sums(dot, qq, xx);
      } else {
        score = -__dsqrt_rn(ref_sum_halfwarp<REF_L2SQ>(q, x, d, lanes, hl, hmask, base_lane));
      }
      state = 1;
    }
    if (hl == 0) {
      out[c] = g;
      out[nc + c] = g >= 0 ? __double_as_longlong(appr != nullptr ? appr[c] : (double)c) : 0;
      out[2 * nc + c] = __double_as_longlong(score);
      out[3 * nc + c] = state;
    }
  }
}

cudaError_t launch_score_pack(const RankLaunch& L, const double* cand_approx, const int32_t* cand_counts, int64_t* pack,
                              cudaStream_t st, bool foreign_empty) {
  int threads = RANK_THREADS;
  while (threads < 1024 && threads < 16 * L.nc) threads <<= 1;
  score_pack_kernel<<<L.nq, threads, 0, st>>>(L.X, L.n, L.d, L.skip, L.lanes, L.q, L.cand_ids, cand_approx, cand_counts,
                                               L.nc, L.metric == 1, L.id_base, foreign_empty ? 1 : 0, pack);
  count_launch();
  return cudaGetLastError();
}

__global__ void __launch_bounds__(1024)
merge_adc_rerank_kernel(const int64_t* gath, int w, int nq, int nc, int np, int np2, int k,
                        int64_t* __restrict__ ids_out, double* __restrict__ scores_out, int32_t* __restrict__ counts_out,
                        const unsigned long long* wait_flags, unsigned long long wait_seq, PeerPublishDev pub) {
  extern __shared__ __align__(16) ulonglong2 skey[];  // [np] approximate order, then [np2] exact order
  peer_publish_inline(pub, w, wait_seq);
  peer_wait(wait_flags, w, wait_seq);
  ulonglong2* skey2 = skey + np;
  const int qi = blockIdx.x;
  const int total = w * nc;
  auto entry = [&](int p) { return gath + ((size_t)(p / nc) * nq + qi) * 4 * nc + (p % nc); };
  for (int p = threadIdx.x; p < np; p += blockDim.x) {
    Key key = key_empty();
    if (p < total) {
      const int64_t* e = entry(p);
      if (e[3 * nc] >= 0) key = Key{rank_hi_from_dist(__longlong_as_double(e[nc])), (uint64_t)p};
    }
    st_key(skey + p, key);
  }
  __syncthreads();
  block_bitonic_sort_keys(skey, np);  // stable: ties in approx fall back to (rank, position) = ascending global row
  for (int c = threadIdx.x; c < np2; c += blockDim.x) {
    Key key = key_empty();
    if (c < nc) {
      const Key a = ld_key(skey + c);
      if (!key_is_empty(a)) {
        const int64_t* e = entry((int)a.lo);
        if (e[3 * nc] == 1) key = Key{rank_hi_from_score(__longlong_as_double(e[2 * nc])), (uint64_t)c};
      }
    }
    st_key(skey2 + c, key);
  }
  __syncthreads();
  block_bitonic_sort_keys(skey2, np2);
  __shared__ int s_found;
  if (threadIdx.x == 0) s_found = 0;
  __syncthreads();
  int found = 0;
  for (int i = threadIdx.x; i < k; i += blockDim.x) {
    const Key e2 = i < np2 ? ld_key(skey2 + i) : key_empty();
    const bool ok = !key_is_empty(e2);
    int64_t id = -1;
    if (ok) id = entry((int)ld_key(skey + e2.lo).lo)[0];
    ids_out[(size_t)qi * k + i] = id;
    scores_out[(size_t)qi * k + i] = ok ? score_from_rank_hi(e2.hi) : __longlong_as_double(0x7ff8000000000000ll);
    found += ok ? 1 : 0;
  }
  if (found) atomicAdd(&s_found, found);
  __syncthreads();
  if (threadIdx.x == 0) counts_out[qi] = s_found;
}

// ---- graph construction: GraphBuilder.buildL2Neighbors / buildPrunedNeighbors (J/graph/GraphBuilder.java:41-109) ----
// For node i the reference sorts ALL j != i by l2Squared(v_i, v_j) ascending with a stable sort (ties keep the lower j) and
// takes the first L; the pruned variant then walks that list and drops a candidate u when a kept neighbour p has
// l2Squared(v_u, v_p) <= alpha * l2Squared(v_i, v_u).  The O(n^2) part is the brute-force scan (K1 / K2: every row of the
// segment is a query against the segment itself); this kernel turns each row's nominated candidates into the reference's
// list.  The scan orders by sqrt(l2Squared) ("B order"), the reference by l2Squared itself ("A order"): sqrt is monotone
// but not injective in double, so A is recovered by re-scoring the candidates here, and a row whose candidate list may
// have cut a group of equal-sqrt rows short is FLAGGED and redone exactly (knn_exact_kernel).
constexpr int KNN_THREADS = 256;
constexpr int KNN_MAX_DEGREE = 512;

__global__ void __launch_bounds__(KNN_THREADS)
knn_finalize_kernel(const float* __restrict__ X, int64_t n, int d, int lanes, int64_t row0,
                    const int64_t* __restrict__ cand_ids, const double* __restrict__ cand_scores,
                    const int32_t* __restrict__ cand_counts, int kq, int np, int take, int degree, int keep, int check_closure,
                    double alpha, int prune, int64_t id_base, int32_t* __restrict__ neighbors, int32_t* __restrict__ counts,
                    int32_t* __restrict__ flags) {
  // take: candidates considered (min(degree or lBuild, n-1)); keep: most neighbours kept (<= degree); degree: row pitch of the output
  extern __shared__ __align__(16) ulonglong2 skey[];  // [np]
  __shared__ int s_sel[KNN_MAX_DEGREE];
  __shared__ int s_nsel, s_drop, s_flag;
  const int qi = blockIdx.x;
  const int64_t i = row0 + qi;
  const float* xi = X + (size_t)i * d;
  const int64_t* cand = cand_ids + (size_t)qi * kq;
  const int cnt = cand_counts[qi] < kq ? cand_counts[qi] : kq;
  const int lane = threadIdx.x & 31, hl = lane & 15, hw = threadIdx.x >> 4, nhw = blockDim.x >> 4;
  const unsigned hmask = (lane < 16) ? 0x0000ffffu : 0xffff0000u;
  const int base_lane = lane & 16;
  if (threadIdx.x == 0) s_flag = 0;
  for (int c = threadIdx.x; c < np; c += blockDim.x) st_key(skey + c, key_empty());
  __syncthreads();
  for (int c = hw; c < cnt; c += nhw) {
    const int64_t j = cand[c] - id_base;
    if (j < 0 || j >= n || j == i) continue;  // uniform across the half-warp
    const double dd = ref_sum_halfwarp<REF_L2SQ>(xi, X + (size_t)j * d, d, lanes, hl, hmask, base_lane);
    if (hl == 0) st_key(skey + c, Key{rank_hi_from_dist(dd), (uint64_t)j});
  }
  if (check_closure && threadIdx.x == 0) {
    // B order = list order.  t = the score of the take-th entry that is a real neighbour (not the node itself, not a
    // NaN distance); the list is closed when it is exhausted or its last entry is strictly farther than t.
    bool closed = cnt < kq;
    if (!closed) {
      int seen = 0, pt = -1;
      for (int c = 0; c < cnt && pt < 0; c++) {
        const int64_t j = cand[c] - id_base;
        const double sc = cand_scores[(size_t)qi * kq + c];
        if (j == i || sc != sc) continue;
        if (++seen == take) pt = c;
      }
      if (pt >= 0) {
        const double t = cand_scores[(size_t)qi * kq + pt], last = cand_scores[(size_t)qi * kq + cnt - 1];
        closed = last == last && last < t;  // scores are -distance: strictly smaller = strictly farther
      }
    }
    s_flag = closed ? 0 : 1;
  }
  __syncthreads();
  block_bitonic_sort_keys(skey, np);  // (l2Squared, j) ascending, NaN distances last, empty slots at the end
  int total = 0;  // real candidates, capped at `take`
  if (!prune) {
    for (int c = threadIdx.x; c < degree; c += blockDim.x) {
      const Key e = c < np ? ld_key(skey + c) : key_empty();
      neighbors[(size_t)i * degree + c] = (c < keep && !key_is_empty(e)) ? (int32_t)e.lo : -1;
    }
    if (threadIdx.x == 0) {
      while (total < keep && total < np && !key_is_empty(ld_key(skey + total))) total++;
      counts[i] = total;
      if (flags != nullptr) flags[qi] = s_flag;
    }
    return;
  }
  // pruned: greedy in A order over the first `take` (= min(lBuild, n-1)) candidates, at most `degree` kept
  if (threadIdx.x == 0) s_nsel = 0;
  __syncthreads();
  for (int c = 0; c < take && c < np; c++) {
    const Key e = ld_key(skey + c);
    if (key_is_empty(e)) break;                 // uniform: shared memory
    if (s_nsel >= keep) break;
    const int u = (int)e.lo;
    const double diu = dist_from_rank_hi(e.hi);
    if (threadIdx.x == 0) s_drop = 0;
    __syncthreads();
    const int ns = s_nsel;
    const float* xu = X + (size_t)u * d;
    for (int t = hw; t < ns; t += nhw) {
      const double dup = ref_sum_halfwarp<REF_L2SQ>(xu, X + (size_t)s_sel[t] * d, d, lanes, hl, hmask, base_lane);
      if (hl == 0 && dup <= __dmul_rn(alpha, diu)) s_drop = 1;  // :101 (NaN compares false: kept, as in Java)
    }
    __syncthreads();
    if (threadIdx.x == 0 && !s_drop) s_sel[s_nsel++] = u;
    __syncthreads();
  }
  for (int c = threadIdx.x; c < degree; c += blockDim.x) neighbors[(size_t)i * degree + c] = c < s_nsel ? s_sel[c] : -1;
  if (threadIdx.x == 0) {
    counts[i] = s_nsel;
    if (flags != nullptr) flags[qi] = s_flag;
  }
}

// exact A-order candidates of ONE node: every row scored with the reference arithmetic (a half-warp per row), the kq
// smallest (l2Squared, j) kept.  Slow (the whole segment per node) and exact; only flagged nodes come here.
__global__ void __launch_bounds__(KNN_THREADS)
knn_exact_kernel(const float* __restrict__ X, int64_t n, int d, int lanes, const int32_t* __restrict__ nodes, int kq, int kp,
                 int64_t id_base, int64_t* __restrict__ cand_ids, double* __restrict__ cand_scores, int32_t* __restrict__ cand_counts) {
  extern __shared__ __align__(128) ulonglong2 smem[];
  const int qi = blockIdx.x;
  const int64_t i = nodes[qi];
  const float* xi = X + (size_t)i * d;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5, hl = lane & 15;
  const unsigned hmask = (lane < 16) ? 0x0000ffffu : 0xffff0000u;
  const int base_lane = lane & 16;
  const int stride_keys = kp + TOPK_BUF;
  WarpTopK tk;
  tk.init(smem + (size_t)warp * stride_keys, kp, kq, lane);
  for (int64_t j0 = (int64_t)warp * 32; j0 < n; j0 += (int64_t)nw * 32) {
    // 16 rounds x 2 half-warps score 32 consecutive rows; lane l ends up holding the key of row j0 + l
    Key mine = key_empty();
    bool have = false;
    for (int r = 0; r < 16; r++) {
      const int64_t j = j0 + 2 * r + (lane >> 4);  // uniform within a half-warp
      double dd = 0.0;
      if (j < n) dd = ref_sum_halfwarp<REF_L2SQ>(xi, X + (size_t)j * d, d, lanes, hl, hmask, base_lane);
      const double v0 = __shfl_sync(FULL_MASK, dd, 0), v1 = __shfl_sync(FULL_MASK, dd, 16);
      if (lane == 2 * r) {
        mine = Key{rank_hi_from_dist(v0), (uint64_t)(j0 + 2 * r)};
        have = j0 + 2 * r < n && j0 + 2 * r != i;
      }
      if (lane == 2 * r + 1) {
        mine = Key{rank_hi_from_dist(v1), (uint64_t)(j0 + 2 * r + 1)};
        have = j0 + 2 * r + 1 < n && j0 + 2 * r + 1 != i;
      }
    }
    tk.push(mine, have, lane);
  }
  tk.flush(lane);
  block_combine_lists(smem, stride_keys, nw, kp, warp, lane);
  __shared__ int s_found;
  if (threadIdx.x == 0) s_found = 0;
  __syncthreads();
  int found = 0;
  for (int c = threadIdx.x; c < kq; c += blockDim.x) {
    const Key e = ld_key(smem + c);
    const bool ok = !key_is_empty(e);
    cand_ids[(size_t)qi * kq + c] = ok ? id_base + (int64_t)e.lo : -1;
    cand_scores[(size_t)qi * kq + c] = ok ? dist_from_rank_hi(e.hi) : __longlong_as_double(0x7ff8000000000000ll);
    found += ok ? 1 : 0;
  }
  if (found) atomicAdd(&s_found, found);
  __syncthreads();
  if (threadIdx.x == 0) cand_counts[qi] = s_found;
}

cudaError_t launch_knn_finalize(const float* X, int64_t n, int d, int lanes, int64_t row0, int nrows, const int64_t* cand_ids,
                                const double* cand_scores, const int32_t* cand_counts, int kq, int take, int degree, int keep,
                                bool check_closure, double alpha, bool prune, int64_t id_base, int32_t* neighbors,
                                int32_t* counts, int32_t* flags, cudaStream_t st) {
  if (degree > KNN_MAX_DEGREE || keep > degree || kq > RANK_MAX_CAND) return cudaErrorInvalidValue;
  int np = 2;
  while (np < kq) np <<= 1;
  const size_t smem = (size_t)np * 16;
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(knn_finalize_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
  }
  knn_finalize_kernel<<<nrows, KNN_THREADS, smem, st>>>(X, n, d, lanes, row0, cand_ids, cand_scores, cand_counts, kq, np, take,
                                                        degree, keep, check_closure ? 1 : 0, alpha, prune ? 1 : 0, id_base,
                                                        neighbors, counts, flags);
  count_launch();
  return cudaGetLastError();
}

cudaError_t launch_knn_exact(const float* X, int64_t n, int d, int lanes, const int32_t* nodes, int nnodes, int kq, int64_t id_base,
                             int64_t* cand_ids, double* cand_scores, int32_t* cand_counts, cudaStream_t st) {
  const int kp = topk_pad(kq);
  const size_t smem = (size_t)(KNN_THREADS / 32) * topk_warp_smem(kp);
  if (smem > 200 * 1024) return cudaErrorInvalidValue;
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(knn_exact_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
  }
  knn_exact_kernel<<<nnodes, KNN_THREADS, smem, st>>>(X, n, d, lanes, nodes, kq, kp, id_base, cand_ids, cand_scores, cand_counts);
  count_launch();
  return cudaGetLastError();
}

// ---- an empty shard's contribution to an exchange --------------------------------------------------------
// kind 0: brute-force / ADC lists [nq][2k] (ids -1, score bits NaN); kind 1: ADC + re-rank packs [nq][4][k] (state -1)
__global__ void fill_pack_kernel(int64_t* __restrict__ pack, int nq, int k, int kind, int32_t* __restrict__ counts) {
  const int64_t per = kind == 0 ? 2 * (int64_t)k : 4 * (int64_t)k;
  const int64_t total = (int64_t)nq * per;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t j = i % per;
    int64_t v;
    if (kind == 0) v = j < k ? -1 : 0x7ff8000000000000ll;
    else v = (j < k || j >= 3 * (int64_t)k) ? -1 : 0;
    pack[i] = v;
  }
  if (counts != nullptr)
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nq; i += (int64_t)gridDim.x * blockDim.x) counts[i] = 0;
}
cudaError_t launch_fill_pack(int64_t* pack, int nq, int k, int kind, int32_t* counts, cudaStream_t st) {
  const int64_t total = (int64_t)nq * (kind == 0 ? 2 : 4) * k;
  int ctas = (int)((total + 255) / 256);
  ctas = ctas < 1 ? 1 : (ctas > 1024 ? 1024 : ctas);
  fill_pack_kernel<<<ctas, 256, 0, st>>>(pack, nq, k, kind, counts);
  count_launch();
  return cudaGetLastError();
}

// ---- all-reduce over the peer buffers (K10): PqTrainer's per-cluster sums and counts across row shards -------
// Every rank has pushed [nf floats | ni ints] into slot `rank` of every peer (launch_peer_publish).  These kernels
// wait for the arrival flags in this rank's own buffer and combine the `w` copies from local memory:
//   floats  mode 0: fp32 sum in ASCENDING RANK ORDER ((c0 + c1) + c2 ...): deterministic and identical on every rank
//           mode 1: bitwise OR (rows owned by exactly one rank, zero elsewhere: an exact broadcast, -0.0f included)
//           mode 2: the copy of rank w-1 (the exact-order chain has passed through every rank)
//   ints    integer sum.
__global__ void __launch_bounds__(256)
peer_reduce_kernel(const unsigned char* gath, int w, size_t stride, int64_t nf, int64_t ni, int mode,
                   float* __restrict__ out_f, int32_t* __restrict__ out_i, const unsigned long long* flags,
                   unsigned long long seq) {
  peer_wait(flags, w, seq);
  const int64_t i0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x, step = (int64_t)gridDim.x * blockDim.x;
  const size_t ioff = (size_t)((nf + 3) / 4 * 4) * 4;
  for (int64_t i = i0; i < nf; i += step) {
    if (mode == 0) {
      float acc = reinterpret_cast<const float*>(gath)[i];
      for (int r = 1; r < w; r++) acc = __fadd_rn(acc, reinterpret_cast<const float*>(gath + (size_t)r * stride)[i]);
      out_f[i] = acc;
    } else if (mode == 1) {
      unsigned int acc = 0u;
      for (int r = 0; r < w; r++) acc |= reinterpret_cast<const unsigned int*>(gath + (size_t)r * stride)[i];
      out_f[i] = __uint_as_float(acc);
    } else {
      out_f[i] = reinterpret_cast<const float*>(gath + (size_t)(w - 1) * stride)[i];
    }
  }
  for (int64_t i = i0; i < ni; i += step) {
    int32_t acc = 0;
    for (int r = 0; r < w; r++) acc += reinterpret_cast<const int32_t*>(gath + (size_t)r * stride + ioff)[i];
    out_i[i] = acc;
  }
}
cudaError_t launch_peer_reduce(const void* gath, int w, size_t stride, int64_t nf, int64_t ni, int mode, float* out_f,
                               int32_t* out_i, const unsigned long long* flags, unsigned long long seq, cudaStream_t st) {
  if (w < 1 || w > VS_PEER_MAX_WORLD) return cudaErrorInvalidValue;
  const int64_t m = nf > ni ? nf : ni;
  int ctas = (int)((m + 255) / 256);
  ctas = ctas < 1 ? 1 : (ctas > 296 ? 296 : ctas);
  peer_reduce_kernel<<<ctas, 256, 0, st>>>(static_cast<const unsigned char*>(gath), w, stride, nf, ni, mode, out_f, out_i, flags, seq);
  count_launch();
  return cudaGetLastError();
}
// waits (one warp) until the flag of source rank `src` has reached seq: orders the stream behind ONE peer's publish
__global__ void peer_wait_one_kernel(const unsigned long long* flags, int src, unsigned long long seq) {
  if (threadIdx.x == 0) {
    const unsigned long long t0 = global_timer_ns();
    while (ld_acquire_sys(flags + src) < seq) {
      __nanosleep(64);
      if (global_timer_ns() - t0 > VS_PEER_TIMEOUT_NS) __trap();
    }
  }
}
cudaError_t launch_peer_wait_one(const unsigned long long* flags, int src, unsigned long long seq, cudaStream_t st) {
  peer_wait_one_kernel<<<1, 32, 0, st>>>(flags, src, seq);
  count_launch();
  return cudaGetLastError();
}

static PeerPublishDev make_publish(const PeerPublish* pub, int w, int nq) {
  PeerPublishDev pd{};
  if (pub != nullptr && nq == 1 && w <= VS_PEER_MAX_WORLD) {
    for (int p = 0; p < w; p++) pd.peers.base[p] = pub->bases[p];
    pd.payload = static_cast<const uint4*>(pub->payload);
    pd.n16 = pub->bytes / 16;
    pd.data_off = pub->data_off;
    pd.flag_off = pub->flag_off;
    pd.rank = pub->rank;
  }
  return pd;
}

cudaError_t launch_peer_publish(unsigned char* const* bases, int w, int rank, const void* payload, size_t bytes,
                                size_t data_off, size_t flag_off, unsigned long long seq, unsigned int* ticket,
                                cudaStream_t st) {
  if (w < 1 || w > VS_PEER_MAX_WORLD || (bytes & 15) != 0 || (data_off & 15) != 0) return cudaErrorInvalidValue;
  PeerBases pb;
  for (int p = 0; p < VS_PEER_MAX_WORLD; p++) pb.base[p] = p < w ? bases[p] : nullptr;
  const size_t n16 = bytes / 16;
  const int threads = n16 >= 1024 ? 1024 : (n16 > 32 ? (int)((n16 + 31) / 32 * 32) : 32);
  int ctas = (int)((n16 * (size_t)w + 16383) / 16384);  // about 16 stores per thread
  ctas = ctas < 1 ? 1 : (ctas > 16 ? 16 : ctas);
  peer_publish_kernel<<<ctas, threads < w ? 32 : threads, 0, st>>>(pb, w, rank, static_cast<const uint4*>(payload), n16, data_off,
                                                                   flag_off, seq, ticket);
  count_launch();
  return cudaGetLastError();
}

cudaError_t launch_merge_adc_rerank(const int64_t* gath, int w, int nq, int nc, int k, int64_t* ids_out,
                                    double* scores_out, int32_t* counts_out, cudaStream_t st,
                                    const unsigned long long* wait_flags, unsigned long long wait_seq, const PeerPublish* pub) {
  if ((int64_t)w * nc > RANK_MAX_CAND) return cudaErrorInvalidValue;
  int np = 2, np2 = 2;
  while (np < w * nc) np <<= 1;
  while (np2 < nc) np2 <<= 1;
  const size_t smem = (size_t)(np + np2) * 16;
  // largest case: RANK_MAX_CAND gathered entries plus their first TOPK_MAX_K by approximate distance
  cudaError_t e = cudaFuncSetAttribute(merge_adc_rerank_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)((size_t)(RANK_MAX_CAND + TOPK_MAX_K) * 16));
  if (e != cudaSuccess) return e;
  if (smem > (size_t)(RANK_MAX_CAND + TOPK_MAX_K) * 16) return cudaErrorInvalidValue;
  int threads = RANK_THREADS;
  while (threads < 1024 && threads < np / 2) threads <<= 1;
  merge_adc_rerank_kernel<<<nq, threads, smem, st>>>(gath, w, nq, nc, np, np2, k, ids_out, scores_out, counts_out, wait_flags,
                                                     wait_seq, make_publish(pub, w, nq));
  count_launch();
  return cudaGetLastError();
}

cudaError_t launch_merge_packed(const int64_t* gath, int w, int nq, int k, bool descending,
                                int64_t* ids_out, double* scores_out, int32_t* counts_out, cudaStream_t st,
                                const unsigned long long* wait_flags, unsigned long long wait_seq, const PeerPublish* pub) {
  const int kp = topk_pad(k);
  const size_t smem = (size_t)(RANK_THREADS / 32) * topk_warp_smem(kp);
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(merge_packed_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
  }
  merge_packed_kernel<<<nq, RANK_THREADS, smem, st>>>(gath, w, nq, k, kp, descending ? 1 : 0, ids_out, scores_out, counts_out,
                                                      wait_flags, wait_seq, make_publish(pub, w, nq));
  count_launch();
  return cudaGetLastError();
}

cudaError_t launch_rank(const RankLaunch& L, cudaStream_t st) {
  if (L.nc > RANK_MAX_CAND) return cudaErrorInvalidValue;
  int np = 2;
  while (np < L.nc) np <<= 1;
  const size_t smem = (size_t)np * 16;
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(rank_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
  }
  // a half-warp per candidate: as many half-warps as candidates (up to 64) so that the gather of the
  // candidate rows is one or two latency-bound rounds
  int threads = RANK_THREADS;
  while (threads < 1024 && threads < 16 * L.nc) threads <<= 1;
  count_launch();
  return launch_pdl(rank_kernel, dim3(L.nq), dim3(threads), smem, st, L.X, L.n, L.d, L.skip, L.lanes, L.q, L.cand_ids, L.nc, np,
                    L.k, (int)(L.metric == 1), L.id_base, L.ids_out, L.scores_out, L.counts_out);
}

cudaError_t launch_merge(const int64_t* ids, const double* scores, int64_t total, int k,
                         int64_t* ids_out, double* scores_out, int32_t* count_out, cudaStream_t st) {
  const int kp = topk_pad(k);
  const size_t smem = (size_t)(RANK_THREADS / 32) * topk_warp_smem(kp);
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(merge_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
  }
  merge_kernel<<<1, RANK_THREADS, smem, st>>>(ids, scores, total, k, kp, ids_out, scores_out, count_out);
  count_launch();
  return cudaGetLastError();
}

}  // namespace vs
