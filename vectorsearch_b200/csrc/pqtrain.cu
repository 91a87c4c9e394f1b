// pqtrain.cu -- K4: PqTrainer.train (J/pq/PqTrainer.java:28-91) on device-resident rows.
//
// Per subspace the reference runs Lloyd's k-means: K initial centroids drawn WITH replacement by
// rnd.nextInt(n) (:47-50), then `iterations` x { assign (strict-< argmin, :56-68); update:
// float sums `newC[a][d] += x[d]` IN ROW ORDER plus int counts (:70-77); an empty cluster becomes
// a copy of data.get(rnd.nextInt(n)) (:79-82), the others are divided by their count in fp32
// (:84) }.  One java.util.Random(seed) is shared by all subspaces in sequence (:37).
//
// What is parallel here and what is not:
//  * assignment: pq.cu, all subspaces of a wave in one pass over the rows;
//  * update: fp32 addition is not associative, so a tree/atomic reduction would not reproduce the
//    reference's sums.  Rows are therefore stably partitioned by cluster (per-warp-block
//    histograms -> exclusive scan -> in-order scatter) and every (subspace, cluster, component)
//    chain is summed sequentially in row order by its own thread: M*K*subDim independent chains
//    (32 768 for the production shape) keep the GPU busy while each chain's order is the
//    reference's.  Centroids are bit-identical to the JVM's.
//  * the shared Random couples the subspaces only through the NUMBER of draws each one consumes
//    (K for the init + one per empty cluster).  All remaining subspaces run together with their
//    stream positions predicted (duplicate init draws produce exactly one empty cluster each in
//    iteration 1); the prediction is verified afterwards and the subspaces behind the first
//    mismatch are replayed with the corrected position.
#include <algorithm>
#include <cstring>
#include <vector>

#include "../../include/vsgpu.h"
#include "common.cuh"
#include "host.h"
#include "kernels.h"

namespace vs {

namespace {

// ---- java.util.Random on the host (JDK core, fully specified) -----------------------------------------
struct JRandom {
  uint64_t s;
  static constexpr uint64_t MULT = 0x5DEECE66DULL, ADD = 0xBULL, MASK = (1ULL << 48) - 1;
  explicit JRandom(int64_t seed) : s(((uint64_t)seed ^ MULT) & MASK) {}
  int32_t next(int bits) {
    s = (s * MULT + ADD) & MASK;
    return (int32_t)(uint32_t)(s >> (48 - bits));
  }
  int32_t nextInt(int32_t bound) {  // bound > 0
    int32_t r = next(31);
    const int32_t m = bound - 1;
    if ((bound & m) == 0) return (int32_t)(((int64_t)bound * (int64_t)r) >> 31);
    for (int32_t u = r;;) {
      r = u % bound;
      if ((int32_t)((uint32_t)u - (uint32_t)r + (uint32_t)m) >= 0) return r;
      u = next(31);
    }
  }
};

constexpr int RB = 2048;  // rows per warp-block of the stable partition

// centroids[s][ci] <- sub-vector s of row rows[(s - s_begin) * K + ci]   (rows < 0: leave as is)
__global__ void gather_subvectors_kernel(const float* __restrict__ X, int d, int K, int sd,
                                         const int64_t* __restrict__ rows, float* __restrict__ centroids,
                                         int s_begin, int s_end) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t total = (int64_t)(s_end - s_begin) * K * sd;
  if (t >= total) return;
  const int comp = (int)(t % sd);
  const int64_t e = t / sd;
  const int ci = (int)(e % K);
  const int s = s_begin + (int)(e / K);
  const int64_t row = rows[e];
  if (row < 0) return;
  centroids[((size_t)s * K + ci) * sd + comp] = X[(size_t)row * d + (size_t)s * sd + comp];
}

// Sharded training: out[sl][ci][comp] <- sub-vector of GLOBAL row rows[sl * K + ci] if this rank owns it
// (row_lo <= row < row_lo + n), else untouched (the buffer is zeroed first and summed across ranks)
__global__ void gather_owned_kernel(const float* __restrict__ X, int64_t n, int d, int K, int sd, int64_t row_lo,
                                    const int64_t* __restrict__ rows, float* __restrict__ out, int s_begin, int ns) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (int64_t)ns * K * sd) return;
  const int comp = (int)(t % sd);
  const int64_t e = t / sd;
  const int s = s_begin + (int)(e / K);
  const int64_t row = rows[e] - row_lo;
  if (rows[e] < 0 || row < 0 || row >= n) return;
  out[t] = X[(size_t)row * d + (size_t)s * sd + comp];
}

// one warp per (row block, subspace): histogram of assignments
__global__ void __launch_bounds__(32)
hist_kernel(const int32_t* __restrict__ assign, int64_t n, int K, int nb, int s_begin,
            int32_t* __restrict__ blockhist) {
  extern __shared__ int32_t h[];
  const int b = blockIdx.x;
  const int sl = blockIdx.y;
  const int lane = threadIdx.x;
  for (int i = lane; i < K; i += 32) h[i] = 0;
  __syncwarp();
  const int32_t* a = assign + (size_t)(s_begin + sl) * n;
  const int64_t r0 = (int64_t)b * RB;
  const int64_t r1 = r0 + RB < n ? r0 + RB : n;
  for (int64_t r = r0 + lane; r < r1; r += 32) atomicAdd(&h[a[r]], 1);
  __syncwarp();
  int32_t* out = blockhist + ((size_t)sl * nb + b) * K;
  for (int i = lane; i < K; i += 32) out[i] = h[i];
}

// thread per (subspace, cluster): exclusive prefix over row blocks, in place; totals -> counts
__global__ void scan_blocks_kernel(int32_t* __restrict__ blockhist, int K, int nb, int ns,
                                   int32_t* __restrict__ counts /* [ns][K] */) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= ns * K) return;
  const int sl = t / K, ci = t % K;
  int32_t run = 0;
  int32_t* p = blockhist + (size_t)sl * nb * K + ci;
  for (int b = 0; b < nb; b++) {
    const int32_t v = p[(size_t)b * K];
    p[(size_t)b * K] = run;
    run += v;
  }
  counts[t] = run;
}

// one thread per subspace: offsets[sl][0..K] = exclusive scan of counts
__global__ void scan_clusters_kernel(const int32_t* __restrict__ counts, int K, int ns,
                                     int32_t* __restrict__ offsets /* [ns][K+1] */) {
  const int sl = blockIdx.x * blockDim.x + threadIdx.x;
  if (sl >= ns) return;
  int32_t run = 0;
  for (int ci = 0; ci < K; ci++) {
    offsets[(size_t)sl * (K + 1) + ci] = run;
    run += counts[(size_t)sl * K + ci];
  }
  offsets[(size_t)sl * (K + 1) + K] = run;
}

// one warp per (row block, subspace): rows visited in ascending order, each written to its
// cluster's next free slot -> order[sl][...] lists every cluster's members in row order
__global__ void __launch_bounds__(32)
scatter_kernel(const int32_t* __restrict__ assign, int64_t n, int K, int nb, int s_begin,
               const int32_t* __restrict__ blockprefix, const int32_t* __restrict__ offsets,
               int32_t* __restrict__ order) {
  extern __shared__ int32_t pos[];
  const int b = blockIdx.x;
  const int sl = blockIdx.y;
  const int lane = threadIdx.x;
  const int32_t* bp = blockprefix + ((size_t)sl * nb + b) * K;
  const int32_t* off = offsets + (size_t)sl * (K + 1);
  for (int i = lane; i < K; i += 32) pos[i] = off[i] + bp[i];
  __syncwarp();
  const int32_t* a = assign + (size_t)(s_begin + sl) * n;
  int32_t* ord = order + (size_t)sl * n;
  const int64_t r0 = (int64_t)b * RB;
  const int64_t r1 = r0 + RB < n ? r0 + RB : n;
  for (int64_t rb = r0; rb < r1; rb += 32) {
    const int64_t r = rb + lane;
    const bool live = r < r1;
    const int32_t ci = live ? a[r] : -1 - lane;  // dead lanes get unique keys
    const unsigned peers = __match_any_sync(FULL_MASK, ci);
    const int rank = __popc(peers & ((1u << lane) - 1u));
    int32_t base = 0;
    if (live) base = pos[ci];
    __syncwarp();
    if (live) {
      ord[base + rank] = (int32_t)r;
      if (rank == 0) pos[ci] = base + __popc(peers);
    }
    __syncwarp();
  }
}

// thread per (subspace, cluster, component): sequential fp32 sum over the cluster's rows in row order
__global__ void __launch_bounds__(64)
chain_sum_kernel(const float* __restrict__ X, int64_t n, int d, int K, int sd, int s_begin, int ns,
                 const int32_t* __restrict__ order, const int32_t* __restrict__ offsets,
                 const float* __restrict__ carry /* nullable: running sums of the lower row ranges */,
                 float* __restrict__ sums /* [ns][K][sd] */) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (int64_t)ns * K * sd) return;
  const int comp = (int)(t % sd);
  const int64_t e = t / sd;
  const int ci = (int)(e % K);
  const int sl = (int)(e / K);
  const int32_t* ord = order + (size_t)sl * n;
  const int32_t j0 = offsets[(size_t)sl * (K + 1) + ci];
  const int32_t j1 = offsets[(size_t)sl * (K + 1) + ci + 1];
  const float* col = X + (size_t)(s_begin + sl) * sd + comp;
  float acc = carry != nullptr ? carry[t] : 0.0f;
  int32_t j = j0;
  // The additions are a dependent chain (that IS the reference's order); the loads are not.  The kernel is bound by how
  // many row fetches are in flight (32 768 threads over 148 SMs is seven warps per SM), so every thread keeps CH_U gathers
  // outstanding and fetches the NEXT batch's row indices before it adds the current one.
  constexpr int CH_U = 32;
  int32_t idx[CH_U];
  if (j + CH_U <= j1) {
#pragma unroll
    for (int u = 0; u < CH_U; u++) idx[u] = ord[j + u];
  }
  for (; j + CH_U <= j1; j += CH_U) {
    float v[CH_U];
#pragma unroll
    for (int u = 0; u < CH_U; u++) v[u] = __ldg(col + (size_t)idx[u] * d);
    if (j + 2 * CH_U <= j1) {
#pragma unroll
      for (int u = 0; u < CH_U; u++) idx[u] = ord[j + CH_U + u];
    }
#pragma unroll
    for (int u = 0; u < CH_U; u++) acc = __fadd_rn(acc, v[u]);
  }
  for (; j + 8 <= j1; j += 8) {
    float v[8];
#pragma unroll
    for (int u = 0; u < 8; u++) v[u] = __ldg(col + (size_t)ord[j + u] * d);
#pragma unroll
    for (int u = 0; u < 8; u++) acc = __fadd_rn(acc, v[u]);
  }
  for (; j < j1; j++) acc = __fadd_rn(acc, __ldg(col + (size_t)ord[j] * d));
  sums[t] = acc;
}

// centroid = count ? sum / (float)count : sub-vector of the re-init row   (PqTrainer.java:78-87)
__global__ void finalize_kernel(const float* __restrict__ X, int d, int K, int sd, int s_begin, int ns,
                                const float* __restrict__ sums, const int32_t* __restrict__ counts,
                                const int64_t* __restrict__ reinit, const float* __restrict__ reinit_vals,
                                float* __restrict__ centroids) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (int64_t)ns * K * sd) return;
  const int comp = (int)(t % sd);
  const int64_t e = t / sd;
  const int ci = (int)(e % K);
  const int sl = (int)(e / K);
  const int s = s_begin + sl;
  const int32_t c = counts[e];
  float v;
  if (c == 0)  // reinit_vals: the sub-vector was fetched from the rank that owns the row (sharded training)
    v = reinit_vals != nullptr ? reinit_vals[t] : X[(size_t)reinit[e] * d + (size_t)s * sd + comp];
  else
    v = __fdiv_rn(sums[t], (float)c);
  centroids[((size_t)s * K + ci) * sd + comp] = v;
}

// Scratch of a training run comes from the stream-ordered pool the tensor-core assignment uses (pq_tc.cu): allocation and
// release are ordered on the stream, cached between calls -- a training call neither reaches the driver's allocator nor
// synchronises the device on its way out (eight cudaMalloc + cudaFree pairs were ~15 % of a 20 ms run on 8 GPUs).
struct DevBuf {
  void* p = nullptr;
  cudaStream_t st = nullptr;
  cudaError_t alloc(size_t bytes, cudaStream_t s) {
    st = s;
    return pq_pool_alloc(&p, bytes ? bytes : 1, s);
  }
  ~DevBuf() {
    if (p) cudaFreeAsync(p, st);
  }
  template <typename T>
  T* as() { return static_cast<T*>(p); }
};

#define TCK(call, what)                                 \
  do {                                                  \
    cudaError_t _e = (call);                            \
    if (_e != cudaSuccess) return cuda_fail(_e, what);  \
  } while (0)

}  // namespace

// comm == nullptr: one process owns all n rows (bit-identical to the reference).  Otherwise this process owns
// rows [comm->row_lo, comm->row_lo + n) of comm->n_total and the per-cluster sums and counts are combined
// across the ranks every iteration through comm->allreduce: in row order, rank after rank (exact_order:
// centroids bit-identical to the reference on any number of ranks), or with ONE all-reduce (fp32 additions
// re-associated across shards: a different, statistically equivalent Lloyd trajectory -- k-means amplifies
// the last-bit difference as soon as one row changes cluster).  Rows another rank owns (initial centroids,
// re-initialised empty clusters) arrive through a zero-padded sum.  Every rank takes the same decisions:
// counts and random draws are global.
int pq_train_device(cudaStream_t st, const float* dX, int64_t n, int d, int M, int K, int iterations,
                    int64_t seed, int lanes_, float* centroids_out, const TrainComm* comm) {
  const int sd = d / M;
  const int64_t n_draw = comm ? comm->n_total : n;  // bound of rnd.nextInt(n)
  auto reduce = [&](int kind, int64_t count) -> int {
    cudaError_t e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) return cuda_fail(e, "sync before all-reduce");
    if (comm->allreduce(comm->user, kind, count) != 0) return fail(VS_ESTATE, "all-reduce callback failed");
    return VS_OK;
  };
  // Peer transport (comm->peer != nullptr): the exchange is the library's own -- every rank pushes
  // [floats | ints] into all peers' buffers over NVLink and a reduce kernel waits for the arrival flags.  Nothing
  // leaves the stream: no callback, no host synchronisation per reduction.
  const bool peer = comm != nullptr && comm->peer != nullptr;
  const size_t max_f = (size_t)M * K * sd, max_i = (size_t)M * K;
  auto ioff_of = [](size_t nf) { return ((nf + 3) / 4 * 4) * 4; };
  auto bytes_of = [&](size_t nf, size_t ni) { return (ioff_of(nf) + ni * 4 + 15) / 16 * 16; };
  DevBuf b_send, b_reinit;
  if (peer) {
    if (b_send.alloc(bytes_of(max_f, max_i), st) != cudaSuccess || b_reinit.alloc(max_f * 4, st) != cudaSuccess) {
      cudaGetLastError();
      return fail(VS_ENOMEM, "cudaMalloc(exchange buffers)");
    }
  }
  float* d_send_f = b_send.as<float>();
  // rows owned by exactly one rank (initial centroids, re-initialised clusters) -> the same values on every rank
  auto peer_bcast_rows = [&](const int64_t* d_rows_, int s_begin, int ns_, float* out) -> int {
    const int64_t total = (int64_t)ns_ * K * sd;
    PeerXchg x;
    int r_ = comm->peer_begin(comm->peer, st, bytes_of((size_t)total, 0), &x);
    if (r_ != VS_OK) return r_;
    cudaError_t e = cudaMemsetAsync(d_send_f, 0, bytes_of((size_t)total, 0), st);
    if (e != cudaSuccess) return cuda_fail(e, "memset");
    gather_owned_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(dX, n, d, K, sd, comm->row_lo, d_rows_, d_send_f, s_begin, ns_);
    count_launch();
    if ((e = cudaGetLastError()) != cudaSuccess) return cuda_fail(e, "gather launch");
    if ((r_ = comm->peer_publish(comm->peer, st, d_send_f, &x)) != VS_OK) return r_;
    int enq = 0;
    if ((r_ = comm->peer_wait(comm->peer, st, &x, -1, &enq)) != VS_OK) return r_;
    if ((e = launch_peer_reduce(x.gath, comm->world, x.stride, total, 0, 1, out, nullptr, enq ? nullptr : x.flags, x.seq, st)) != cudaSuccess)
      return cuda_fail(e, "peer reduce launch");
    return VS_OK;
  };
  if ((size_t)K * 4 > 96 * 1024) return fail(VS_EINVAL, "K too large for the device trainer (max 24576)");
  const int nb = (int)((n + RB - 1) / RB);
  DevBuf b_cent, b_assign, b_order, b_hist, b_counts, b_offsets, b_sums, b_rows;
  TCK(b_cent.alloc((size_t)M * K * sd * 4, st), "cudaMalloc(centroids)");
  TCK(b_assign.alloc((size_t)M * n * 4, st), "cudaMalloc(assign)");
  TCK(b_order.alloc((size_t)M * n * 4, st), "cudaMalloc(order)");
  TCK(b_hist.alloc((size_t)M * nb * K * 4, st), "cudaMalloc(blockhist)");
  TCK(b_counts.alloc((size_t)M * K * 4, st), "cudaMalloc(counts)");
  TCK(b_offsets.alloc((size_t)M * (K + 1) * 4, st), "cudaMalloc(offsets)");
  TCK(b_sums.alloc((size_t)M * K * sd * 4, st), "cudaMalloc(sums)");
  TCK(b_rows.alloc((size_t)M * K * 8, st), "cudaMalloc(rows)");
  float* d_cent = b_cent.as<float>();
  int32_t* d_assign = b_assign.as<int32_t>();
  int32_t* d_order = b_order.as<int32_t>();
  int32_t* d_hist = b_hist.as<int32_t>();
  int32_t* d_counts = b_counts.as<int32_t>();
  int32_t* d_offsets = b_offsets.as<int32_t>();
  float* d_sums = b_sums.as<float>();
  int64_t* d_rows = b_rows.as<int64_t>();

  const size_t ksmem = (size_t)K * 4;
  if (ksmem > 48 * 1024) {
    TCK(cudaFuncSetAttribute(hist_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ksmem), "smem attr");
    TCK(cudaFuncSetAttribute(scatter_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ksmem), "smem attr");
  }

  struct TcScopeGuard {  // the rows do not change during training: their tensor-core operand image is built once
    TcScopeGuard() { pq_tc_scope_begin(); }
    ~TcScopeGuard() { pq_tc_scope_end(); }
  } tc_scope_guard;
  JRandom rnd(seed);
  std::vector<int64_t> h_rows((size_t)M * K);
  std::vector<int32_t> h_counts((size_t)M * K);
  std::vector<JRandom> cur(M, JRandom(0));
  std::vector<int> predicted(M), actual(M);

  int s0 = 0;
  while (s0 < M) {
    const int ns = M - s0;
    // ---- plan the wave: init draws at predicted stream positions --------------------------------
    JRandom r = rnd;
    for (int s = s0; s < M; s++) {
      int64_t* rows = h_rows.data() + (size_t)(s - s0) * K;
      for (int ci = 0; ci < K; ci++) rows[ci] = r.nextInt((int32_t)n_draw);  // PqTrainer.java:48
      cur[s] = r;
      int dups = 0;
      if (iterations > 0) {
        std::vector<int64_t> sorted(rows, rows + K);
        std::sort(sorted.begin(), sorted.end());
        for (int i = 1; i < K; i++) dups += sorted[i] == sorted[i - 1];
      }
      predicted[s] = dups;
      actual[s] = 0;
      for (int e = 0; e < dups; e++) r.nextInt((int32_t)n_draw);
    }
    TCK(cudaMemcpyAsync(d_rows, h_rows.data(), (size_t)ns * K * 8, cudaMemcpyHostToDevice, st), "H2D init rows");
    {
      const int64_t total = (int64_t)ns * K * sd;
      if (peer) {
        int r_ = peer_bcast_rows(d_rows, s0, ns, d_cent + (size_t)s0 * K * sd);
        if (r_ != VS_OK) return r_;
      } else if (comm) {
        TCK(cudaMemsetAsync(comm->d_f32, 0, (size_t)total * 4, st), "memset");
        gather_owned_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(dX, n, d, K, sd, comm->row_lo, d_rows, comm->d_f32, s0, ns);
        count_launch();
        TCK(cudaGetLastError(), "gather launch");
        { int r_ = reduce(0, total); if (r_ != VS_OK) return r_; }
        TCK(cudaMemcpyAsync(d_cent + (size_t)s0 * K * sd, comm->d_f32, (size_t)total * 4, cudaMemcpyDeviceToDevice, st), "D2D centroids");
      } else {
        gather_subvectors_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(dX, d, K, sd, d_rows, d_cent, s0, M);
        count_launch();
        TCK(cudaGetLastError(), "gather launch");
      }
    }
    TCK(cudaStreamSynchronize(st), "sync");  // h_rows is reused below

    for (int it = 0; it < iterations; it++) {
      PqAssignLaunch L{};
      L.X = dX; L.n = n; L.d = d; L.M = M; L.K = K; L.subDim = sd; L.centroids = d_cent; L.lanes = lanes_;
      L.codes_u8 = nullptr; L.assign_i32 = d_assign; L.s_begin = s0; L.s_end = M;
      TCK(launch_pq_assign(L, st), "pq assign launch");
      hist_kernel<<<dim3(nb, ns), 32, ksmem, st>>>(d_assign, n, K, nb, s0, d_hist);
      count_launch();
      TCK(cudaGetLastError(), "hist launch");
      scan_blocks_kernel<<<(ns * K + 127) / 128, 128, 0, st>>>(d_hist, K, nb, ns, d_counts);
      count_launch();
      scan_clusters_kernel<<<(ns + 31) / 32, 32, 0, st>>>(d_counts, K, ns, d_offsets);
      count_launch();
      scatter_kernel<<<dim3(nb, ns), 32, ksmem, st>>>(d_assign, n, K, nb, s0, d_hist, d_offsets, d_order);
      count_launch();
      TCK(cudaGetLastError(), "scatter launch");
      const int64_t total = (int64_t)ns * K * sd;
      if (peer) {
        // ONE exchange per iteration carries sums and counts.  exact_order: the running sums pass from rank to rank
        // in ascending row order inside that exchange -- rank r waits (on the device) for rank r-1's copy, continues
        // its chains from there and publishes; the copy of the last rank is the reference's sum, bit for bit.
        const size_t ioff = ioff_of((size_t)total);
        PeerXchg x;
        { int r_ = comm->peer_begin(comm->peer, st, bytes_of((size_t)total, (size_t)ns * K), &x); if (r_ != VS_OK) return r_; }
        const float* carry = nullptr;
        if (comm->exact_order && comm->rank > 0) {
          int enq1 = 0;
          { int r_ = comm->peer_wait(comm->peer, st, &x, comm->rank - 1, &enq1); if (r_ != VS_OK) return r_; }
          if (!enq1) TCK(launch_peer_wait_one(x.flags, comm->rank - 1, x.seq, st), "peer wait launch");
          carry = reinterpret_cast<const float*>(x.gath + (size_t)(comm->rank - 1) * x.stride);
        }
        chain_sum_kernel<<<(unsigned)((total + 63) / 64), 64, 0, st>>>(dX, n, d, K, sd, s0, ns, d_order, d_offsets, carry, d_send_f);
        count_launch();
        TCK(cudaGetLastError(), "chain launch");
        TCK(cudaMemcpyAsync(reinterpret_cast<char*>(d_send_f) + ioff, d_counts, (size_t)ns * K * 4, cudaMemcpyDeviceToDevice, st), "D2D counts");
        { int r_ = comm->peer_publish(comm->peer, st, d_send_f, &x); if (r_ != VS_OK) return r_; }
        int enq = 0;
        { int r_ = comm->peer_wait(comm->peer, st, &x, -1, &enq); if (r_ != VS_OK) return r_; }
        TCK(launch_peer_reduce(x.gath, comm->world, x.stride, total, (int64_t)ns * K, comm->exact_order ? 2 : 0, d_sums, d_counts,
                               enq ? nullptr : x.flags, x.seq, st), "peer reduce launch");
      } else if (comm && comm->exact_order) {
        // The reference adds the rows of a cluster in ascending row order.  Shards are ascending row ranges, so
        // rank r continues the running sums of ranks 0..r-1: world rounds, in each of which exactly one rank
        // contributes (the others add zeros, x + 0.0f == x) -- the sums, hence the centroids, are the reference's bit for bit.
        for (int r = 0; r < comm->world; r++) {
          if (r == comm->rank) {
            chain_sum_kernel<<<(unsigned)((total + 63) / 64), 64, 0, st>>>(dX, n, d, K, sd, s0, ns, d_order, d_offsets,
                                                                              r == 0 ? nullptr : comm->d_f32, d_sums);
            count_launch();
            TCK(cudaGetLastError(), "chain launch");
            TCK(cudaMemcpyAsync(comm->d_f32, d_sums, (size_t)total * 4, cudaMemcpyDeviceToDevice, st), "D2D sums");
          } else {
            TCK(cudaMemsetAsync(comm->d_f32, 0, (size_t)total * 4, st), "memset");
          }
          { int r_ = reduce(0, total); if (r_ != VS_OK) return r_; }
        }
        TCK(cudaMemcpyAsync(d_sums, comm->d_f32, (size_t)total * 4, cudaMemcpyDeviceToDevice, st), "D2D sums");
        TCK(cudaMemcpyAsync(comm->d_i32, d_counts, (size_t)ns * K * 4, cudaMemcpyDeviceToDevice, st), "D2D counts");
        { int r_ = reduce(1, (int64_t)ns * K); if (r_ != VS_OK) return r_; }
        TCK(cudaMemcpyAsync(d_counts, comm->d_i32, (size_t)ns * K * 4, cudaMemcpyDeviceToDevice, st), "D2D counts");
      } else {
      chain_sum_kernel<<<(unsigned)((total + 63) / 64), 64, 0, st>>>(dX, n, d, K, sd, s0, ns, d_order, d_offsets, nullptr, d_sums);
      count_launch();
      TCK(cudaGetLastError(), "chain launch");
      if (comm) {  // per-cluster sums and counts of all shards, one all-reduce each (sums re-associated across shards)
        TCK(cudaMemcpyAsync(comm->d_f32, d_sums, (size_t)total * 4, cudaMemcpyDeviceToDevice, st), "D2D sums");
        TCK(cudaMemcpyAsync(comm->d_i32, d_counts, (size_t)ns * K * 4, cudaMemcpyDeviceToDevice, st), "D2D counts");
        { int r_ = reduce(0, total); if (r_ != VS_OK) return r_; }
        { int r_ = reduce(1, (int64_t)ns * K); if (r_ != VS_OK) return r_; }
        TCK(cudaMemcpyAsync(d_sums, comm->d_f32, (size_t)total * 4, cudaMemcpyDeviceToDevice, st), "D2D sums");
        TCK(cudaMemcpyAsync(d_counts, comm->d_i32, (size_t)ns * K * 4, cudaMemcpyDeviceToDevice, st), "D2D counts");
      }
      }
      TCK(cudaMemcpyAsync(h_counts.data(), d_counts, (size_t)ns * K * 4, cudaMemcpyDeviceToHost, st), "D2H counts");
      TCK(cudaStreamSynchronize(st), "sync");
      bool any_empty = false;
      for (int s = s0; s < M; s++) {
        for (int ci = 0; ci < K; ci++) {
          const size_t e = (size_t)(s - s0) * K + ci;
          if (h_counts[e] == 0) {
            h_rows[e] = cur[s].nextInt((int32_t)n_draw);  // PqTrainer.java:81
            any_empty = true;
            actual[s]++;
          } else {
            h_rows[e] = -1;
          }
        }
      }
      TCK(cudaMemcpyAsync(d_rows, h_rows.data(), (size_t)ns * K * 8, cudaMemcpyHostToDevice, st), "H2D reinit rows");
      const float* reinit_vals = nullptr;
      if (peer) {
        if (any_empty) {
          int r_ = peer_bcast_rows(d_rows, s0, ns, b_reinit.as<float>());
          if (r_ != VS_OK) return r_;
        }
        reinit_vals = b_reinit.as<float>();
      } else if (comm && any_empty) {  // the re-init rows may live on other ranks (every rank sees the same empties)
        TCK(cudaMemsetAsync(comm->d_f32, 0, (size_t)total * 4, st), "memset");
        gather_owned_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(dX, n, d, K, sd, comm->row_lo, d_rows, comm->d_f32, s0, ns);
        count_launch();
        TCK(cudaGetLastError(), "gather launch");
        { int r_ = reduce(0, total); if (r_ != VS_OK) return r_; }
        reinit_vals = comm->d_f32;
      } else if (comm) {
        reinit_vals = comm->d_f32;  // unused: no cluster is empty
      }
      finalize_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(dX, d, K, sd, s0, ns, d_sums, d_counts, d_rows, reinit_vals, d_cent);
      count_launch();
      TCK(cudaGetLastError(), "finalize launch");
      TCK(cudaStreamSynchronize(st), "sync");
    }
    // ---- verify the predicted stream positions -------------------------------------------------------
    int s = s0;
    for (; s < M; s++)
      if (actual[s] != predicted[s]) break;
    if (s >= M) break;   // every subspace of the wave started where the reference would have
    rnd = cur[s];        // subspace s itself is right; everything behind it is replayed
    s0 = s + 1;
  }
  TCK(cudaMemcpyAsync(centroids_out, d_cent, (size_t)M * K * sd * 4, cudaMemcpyDeviceToHost, st), "D2H centroids");
  TCK(cudaStreamSynchronize(st), "sync");
  return VS_OK;
}

}  // namespace vs
