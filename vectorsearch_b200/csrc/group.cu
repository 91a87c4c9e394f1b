// group.cu -- one process, several GPUs (vs_init_multi): the in-library coordinator.
//
// The reference's caller is ONE JVM that fans a query out over its segments with CompletableFuture.allOf and merges
// the per-segment lists (J/fdb/FdbVectorIndex.java:418-437).  The GPU counterpart keeps that shape inside the library:
// vs_init_multi binds the GPUs of the box, every GPU gets one worker thread (its own CUDA stream, scratch and staging
// through the ordinary per-thread context of api.cu) and one rank of an in-process peer communicator (vs_peer_*: the
// buffers are plain device pointers here, peer access enabled once -- no IPC, no NCCL, no second process).
//   vs_segment_upload / _generate   split the rows by ascending range over the GPUs (a "sharded" handle)
//   vs_bruteforce_topk, vs_adc_topk, vs_adc_rerank_topk, vs_rerank_topk
//                                   every worker runs the one-call exchange form on its shard (scan -> push into all peers'
//                                   buffers over NVLink -> merge); rank 0's merged result is the caller's
//   vs_pq_train                     vs_pq_train_sharded_peer on every worker (all-reduce of sums and counts over the peer buffers)
//   vs_pq_encode_batch, vs_segment_attach_pq, vs_adc_gather      embarrassingly parallel per shard
// Results equal the single-GPU results bit for bit: shards are ascending row ranges, every merge is the reference's
// stable merge with the shards in the role of segments.
//
// A job is pushed to ALL workers under one lock, so every worker sees the jobs -- hence the collectives -- in the same
// order; calls from several request threads are serialised per group (one lane).
#include <algorithm>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <cstring>
#include <deque>
#include <functional>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <unordered_map>
#include <vector>

#include "../../include/vsgpu.h"
#include "host.h"

namespace vs {

int fail(int code, const char* fmt, ...);

namespace {

struct Job {
  std::function<int(int)> fn;
  std::vector<int> rc;
  std::vector<std::string> err;
  std::atomic<int> pending{0};
  std::mutex mu;
  std::condition_variable cv;
};

struct Worker {
  int slot = 0, dev = 0;
  std::thread th;
  std::mutex mu;
  std::condition_variable cv;
  std::deque<std::shared_ptr<Job>> q;
  std::atomic<int> queued{0};
  bool stop = false;
  uint64_t comm = 0;
};

struct Sharded {
  int64_t n = 0, id_base = 0;
  int d = 0, M = 0, K = 0;
  std::vector<uint64_t> h;   // per-slot segment handles
  std::vector<int64_t> lo;   // world + 1 row boundaries
};

struct Group {
  int n = 0;
  std::vector<std::unique_ptr<Worker>> w;
  std::mutex submit_mu;
  std::mutex reg_mu;
  std::unordered_map<uint64_t, std::shared_ptr<Sharded>> segs;
  uint64_t next = 1;
};

std::mutex g_group_mu;
std::shared_ptr<Group> g_group;
thread_local bool t_is_worker = false;
constexpr uint64_t SHARDED_TAG = 0x5348000000000000ull;  // "SH": sharded handles never collide with plain ones
constexpr int64_t GROUP_SLOT_BYTES = 2 << 20;

std::shared_ptr<Group> group_get() {
  std::lock_guard<std::mutex> g(g_group_mu);
  return g_group;
}

void worker_main(Worker* w) {
  t_is_worker = true;
  set_thread_device(w->dev);
  cudaSetDevice(w->dev);
  for (;;) {
    std::shared_ptr<Job> job;
    // queries arrive back to back: spin briefly before sleeping on the condition variable
    for (int spin = 0; spin < 2000 && w->queued.load(std::memory_order_acquire) == 0; spin++) {
#if defined(__x86_64__)
      __builtin_ia32_pause();
#endif
    }
    {
      std::unique_lock<std::mutex> lk(w->mu);
      w->cv.wait(lk, [&] { return w->stop || !w->q.empty(); });
      if (w->q.empty()) return;  // stop
      job = std::move(w->q.front());
      w->q.pop_front();
      w->queued.fetch_sub(1, std::memory_order_relaxed);
    }
    set_error("");
    const int rc = job->fn(w->slot);
    job->rc[w->slot] = rc;
    if (rc != VS_OK) job->err[w->slot] = vs_last_error();
    if (job->pending.fetch_sub(1, std::memory_order_acq_rel) == 1) {
      std::lock_guard<std::mutex> lk(job->mu);
      job->cv.notify_all();
    }
  }
}

// runs fn(slot) on every worker; returns the first failure (its message becomes the caller's vs_last_error)
int fan_out(const std::shared_ptr<Group>& G, std::function<int(int)> fn) {
  auto job = std::make_shared<Job>();
  job->fn = std::move(fn);
  job->rc.assign(G->n, VS_OK);
  job->err.resize(G->n);
  job->pending.store(G->n);
  {
    std::lock_guard<std::mutex> sg(G->submit_mu);
    for (auto& w : G->w) {
      {
        std::lock_guard<std::mutex> lk(w->mu);
        w->q.push_back(job);
        w->queued.fetch_add(1, std::memory_order_release);
      }
      w->cv.notify_one();
    }
  }
  for (int spin = 0; spin < 4000 && job->pending.load(std::memory_order_acquire) != 0; spin++) {
#if defined(__x86_64__)
    __builtin_ia32_pause();
#endif
  }
  if (job->pending.load(std::memory_order_acquire) != 0) {
    std::unique_lock<std::mutex> lk(job->mu);
    job->cv.wait(lk, [&] { return job->pending.load(std::memory_order_acquire) == 0; });
  }
  for (int i = 0; i < G->n; i++)
    if (job->rc[i] != VS_OK) {
      set_error((std::string("GPU slot ") + std::to_string(i) + ": " + job->err[i]).c_str());
      return job->rc[i];
    }
  return VS_OK;
}

std::shared_ptr<Sharded> sharded_lookup(const std::shared_ptr<Group>& G, uint64_t h) {
  if (!G) return nullptr;
  std::lock_guard<std::mutex> g(G->reg_mu);
  auto it = G->segs.find(h);
  return it == G->segs.end() ? nullptr : it->second;
}

std::vector<int64_t> boundaries(int64_t n, int world) {
  std::vector<int64_t> lo(world + 1);
  const int64_t base = n / world, rem = n % world;
  for (int r = 0; r <= world; r++) lo[r] = r * base + (r < rem ? r : rem);
  return lo;
}

uint64_t sharded_register(const std::shared_ptr<Group>& G, std::shared_ptr<Sharded> s) {
  std::lock_guard<std::mutex> g(G->reg_mu);
  const uint64_t h = SHARDED_TAG | G->next++;
  G->segs[h] = std::move(s);
  return h;
}

int free_shards(const std::shared_ptr<Group>& G, const std::shared_ptr<Sharded>& s) {
  return fan_out(G, [s](int r) { return s->h[r] ? vs_segment_free(s->h[r]) : VS_OK; });
}

#define GRP(G, S)                                                                    \
  std::shared_ptr<Group> G = group_get();                                            \
  std::shared_ptr<Sharded> S = sharded_lookup(G, h);                                 \
  if (!S) return fail(VS_EHANDLE, "unknown (or freed) sharded segment handle")

}  // namespace

bool group_is_sharded(uint64_t h) { return (h & 0xffff000000000000ull) == SHARDED_TAG; }
int group_size() {
  std::shared_ptr<Group> G = group_get();
  return G ? G->n : 0;
}
bool group_wants_sharding() { return !t_is_worker && group_size() > 1; }

int group_start(int n, const int* cuda_devs) {
  auto G = std::make_shared<Group>();
  G->n = n;
  for (int i = 0; i < n; i++) {
    auto w = std::make_unique<Worker>();
    w->slot = i;
    w->dev = cuda_devs[i];
    G->w.push_back(std::move(w));
  }
  for (auto& w : G->w) w->th = std::thread(worker_main, w.get());
  std::vector<int> devs(cuda_devs, cuda_devs + n);
  std::vector<uint64_t> bases(n, 0);
  // peer access both ways between every pair, one communicator rank per GPU
  int rc = fan_out(G, [&](int r) {
    for (int p = 0; p < n; p++) {
      if (p == r || devs[p] == devs[r]) continue;
      int can = 0;
      cudaError_t e = cudaDeviceCanAccessPeer(&can, devs[r], devs[p]);
      if (e != cudaSuccess || !can) {
        cudaGetLastError();
        return fail(VS_ECUDA, "GPU %d cannot access GPU %d's memory (no NVLink / PCIe peer path)", devs[r], devs[p]);
      }
      e = cudaDeviceEnablePeerAccess(devs[p], 0);
      if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) {
        cudaGetLastError();
        return fail(VS_ECUDA, "cudaDeviceEnablePeerAccess(%d -> %d): %s", devs[r], devs[p], cudaGetErrorString(e));
      }
      cudaGetLastError();
    }
    uint8_t handle[VS_PEER_HANDLE_BYTES];
    int rr = vs_peer_create(r, n, GROUP_SLOT_BYTES, 8, &G->w[r]->comm, handle);
    if (rr != VS_OK) return rr;
    return vs_peer_base(G->w[r]->comm, &bases[r]);
  });
  if (rc == VS_OK) rc = fan_out(G, [&](int r) { return vs_peer_connect_ptrs(G->w[r]->comm, bases.data()); });
  {
    std::lock_guard<std::mutex> g(g_group_mu);
    g_group = G;
  }
  if (rc != VS_OK) {
    const std::string msg = vs_last_error();
    group_stop();
    set_error(msg.c_str());
  }
  return rc;
}

int group_stop() {
  std::shared_ptr<Group> G;
  {
    std::lock_guard<std::mutex> g(g_group_mu);
    G = std::move(g_group);
    g_group.reset();
  }
  if (!G) return VS_OK;
  std::vector<std::shared_ptr<Sharded>> segs;
  {
    std::lock_guard<std::mutex> g(G->reg_mu);
    for (auto& kv : G->segs) segs.push_back(kv.second);
    G->segs.clear();
  }
  for (auto& s : segs) free_shards(G, s);
  fan_out(G, [&](int r) {
    cudaDeviceSynchronize();
    return VS_OK;
  });
  fan_out(G, [&](int r) { return G->w[r]->comm ? vs_peer_destroy(G->w[r]->comm) : VS_OK; });
  for (auto& w : G->w) {
    {
      std::lock_guard<std::mutex> lk(w->mu);
      w->stop = true;
    }
    w->cv.notify_one();
  }
  for (auto& w : G->w)
    if (w->th.joinable()) w->th.join();
  return VS_OK;
}

// ---- residency ----------------------------------------------------------------------------------------------------
static int make_sharded(const std::shared_ptr<Group>& G, int64_t n, int d, int64_t id_base,
                        const std::function<int(int, int64_t, int64_t, uint64_t*)>& make, uint64_t* handle_out) {
  if (n < 0 || d <= 0) return fail(VS_EINVAL, "n must be >= 0 and d positive");
  auto s = std::make_shared<Sharded>();
  s->n = n;
  s->d = d;
  s->id_base = id_base;
  s->h.assign(G->n, 0);
  s->lo = boundaries(n, G->n);
  int rc = fan_out(G, [&](int r) { return make(r, s->lo[r], s->lo[r + 1] - s->lo[r], &s->h[r]); });
  if (rc != VS_OK) {
    const std::string msg = vs_last_error();
    free_shards(G, s);
    set_error(msg.c_str());
    return rc;
  }
  *handle_out = sharded_register(G, s);
  return VS_OK;
}

int group_segment_upload(const float* rows, int64_t n, int d, const uint8_t* skip, int64_t id_base, uint64_t* handle_out) {
  std::shared_ptr<Group> G = group_get();
  if (!G) return fail(VS_ESTATE, "no device group is bound");
  return make_sharded(G, n, d, id_base, [&](int, int64_t lo, int64_t cnt, uint64_t* out) {
    return vs_segment_upload(rows ? rows + (size_t)lo * d : nullptr, cnt, d, skip ? skip + lo : nullptr, id_base + lo, out);
  }, handle_out);
}

int group_segment_generate(int64_t seed, int64_t first_row, int64_t n, int d, int64_t id_base, uint64_t* handle_out) {
  std::shared_ptr<Group> G = group_get();
  if (!G) return fail(VS_ESTATE, "no device group is bound");
  return make_sharded(G, n, d, id_base, [&](int, int64_t lo, int64_t cnt, uint64_t* out) {
    return vs_segment_generate(seed, first_row + lo, cnt, d, id_base + lo, out);
  }, handle_out);
}

int group_segment_set_skip(uint64_t h, const uint8_t* skip) {
  GRP(G, s);
  return fan_out(G, [&](int r) { return vs_segment_set_skip(s->h[r], skip ? skip + s->lo[r] : nullptr); });
}

int group_segment_info(uint64_t h, int64_t* n, int32_t* d, int32_t* M, int32_t* K, int64_t* id_base) {
  GRP(G, s);
  if (n) *n = s->n;
  if (d) *d = s->d;
  if (M) *M = s->M;
  if (K) *K = s->K;
  if (id_base) *id_base = s->id_base;
  return VS_OK;
}

// rows / codes [first, first + count) live on the shards whose ranges overlap it
template <typename F>
static int for_overlap(const std::shared_ptr<Group>& G, const std::shared_ptr<Sharded>& s, int64_t first, int64_t count, F f) {
  if (first < 0 || count < 0 || first + count > s->n) return fail(VS_EINVAL, "row range out of bounds");
  return fan_out(G, [&](int r) {
    const int64_t a = std::max(first, s->lo[r]), b = std::min(first + count, s->lo[r + 1]);
    if (a >= b) return (int)VS_OK;
    return f(r, a - s->lo[r], b - a, a - first);
  });
}

int group_segment_download_rows(uint64_t h, int64_t first, int64_t count, float* out) {
  GRP(G, s);
  if (!out && count > 0) return fail(VS_EINVAL, "null output pointer");
  return for_overlap(G, s, first, count, [&](int r, int64_t lfirst, int64_t cnt, int64_t ooff) {
    return (int)vs_segment_download_rows(s->h[r], lfirst, cnt, out + (size_t)ooff * s->d);
  });
}

int group_segment_attach_pq(uint64_t h, const float* centroids, int M, int K, const uint8_t* codes) {
  GRP(G, s);
  int rc = fan_out(G, [&](int r) { return vs_segment_attach_pq(s->h[r], centroids, M, K, codes ? codes + (size_t)s->lo[r] * M : nullptr); });
  if (rc == VS_OK) {
    s->M = M;
    s->K = K;
  }
  return rc;
}

int group_segment_download_codes(uint64_t h, int64_t first, int64_t count, uint8_t* out) {
  GRP(G, s);
  if (s->M == 0) return fail(VS_ESTATE, "segment has no PQ attached");
  if (!out && count > 0) return fail(VS_EINVAL, "null output pointer");
  return for_overlap(G, s, first, count, [&](int r, int64_t lfirst, int64_t cnt, int64_t ooff) {
    return (int)vs_segment_download_codes(s->h[r], lfirst, cnt, out + (size_t)ooff * s->M);
  });
}

int group_segment_free(uint64_t h) {
  std::shared_ptr<Group> G = group_get();
  std::shared_ptr<Sharded> s;
  if (G) {
    std::lock_guard<std::mutex> g(G->reg_mu);
    auto it = G->segs.find(h);
    if (it != G->segs.end()) {
      s = it->second;
      G->segs.erase(it);
    }
  }
  if (!s) return fail(VS_EHANDLE, "unknown (or freed) sharded segment handle");
  return free_shards(G, s);
}

// ---- queries: the one-call exchange form on every worker; rank 0's merged lists are the caller's -------------------
template <typename CALL>
static int query_fan_out(const std::shared_ptr<Group>& G, int nq, int kout, int64_t* ids, double* scores, int32_t* counts,
                         CALL call) {
  if (!ids || !scores) return fail(VS_EINVAL, "null output pointer");
  return fan_out(G, [&](int r) {
    if (r == 0) return (int)call(r, ids, scores, counts);
    // the other ranks hold the same merged lists; only their participation is needed
    static thread_local std::vector<int64_t> ti;
    static thread_local std::vector<double> ts;
    static thread_local std::vector<int32_t> tc;
    ti.resize((size_t)nq * kout);
    ts.resize((size_t)nq * kout);
    tc.resize((size_t)nq);
    return (int)call(r, ti.data(), ts.data(), tc.data());
  });
}

// queries per call so that the packed lists of one exchange fit the communicator's slot
static int chunk_for(int nq, size_t bytes_per_query) {
  int64_t c = GROUP_SLOT_BYTES / (int64_t)(bytes_per_query > 0 ? bytes_per_query : 1);
  if (c < 1) c = 1;
  return (int)(c < nq ? c : nq);
}

int group_bruteforce_topk(uint64_t h, const float* q, int nq, int k, int metric, int64_t* ids, double* scores, int32_t* counts) {
  GRP(G, s);
  if (!q || nq <= 0 || k <= 0) return fail(VS_EINVAL, "q must be non-null, nq and k positive");
  const int chunk = chunk_for(nq, (size_t)2 * k * 8);
  for (int q0 = 0; q0 < nq; q0 += chunk) {
    const int c = std::min(chunk, nq - q0);
    int rc = query_fan_out(G, c, k, ids + (size_t)q0 * k, scores + (size_t)q0 * k, counts ? counts + q0 : nullptr,
                           [&](int r, int64_t* oi, double* os, int32_t* oc) {
                             return vs_bruteforce_topk_exchange(s->h[r], G->w[r]->comm, q + (size_t)q0 * s->d, c, k, metric, oi, os, oc);
                           });
    if (rc != VS_OK) return rc;
  }
  return VS_OK;
}

int group_adc_topk(uint64_t h, const float* q, int nq, int n_cand, int64_t* ids, double* approx, int32_t* counts) {
  GRP(G, s);
  if (!q || nq <= 0 || n_cand <= 0) return fail(VS_EINVAL, "q must be non-null, nq and n_cand positive");
  if (s->M == 0) return fail(VS_ESTATE, "segment has no PQ attached");
  const int chunk = chunk_for(nq, (size_t)2 * n_cand * 8);
  for (int q0 = 0; q0 < nq; q0 += chunk) {
    const int c = std::min(chunk, nq - q0);
    int rc = query_fan_out(G, c, n_cand, ids + (size_t)q0 * n_cand, approx + (size_t)q0 * n_cand, counts ? counts + q0 : nullptr,
                           [&](int r, int64_t* oi, double* os, int32_t* oc) {
                             return vs_adc_topk_exchange(s->h[r], G->w[r]->comm, q + (size_t)q0 * s->d, c, n_cand, oi, os, oc);
                           });
    if (rc != VS_OK) return rc;
  }
  return VS_OK;
}

int group_adc_rerank_topk(uint64_t h, const float* q, int nq, int n_cand, int k, int metric, int nor, int64_t* ids,
                          double* scores, int32_t* counts) {
  GRP(G, s);
  if (!q || nq <= 0 || n_cand <= 0 || k <= 0) return fail(VS_EINVAL, "q must be non-null, nq, n_cand and k positive");
  if (s->M == 0) return fail(VS_ESTATE, "segment has no PQ attached");
  const int chunk = chunk_for(nq, (size_t)4 * n_cand * 8);
  for (int q0 = 0; q0 < nq; q0 += chunk) {
    const int c = std::min(chunk, nq - q0);
    int rc = query_fan_out(G, c, k, ids + (size_t)q0 * k, scores + (size_t)q0 * k, counts ? counts + q0 : nullptr,
                           [&](int r, int64_t* oi, double* os, int32_t* oc) {
                             return vs_adc_rerank_topk_exchange(s->h[r], G->w[r]->comm, q + (size_t)q0 * s->d, c, n_cand, k, metric,
                                                                nor, oi, os, oc);
                           });
    if (rc != VS_OK) return rc;
  }
  return VS_OK;
}

int group_rerank_topk(uint64_t h, const float* q, const int64_t* cand, int n_cand, int k, int metric, int nor, int64_t* ids,
                      double* scores, int32_t* count) {
  GRP(G, s);
  if (!q || (!cand && n_cand > 0) || n_cand < 0 || k <= 0) return fail(VS_EINVAL, "null pointer, negative n_cand or k <= 0");
  if (n_cand == 0) {
    for (int i = 0; i < k; i++) {
      ids[i] = -1;
      scores[i] = __builtin_nan("");
    }
    if (count) *count = 0;
    return VS_OK;
  }
  // the exchange form ranks k <= n_cand results; a longer request is padded below
  const int kk = std::min(k, n_cand);
  std::vector<int64_t> oi0(kk);
  std::vector<double> os0(kk);
  int32_t oc0 = 0;
  int rc = query_fan_out(G, 1, kk, oi0.data(), os0.data(), &oc0, [&](int r, int64_t* oi, double* os, int32_t* oc) {
    return vs_rerank_topk_exchange(s->h[r], G->w[r]->comm, q, cand, n_cand, kk, metric, nor, oi, os, oc);
  });
  if (rc != VS_OK) return rc;
  for (int i = 0; i < k; i++) {
    ids[i] = i < kk ? oi0[i] : -1;
    scores[i] = i < kk ? os0[i] : __builtin_nan("");
  }
  if (count) *count = oc0;
  return VS_OK;
}

// ---- builds -----------------------------------------------------------------------------------------------------------
static std::atomic<int> g_train_exact{1};
void group_set_train_exact(int v) { g_train_exact.store(v ? 1 : 0); }

int group_pq_train(uint64_t h, int64_t n, int d, int M, int K, int iterations, int64_t seed, float* centroids_out) {
  GRP(G, s);
  if (s->d != d) return fail(VS_EINVAL, "segment dimension %d != d %d", s->d, d);
  if (n != s->n) return fail(VS_EINVAL, "a sharded segment trains on all of its rows (n must be %lld)", (long long)s->n);
  for (int r = 0; r < G->n; r++)
    if (s->lo[r + 1] == s->lo[r]) return fail(VS_EINVAL, "fewer rows than GPUs: train this segment on one device");
  const size_t cbytes = (size_t)K * d * 4;
  const int exact = g_train_exact.load();
  return fan_out(G, [&](int r) {
    if (r == 0) return (int)vs_pq_train_sharded_peer(s->h[r], G->w[r]->comm, s->n, s->lo[r], exact, M, K, iterations, seed, centroids_out);
    static thread_local std::vector<float> tmp;  // every rank ends with the same centroids
    tmp.resize(cbytes / 4);
    return (int)vs_pq_train_sharded_peer(s->h[r], G->w[r]->comm, s->n, s->lo[r], exact, M, K, iterations, seed, tmp.data());
  });
}

int group_pq_encode(uint64_t h, const float* centroids, int M, int K, int subDim, int64_t n, uint8_t* codes_out) {
  GRP(G, s);
  if (s->d != M * subDim) return fail(VS_EINVAL, "segment dimension %d != M*subDim %d", s->d, M * subDim);
  if (n > s->n) return fail(VS_EINVAL, "n exceeds the segment's row count");
  return fan_out(G, [&](int r) {
    const int64_t cnt = std::min(n, s->lo[r + 1]) - s->lo[r];
    if (cnt <= 0) return (int)VS_OK;
    return (int)vs_pq_encode_batch(centroids, M, K, subDim, nullptr, s->h[r], cnt, codes_out + (size_t)s->lo[r] * M);
  });
}

// ---- BEST_FIRST expansion scoring over a sharded segment: one LUT per shard, every shard answers for its ids ----------
int group_adc_query_begin(uint64_t h, const float* q, std::vector<uint64_t>* shard_q) {
  GRP(G, s);
  if (s->M == 0) return fail(VS_ESTATE, "segment has no PQ attached");
  shard_q->assign(G->n, 0);
  int rc = fan_out(G, [&](int r) { return s->lo[r + 1] > s->lo[r] ? vs_adc_query_begin(s->h[r], q, &(*shard_q)[r]) : VS_OK; });
  if (rc != VS_OK) {
    const std::string msg = vs_last_error();
    group_adc_query_end(*shard_q);
    set_error(msg.c_str());
  }
  return rc;
}

int group_adc_query_gather(uint64_t h, const std::vector<uint64_t>& shard_q, const int64_t* ids, int64_t n_ids, double* out,
                           uint8_t* valid) {
  GRP(G, s);
  const double nan = __builtin_nan("");
  std::vector<std::vector<int64_t>> pos(G->n);
  for (int64_t i = 0; i < n_ids; i++) {
    out[i] = nan;  // ids outside every shard have no code
    if (valid) valid[i] = 0;
    const int64_t row = ids[i] - s->id_base;
    if (row < 0 || row >= s->n) continue;
    const int r = (int)(std::upper_bound(s->lo.begin(), s->lo.end(), row) - s->lo.begin()) - 1;
    pos[r].push_back(i);
  }
  return fan_out(G, [&](int r) {
    const size_t m = pos[r].size();
    if (m == 0) return (int)VS_OK;
    std::vector<int64_t> lid(m);
    std::vector<double> lo(m);
    std::vector<uint8_t> lv(m);
    for (size_t j = 0; j < m; j++) lid[j] = ids[pos[r][j]];
    int rc = vs_adc_query_gather(shard_q[r], lid.data(), (int64_t)m, lo.data(), lv.data());
    if (rc != VS_OK) return rc;
    for (size_t j = 0; j < m; j++) {
      out[pos[r][j]] = lo[j];
      if (valid) valid[pos[r][j]] = lv[j];
    }
    return (int)VS_OK;
  });
}

int group_adc_query_end(const std::vector<uint64_t>& shard_q) {
  std::shared_ptr<Group> G = group_get();
  if (!G) return VS_OK;
  return fan_out(G, [&](int r) { return r < (int)shard_q.size() && shard_q[r] ? vs_adc_query_end(shard_q[r]) : VS_OK; });
}

int group_segment_upload_strided(const uint8_t* bytes, int64_t n, int d, int64_t stride, const uint8_t* skip, int64_t id_base,
                                 uint64_t* handle_out) {
  std::shared_ptr<Group> G = group_get();
  if (!G) return fail(VS_ESTATE, "no device group is bound");
  return make_sharded(G, n, d, id_base, [&](int, int64_t lo, int64_t cnt, uint64_t* out) {
    return vs_segment_upload_strided(bytes ? bytes + (size_t)lo * stride : nullptr, cnt, d, stride, skip ? skip + lo : nullptr,
                                     id_base + lo, out);
  }, handle_out);
}

int group_segment_upload_records(const uint8_t* buf, const int64_t* offsets, int64_t n, int d, int64_t id_base,
                                 int32_t* vec_ids_out, uint64_t* handle_out) {
  std::shared_ptr<Group> G = group_get();
  if (!G) return fail(VS_ESTATE, "no device group is bound");
  return make_sharded(G, n, d, id_base, [&](int, int64_t lo, int64_t cnt, uint64_t* out) {
    return vs_segment_upload_records(buf, offsets + lo, cnt, d, id_base + lo, vec_ids_out ? vec_ids_out + lo : nullptr, out);
  }, handle_out);
}

}  // namespace vs
