// api.cu -- the C ABI of libvsgpu (include/vsgpu.h): handle table, per-thread streams and
// scratch, host<->device staging, launch configuration.  No torch types, no CPU fallback: every
// compute entry point fails with VS_ECUDA when no device is bound.
#include <algorithm>
#include <atomic>
#include <chrono>
#include <climits>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <string>
#include <thread>
#include <unordered_map>
#include <vector>

#include <cuda.h>

#include "../../include/vsgpu.h"
#include "host.h"
#include "kernels.h"
#include "topk.cuh"

namespace vs {

static std::atomic<int64_t> g_launches{0};
void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

// Devices this process drives: slot i -> CUDA ordinal g_devs[i]; slot 0 is the primary device (vs_init(device) binds
// exactly one).  A host thread works on its default device (a group worker's own GPU, else the primary) unless the
// call names a segment, which carries the device its rows live on.
static std::atomic<int> g_device{-1};
static std::atomic<int> g_ndev{0};
static int g_devs[VS_MAX_DEVICES] = {};
static thread_local int t_default_dev = -1;
static std::atomic<int> g_lanes{16};
static std::atomic<int64_t> g_adc_fast_min_rows{16384};   // below this the generic ADC kernel runs
static std::atomic<int64_t> g_adc_fast_cap{4096};         // candidate-list entries per scan CTA
static std::atomic<int64_t> g_batch_min_queries{3};        // query batches at least this large use batch.cu
static std::atomic<int64_t> g_batch_min_rows{16384};       // ... on segments at least this long
static std::atomic<int64_t> g_batch_gm_bytes{int64_t(1) << 30};  // group-minima scratch per query chunk
static std::atomic<int64_t> g_batch_warp_min_q{0};         // > 0: batches this large always select with one warp per query (tests)
static std::atomic<int> g_peer_fused{1};                      // one-query exchanges: publish inside the merge kernel
static std::atomic<int> g_peer_spin_shared{0};                // tests: poll inside the kernels even when ranks share a device
static std::atomic<int64_t> g_scan_reserve_sms{0};          // SMs the one-query scan leaves free (for a collective's CTAs)
// cta_group::2 nomination kernel for batches > 128 queries: 0 never, 1 wherever it fits, 2 (default) long vectors only --
// where both operands stream it moves a third less through L2 and the tensor pipe runs at 97 %; with the query
// block resident (d = 128) it measured equal to the single-CTA kernel
static std::atomic<int64_t> g_batch_pairs{2};
static std::atomic<int64_t> g_batch_prefilter{1};           // candidate groups are pre-filtered on the fp16 copy before exact scoring
static std::atomic<int64_t> g_pdl{1};                       // programmatic dependent launch between the kernels of one call
static std::atomic<int64_t> g_scan_fp16{1};                 // 1: single queries nominate on the fp16 copy (half the bytes), 0: fp32 scan
static std::atomic<int64_t> g_adc_reserve_sms{0};           // SMs the fast ADC scan leaves free (queries alternating between streams)
static std::atomic<int64_t> g_scan_half_ctas{0};            // scan_half_kernel CTAs per SM: 0 = automatic (two where they fit), 1, 2 (diagnostics)
static std::atomic<int64_t> g_batch_select_ctas{0};         // 0 = automatic selection CTAs per query, else 1..64 (diagnostics)
static std::atomic<int64_t> g_batch_group{0};               // 0 = automatic rows per nomination group, else 16 / 32 / 64
static std::atomic<int64_t> g_batch_fp16{1};                // nominate on fp16 operand copies (0: the fp32 rows as tf32)
static int g_sms = 0;
static std::mutex g_mu;
static std::unordered_map<uint64_t, std::shared_ptr<Segment>> g_segments;
static uint64_t g_next_handle = 1;

static thread_local std::string t_err;

int fail(int code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  t_err = buf;
  return code;
}

int cuda_fail(cudaError_t e, const char* what) {
  // leave the sticky error (if any) visible in the message; clear the non-sticky ones
  cudaGetLastError();
  return fail(e == cudaErrorMemoryAllocation ? VS_ENOMEM : VS_ECUDA, "%s: %s", what, cudaGetErrorString(e));
}

int lanes() { return g_lanes.load(); }
void set_thread_device(int cuda_dev) { t_default_dev = cuda_dev; }
int primary_device() { return g_device.load(); }
static bool device_registered(int dev) {
  const int n = g_ndev.load();
  for (int i = 0; i < n; i++)
    if (g_devs[i] == dev) return true;
  return false;
}
void set_error(const char* msg) { t_err = msg ? msg : ""; }
bool pdl_enabled() { return g_pdl.load() != 0; }
int sm_count() { return g_sms; }

// ---- per-thread context ------------------------------------------------------------------------------
struct ThreadCtx {
  int device = -1;
  cudaStream_t stream = nullptr;
  void* d_buf = nullptr;
  size_t d_cap = 0;
  void* h_buf = nullptr;  // pinned
  size_t h_cap = 0;
  unsigned long long* d_ticket = nullptr;  // [queries][4] control words, zero between launches
  size_t ticket_cap = 0;
  unsigned int* d_fs = nullptr;  // [queries][FS_WORDS] ADC fast-scan histogram + control, zero between launches
  size_t fs_cap = 0;
  // The device scratch above is the ACTIVE set.  Work enqueued on different caller streams (the *_dev entry
  // points) may overlap on the GPU, so every stream gets its own set: ctx_use_stream parks the active set in
  // its slot and activates the one of the requested stream.
  struct DevScratch {
    void* key = nullptr;  // the stream this set belongs to
    bool used = false;
    void* d_buf = nullptr;
    size_t d_cap = 0;
    unsigned long long* d_ticket = nullptr;
    size_t ticket_cap = 0;
    unsigned int* d_fs = nullptr;
    size_t fs_cap = 0;
  };
  static constexpr int MAX_SLOTS = 4;
  DevScratch slots[MAX_SLOTS];
  int active = -1;
  std::vector<void*> parked_host;  // outgrown pinned buffers, released with the context
  void park() {
    if (active < 0) return;
    DevScratch& a = slots[active];
    a.d_buf = d_buf; a.d_cap = d_cap; a.d_ticket = d_ticket; a.ticket_cap = ticket_cap; a.d_fs = d_fs; a.fs_cap = fs_cap;
  }
  void activate(int i) {
    const DevScratch& a = slots[i];
    d_buf = a.d_buf; d_cap = a.d_cap; d_ticket = a.d_ticket; ticket_cap = a.ticket_cap; d_fs = a.d_fs; fs_cap = a.fs_cap;
    active = i;
  }
  ~ThreadCtx() {
    // process teardown: the context may already be gone; ignore errors
    if (device >= 0 && device_registered(device)) {
      cudaSetDevice(device);
      park();
      for (DevScratch& a : slots) {
        if (a.d_buf) cudaFree(a.d_buf);
        if (a.d_ticket) cudaFree(a.d_ticket);
        if (a.d_fs) cudaFree(a.d_fs);
      }
      if (h_buf) cudaFreeHost(h_buf);
      for (void* p : parked_host) cudaFreeHost(p);
      if (stream) cudaStreamDestroy(stream);
    }
  }
};
static thread_local ThreadCtx t_ctxs[VS_MAX_DEVICES];  // one per CUDA ordinal this thread has worked on
static void park_host(ThreadCtx* c, void* p) { c->parked_host.push_back(p); }

// Makes the device scratch of `stream` the active one (at most MAX_SLOTS streams per host thread; beyond that
// the least recently bound caller stream is drained and its set is handed over).
static int ctx_use_stream(ThreadCtx* c, void* stream) {
  if (c->active >= 0 && c->slots[c->active].key == stream) return VS_OK;
  c->park();
  int free_slot = -1;
  for (int i = 0; i < ThreadCtx::MAX_SLOTS; i++) {
    if (c->slots[i].used && c->slots[i].key == stream) {
      c->activate(i);
      return VS_OK;
    }
    if (!c->slots[i].used && free_slot < 0) free_slot = i;
  }
  if (free_slot < 0) {  // recycle a caller-stream slot (never slot 0, the thread's own stream)
    free_slot = 1 + (c->active >= 1 ? c->active % (ThreadCtx::MAX_SLOTS - 1) : 0);
    cudaError_t e = cudaStreamSynchronize(static_cast<cudaStream_t>(c->slots[free_slot].key));
    if (e == cudaErrorInvalidResourceHandle || e == cudaErrorContextIsDestroyed || e == cudaErrorDeviceUninitialized) {
      // the caller has destroyed that stream since: its work is done or abandoned; drain the device instead
      cudaGetLastError();
      e = cudaDeviceSynchronize();
    }
    if (e != cudaSuccess) return cuda_fail(e, "sync (scratch hand-over)");
  }
  c->slots[free_slot].used = true;
  c->slots[free_slot].key = stream;
  c->activate(free_slot);
  return VS_OK;
}

// A stream the library owns is about to be destroyed: the calling thread's scratch sets must not keep it as a key
// (a recycled slot would otherwise synchronise a dead handle).  The device has been drained by the caller.
static void ctx_forget_stream(int dev, void* stream) {
  if (dev < 0 || dev >= VS_MAX_DEVICES) return;
  ThreadCtx& c = t_ctxs[dev];
  if (c.device != dev) return;
  for (int i = 1; i < ThreadCtx::MAX_SLOTS; i++) {
    if (c.slots[i].used && c.slots[i].key == stream) {
      if (c.active == i) {
        c.park();
        c.activate(0);
      }
      c.slots[i].key = nullptr;  // keeps its buffers; the next new stream takes the slot over
      c.slots[i].used = false;
    }
  }
}

static int ctx_bind_dev(ThreadCtx** out, int dev) {
  if (dev < 0) return fail(VS_ECUDA, "vs_init has not bound a CUDA device (there is no CPU fallback)");
  if (dev >= VS_MAX_DEVICES) return fail(VS_ECUDA, "CUDA device ordinal %d is beyond the %d this build supports", dev, VS_MAX_DEVICES);
  ThreadCtx& c = t_ctxs[dev];
  cudaError_t e = cudaSetDevice(dev);
  if (e != cudaSuccess) return cuda_fail(e, "cudaSetDevice");
  if (c.device != dev) {
    c = ThreadCtx{};
    c.device = dev;
    e = cudaStreamCreateWithFlags(&c.stream, cudaStreamNonBlocking);
    if (e != cudaSuccess) return cuda_fail(e, "cudaStreamCreate");
  }
  *out = &c;
  return ctx_use_stream(&c, c.stream);
}
// the calling thread's default device: a group worker's own GPU, else the primary
int ctx_bind(ThreadCtx** out) { return ctx_bind_dev(out, t_default_dev >= 0 ? t_default_dev : g_device.load()); }
// the device a segment's rows live on
static int ctx_bind_seg(ThreadCtx** out, const Segment* s) { return ctx_bind_dev(out, s->device); }

// Scratch grows inside calls that may be collective: a rank of a peer exchange must never synchronise the whole DEVICE
// on its way to publishing, because a peer that shares the GPU (several ranks on one device: the one-GPU test shape of
// vs_init_multi / vs_peer_connect_ptrs) may be spinning in its merge kernel for exactly that publish.  cudaFree does
// synchronise the device; stream-ordered allocation and release (and waiting for ONE stream) do not.
static cudaError_t scratch_alloc(void** p, size_t bytes, cudaStream_t st) {
  cudaError_t e = cudaMallocAsync(p, bytes, st);
  if (e == cudaSuccess) e = cudaStreamSynchronize(st);  // usable from any stream from here on
  return e;
}
static void scratch_free(void* p, cudaStream_t st) {
  if (p) cudaFreeAsync(p, st);
}
// pinned host buffers that were outgrown are parked (cudaFreeHost may synchronise too) and released with the context
static int ctx_reserve_dev(ThreadCtx* c, size_t bytes) {
  if (bytes <= c->d_cap) return VS_OK;
  // work enqueued on the stream that owns this scratch set may still use the old buffer: order the release behind it
  cudaStream_t owner = (c->active >= 0 && c->slots[c->active].key) ? static_cast<cudaStream_t>(c->slots[c->active].key) : c->stream;
  if (owner != c->stream) {
    cudaEvent_t ev;
    if (cudaEventCreateWithFlags(&ev, cudaEventDisableTiming) == cudaSuccess) {
      cudaEventRecord(ev, owner);
      cudaStreamWaitEvent(c->stream, ev, 0);
      cudaEventDestroy(ev);
    }
  }
  scratch_free(c->d_buf, c->stream);
  c->d_buf = nullptr;
  c->d_cap = 0;
  size_t cap = bytes + bytes / 4 + 4096;
  cudaError_t e = scratch_alloc(&c->d_buf, cap, c->stream);
  if (e != cudaSuccess) return cuda_fail(e, "cudaMallocAsync(scratch)");
  c->d_cap = cap;
  return VS_OK;
}
static int ctx_reserve_host(ThreadCtx* c, size_t bytes) {
  if (bytes <= c->h_cap) return VS_OK;
  cudaStreamSynchronize(c->stream);
  if (c->h_buf) park_host(c, c->h_buf);
  c->h_buf = nullptr;
  c->h_cap = 0;
  size_t cap = bytes + bytes / 4 + 4096;
  cudaError_t e = cudaMallocHost(&c->h_buf, cap);
  if (e != cudaSuccess) return cuda_fail(e, "cudaMallocHost(staging)");
  c->h_cap = cap;
  return VS_OK;
}
static int ctx_reserve_ticket(ThreadCtx* c, size_t n) {
  if (n <= c->ticket_cap) return VS_OK;
  cudaStreamSynchronize(c->stream);
  scratch_free(c->d_ticket, c->stream);
  c->d_ticket = nullptr;
  c->ticket_cap = 0;
  size_t cap = n < 1024 ? 1024 : n * 2;
  void* tp = nullptr;
  cudaError_t e = scratch_alloc(&tp, cap * 4 * sizeof(unsigned long long), c->stream);
  c->d_ticket = static_cast<unsigned long long*>(tp);
  if (e != cudaSuccess) return cuda_fail(e, "cudaMallocAsync(ticket)");
  e = cudaMemsetAsync(c->d_ticket, 0, cap * 4 * sizeof(unsigned long long), c->stream);
  if (e != cudaSuccess) return cuda_fail(e, "cudaMemset(ticket)");
  c->ticket_cap = cap;
  return VS_OK;
}

static int ctx_reserve_fs(ThreadCtx* c, size_t n) {
  if (n <= c->fs_cap) return VS_OK;
  cudaStreamSynchronize(c->stream);
  scratch_free(c->d_fs, c->stream);
  c->d_fs = nullptr;
  c->fs_cap = 0;
  size_t cap = n < 4 ? 4 : n * 2;
  void* fp = nullptr;
  cudaError_t e = scratch_alloc(&fp, cap * FS_WORDS * sizeof(unsigned int), c->stream);
  c->d_fs = static_cast<unsigned int*>(fp);
  if (e != cudaSuccess) return cuda_fail(e, "cudaMallocAsync(fast-scan scratch)");
  e = cudaMemsetAsync(c->d_fs, 0, cap * FS_WORDS * sizeof(unsigned int), c->stream);
  if (e != cudaSuccess) return cuda_fail(e, "cudaMemset(fast-scan scratch)");
  e = cudaStreamSynchronize(c->stream);
  if (e != cudaSuccess) return cuda_fail(e, "sync");
  c->fs_cap = cap;
  return VS_OK;
}

// bump allocator over the per-thread scratch
struct Arena {
  char* base;
  size_t off = 0;
  explicit Arena(void* b) : base(static_cast<char*>(b)) {}
  template <typename T>
  T* take(size_t count) {
    off = (off + 255) & ~size_t(255);
    T* p = reinterpret_cast<T*>(base + off);
    off += count * sizeof(T);
    return p;
  }
  static size_t need(std::initializer_list<size_t> sizes) {
    size_t o = 0;
    for (size_t s : sizes) o = ((o + 255) & ~size_t(255)) + s;
    return o + 256;
  }
};

// ---- segments -----------------------------------------------------------------------------------------
std::shared_ptr<Segment> seg_lookup(uint64_t h) {
  std::lock_guard<std::mutex> lk(g_mu);
  auto it = g_segments.find(h);
  return it == g_segments.end() ? nullptr : it->second;
}
static void seg_destroy(Segment* s) {
  if (!s) return;
  // A query on another thread may have enqueued work on these buffers without having synchronised yet (the _dev
  // entry points): the last reference drains the device before the memory goes back.
  if (device_registered(s->device)) {
    cudaSetDevice(s->device);
    cudaDeviceSynchronize();
    if (s->X) cudaFree(s->X);
    if (s->skip) cudaFree(s->skip);
    if (s->centroids) cudaFree(s->centroids);
    if (s->codes) cudaFree(s->codes);
    for (int m = 0; m < 2; m++) {
      if (s->ab[m]) cudaFree(s->ab[m]);
      if (s->stats[m]) cudaFree(s->stats[m]);
    }
    if (s->Xh) cudaFree(s->Xh);
  }
  delete s;
}
// Handles own one reference; every entry point holds another for the duration of the call, so vs_segment_free on one
// thread cannot pull the rows from under a query running on another (the fan-out of J/fdb/FdbVectorIndex.java:418-432).
static uint64_t seg_register(Segment* s) {
  std::shared_ptr<Segment> sp(s, seg_destroy);
  std::lock_guard<std::mutex> lk(g_mu);
  uint64_t h = g_next_handle++;
  g_segments[h] = std::move(sp);
  return h;
}

#define CK(call, what)                                   \
  do {                                                   \
    cudaError_t _e = (call);                             \
    if (_e != cudaSuccess) return cuda_fail(_e, what);   \
  } while (0)
#define RET(call)                 \
  do {                            \
    int _r = (call);              \
    if (_r != VS_OK) return _r;   \
  } while (0)

// ---- launch configuration ------------------------------------------------------------------------------
struct ScanPlan {
  int grid, threads, kp;
  size_t smem;
};

// occupancy queries (and the smem attribute they set) are cached per kernel configuration
static std::mutex g_occ_mu;
static std::unordered_map<uint64_t, int> g_occ_cache;
template <typename F>
static int cached_occ(uint64_t key, F compute) {
  {
    std::lock_guard<std::mutex> lk(g_occ_mu);
    auto it = g_occ_cache.find(key);
    if (it != g_occ_cache.end()) return it->second;
  }
  const int occ = compute();
  std::lock_guard<std::mutex> lk(g_occ_mu);
  g_occ_cache[key] = occ;
  return occ;
}
// (kernel attributes such as the dynamic shared-memory opt-in are per device: the current device is part of every key)
static uint64_t occ_key(int kind, int a, int b, int c, int threads, size_t smem) {
  int dev = 0;
  cudaGetDevice(&dev);
  uint64_t h = 1469598103934665603ull ^ ((uint64_t)(dev + 1) * 0x9e3779b97f4a7c15ull);
  for (uint64_t v : {(uint64_t)kind, (uint64_t)a, (uint64_t)b, (uint64_t)c, (uint64_t)threads, (uint64_t)smem}) {
    h ^= v;
    h *= 1099511628211ull;
  }
  return h;
}

// scan launch shapes are cached per (n, d, nq > 1, k, metric, lanes)
static std::mutex g_scan_mu;
static std::unordered_map<uint64_t, ScanLaunch> g_scan_cache;
struct AdcPlan {
  bool fast;
  AdcScanLaunch slow;
  AdcFastLaunch fastL;
  int64_t partial_keys;     // keys of per-CTA list scratch per query
  size_t cand_entries;      // candidate-list entries per query (fast path)
  size_t extra_words;       // 8-byte units per query that follow the lists (LUT extremes)
};
static std::unordered_map<uint64_t, AdcPlan> g_adc_cache;

static int plan_scan(const Segment* s, int nq, int k, bool cosine, ScanLaunch* out) {
  const int ln = lanes();
  int sms = g_sms - (int)g_scan_reserve_sms.load();
  if (sms < 1) sms = 1;
  uint64_t key = occ_key(cosine ? 11 : 10, s->d, ln, k, (nq > 1 ? 2 : 1) + 4 * sms, (size_t)s->n);
  {
    std::lock_guard<std::mutex> lk(g_scan_mu);
    auto it = g_scan_cache.find(key);
    if (it != g_scan_cache.end()) {
      *out = it->second;
      return VS_OK;
    }
  }
  ScanLaunch L{};
  L.n = s->n; L.d = s->d; L.nq = nq; L.lanes = ln; L.cosine = cosine; L.k = k;
  if (!scan_configure(L, sms)) return fail(VS_ECUDA, "scan kernel cannot be resident for d=%d k=%d", s->d, k);
  std::lock_guard<std::mutex> lk(g_scan_mu);
  g_scan_cache[key] = L;
  *out = L;
  return VS_OK;
}

static int plan_adc(const Segment* s, int nq, int k, AdcPlan* out) {
  const int64_t min_rows = g_adc_fast_min_rows.load(), cap_opt = g_adc_fast_cap.load();
  uint64_t key = occ_key(20, s->M, s->K, k, nq > 1 ? 2 : 1, (size_t)s->n) ^ ((uint64_t)min_rows * 0x9e3779b97f4a7c15ull) ^
                 ((uint64_t)cap_opt << 17);
  {
    std::lock_guard<std::mutex> lk(g_scan_mu);
    auto it = g_adc_cache.find(key);
    if (it != g_adc_cache.end()) {
      *out = it->second;
      return VS_OK;
    }
  }
  AdcPlan P{};
  P.fast = adc_fast_supported(s->M, s->K) && s->n >= min_rows && s->n < (int64_t(1) << 47);
  if (P.fast) {
    AdcFastLaunch& L = P.fastL;
    L.n = s->n; L.M = s->M; L.K = s->K; L.nq = nq; L.k = k;
    if (!adc_fast_configure(L, g_sms)) return fail(VS_ECUDA, "ADC fast-scan kernel cannot be resident for M=%d k=%d", s->M, k);
    L.cap = (unsigned int)(cap_opt < 1 ? 1 : cap_opt);
    P.partial_keys = L.partial_keys;
    P.cand_entries = (size_t)L.cap * L.grid;
    P.extra_words = (size_t)s->M * 2;  // per-subspace {min, max} of the LUT
  } else {
    AdcScanLaunch& L = P.slow;
    L.n = s->n; L.M = s->M; L.K = s->K; L.nq = nq; L.k = k;
    if (!adc_configure(L, g_sms)) return fail(VS_ECUDA, "ADC scan kernel cannot be resident for M=%d k=%d", s->M, k);
    P.partial_keys = L.partial_keys;
    P.cand_entries = 0;
  }
  std::lock_guard<std::mutex> lk(g_scan_mu);
  g_adc_cache[key] = P;
  *out = P;
  return VS_OK;
}

static void fill_empty(int64_t* ids, double* scores, int32_t* counts, int nq, int k) {
  const double nan = __builtin_nan("");
  for (int64_t i = 0; i < (int64_t)nq * k; i++) {
    ids[i] = -1;
    scores[i] = nan;
  }
  if (counts)
    for (int i = 0; i < nq; i++) counts[i] = 0;
}

// ---- device-side bodies (everything on c->stream, no synchronisation) -------------------------------------
static int bruteforce_dev(ThreadCtx* c, cudaStream_t st, const Segment* s, const float* d_q, int nq, int k,
                          int metric, int64_t* d_ids, double* d_scores, int32_t* d_counts,
                          ulonglong2* d_partial, unsigned long long* d_ticket, const ScanLaunch& p,
                          int64_t out_stride = 0) {
  (void)c;
  (void)k;
  (void)metric;
  ScanLaunch L = p;
  L.X = s->X; L.q = d_q; L.nq = nq; L.skip = s->skip;
  L.partial = d_partial; L.ctrl = d_ticket;
  L.ids_out = d_ids; L.scores_out = d_scores; L.counts_out = d_counts; L.id_base = s->id_base;
  L.out_stride = out_stride;
  CK(launch_scan(L, st), "scan launch");
  return VS_OK;
}

static int adc_dev(cudaStream_t st, const Segment* s, const float* d_q, int nq, int n_cand, double* d_lut,
                   int64_t* d_ids, double* d_approx, int32_t* d_counts, ulonglong2* d_partial,
                   unsigned long long* d_ticket, unsigned int* d_fs, unsigned long long* d_cand, const AdcPlan& p,
                   int64_t out_stride = 0) {
  (void)n_cand;
  if (p.fast) {
    // the candidate scratch also carries the per-subspace LUT extremes: [nq * cand_entries][nq * M * 2]
    unsigned long long* d_mm = d_cand + (size_t)nq * p.cand_entries;
    CK(launch_build_lut_mm(s->centroids, s->M, s->K, s->subDim, d_q, nq, lanes(), d_lut, d_mm, st), "build_lut launch");
    AdcFastLaunch L = p.fastL;
    L.codes = s->codes; L.lut64 = d_lut; L.mm = d_mm; L.nq = nq; L.fs = d_fs; L.cand = d_cand;
    L.partial = d_partial; L.ctrl = d_ticket; L.ids_out = d_ids; L.approx_out = d_approx;
    L.counts_out = d_counts; L.id_base = s->id_base; L.out_stride = out_stride;
    L.reserve_sms = (int)g_adc_reserve_sms.load();
    CK(launch_adc_fast(L, st), "adc fast-scan launch");
    CK(launch_adc_fallback(L, st), "adc fallback launch");
    return VS_OK;
  }
  CK(launch_build_lut(s->centroids, s->M, s->K, s->subDim, d_q, nq, lanes(), d_lut, st), "build_lut launch");
  AdcScanLaunch L = p.slow;
  L.codes = s->codes; L.lut64 = d_lut; L.nq = nq;
  L.partial = d_partial; L.ctrl = d_ticket; L.ids_out = d_ids; L.approx_out = d_approx;
  L.counts_out = d_counts; L.id_base = s->id_base; L.out_stride = out_stride;
  CK(launch_adc_scan(L, st), "adc scan launch");
  return VS_OK;
}

static int rerank_dev(cudaStream_t st, const Segment* s, const float* d_q, int nq, const int64_t* d_cand,
                      int nc, int k, int metric, int64_t* d_ids, double* d_scores, int32_t* d_counts) {
  RankLaunch L{};
  L.X = s->X; L.n = s->n; L.d = s->d; L.skip = s->skip; L.lanes = lanes(); L.q = d_q; L.nq = nq;
  L.cand_ids = d_cand; L.nc = nc; L.k = k; L.metric = metric; L.id_base = s->id_base;
  L.ids_out = d_ids; L.scores_out = d_scores; L.counts_out = d_counts;
  CK(launch_rank(L, st), "rank launch");
  return VS_OK;
}

// ---- batched-query brute force (batch.cu) ---------------------------------------------------------------
static std::unordered_map<uint64_t, BatchLaunch> g_batch_cache;

// 0: per-query fp32 scan (scan.cu); 1: tensor-core nomination (query batches); 2: the CUDA-core scan of the fp16 copy
// for the one or two queries below the batch threshold (option "scan_fp16")
static int batch_wanted(const Segment* s, int nq, bool cosine) {
  if (s->n < g_batch_min_rows.load() || !batch_supported(s->d, lanes(), cosine, s->n)) return 0;
  if (nq >= g_batch_min_queries.load()) return 1;
  return g_scan_fp16.load() != 0 && g_batch_fp16.load() != 0 ? 2 : 0;
}

// HBM a resident segment of n x d rows may grow by at its first queries (wire.cu's residency budget counts it up
// front): the nomination coefficients of both metrics and, unless switched off, the fp16 operand copy.
int64_t nomination_aux_bytes(int64_t n, int d) {
  if (n < g_batch_min_rows.load() || !batch_supported(d, lanes(), false, n)) return 0;
  const int64_t dp = (d + 7) & ~7;
  return 2 * n * 4 + (g_batch_fp16.load() != 0 ? n * dp * 2 : 0);
}

static int plan_batch(const Segment* s, int k, bool cosine, bool half, BatchLaunch* out) {
  const int ln = lanes();
  const int go = (int)g_batch_group.load(), sp = (int)g_batch_select_ctas.load(), so = (int)g_scan_half_ctas.load();
  const uint64_t key = occ_key(cosine ? 31 : 30, s->d, ln, k, (half ? 1 : 0) + 2 * go + 256 * sp + 256 * 128 * so, (size_t)s->n);
  {
    std::lock_guard<std::mutex> lk(g_scan_mu);
    auto it = g_batch_cache.find(key);
    if (it != g_batch_cache.end()) {
      *out = it->second;
      return VS_OK;
    }
  }
  BatchLaunch L{};
  L.n = s->n; L.d = s->d; L.lanes = ln; L.cosine = cosine; L.k = k; L.half = half; L.dp = (s->d + 7) & ~7; L.group_override = go;
  L.select_ctas_override = sp;
  L.sh_override = so;
  if (!batch_configure(L, g_sms)) return fail(VS_ECUDA, "batched scan cannot be resident for d=%d k=%d", s->d, k);
  std::lock_guard<std::mutex> lk(g_scan_mu);
  g_batch_cache[key] = L;
  *out = L;
  return VS_OK;
}

static void batch_invalidate(Segment* s) {  // caller holds s->mu and has synchronised the device work
  // (the fp16 copy does not depend on the skip mask and stays)
  for (int m = 0; m < 2; m++) {
    if (s->ab[m]) cudaFree(s->ab[m]);
    if (s->stats[m]) cudaFree(s->stats[m]);
    s->ab[m] = nullptr;
    s->stats[m] = nullptr;
    s->nonfinite[m] = 0;
  }
}

// Builds (once per segment and metric) the nomination coefficients and the tensor map.  *ok = false when
// the segment must stay on the per-query scan (non-finite rows, tensor map not encodable).
static int batch_prepare(cudaStream_t st, Segment* s, bool cosine, bool* ok, bool* half) {
  const int m = cosine ? 1 : 0;
  std::lock_guard<std::mutex> lk(s->mu);
  *ok = false;
  *half = false;
  if (!s->tm_ok) {
    if (!batch_encode_segment_map(s->tmX, s->X, s->n, s->d, s->d, false, 256) ||
        !batch_encode_segment_map(s->tmX_b128, s->X, s->n, s->d, s->d, false, 128))
      return VS_OK;
    s->tm_ok = true;
  }
  if (!s->ab[m]) {
    void *ab = nullptr, *stv = nullptr;
    CK(cudaMalloc(&ab, (size_t)s->n * sizeof(float)), "cudaMalloc(row coefficients)");
    cudaError_t e = cudaMalloc(&stv, sizeof(SegStats));
    SegStats hs{};
    if (e == cudaSuccess) e = launch_row_prep(s->X, s->n, s->d, s->skip, cosine, static_cast<float*>(ab), static_cast<SegStats*>(stv), g_sms, st);
    if (e == cudaSuccess) e = cudaMemcpyAsync(&hs, stv, sizeof(SegStats), cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) {
      cudaFree(ab);
      if (stv) cudaFree(stv);
      return cuda_fail(e, "row_prep");
    }
    s->ab[m] = ab;
    s->stats[m] = stv;
    s->nonfinite[m] = hs.nonfinite;
    s->xmax2 = hs.xmax2_bits;
  }
  *ok = s->nonfinite[m] == 0;
  if (*ok && !s->xh_tried && g_batch_fp16.load() != 0) {
    // optional fp16 operand copy (n * dp * 2 bytes): twice the tensor rate at half the L2 traffic
    s->xh_tried = true;
    float xmax2;
    memcpy(&xmax2, &s->xmax2, 4);
    int e2 = 0;
    if (xmax2 > 0.0f) frexpf(sqrtf(xmax2) * 1.0001f, &e2);  // max |x| < 2^e2
    const float sx = xmax2 > 0.0f ? ldexpf(1.0f, 14 - e2) : 1.0f;
    const int dp = (s->d + 7) & ~7;
    void* xh = nullptr;
    if (cudaMalloc(&xh, (size_t)s->n * dp * 2) != cudaSuccess) {
      cudaGetLastError();  // not enough memory for the copy: stay on tf32 operands
    } else {
      cudaError_t e = launch_row_convert(s->X, s->n, s->d, dp, sx, xh, g_sms, st);
      if (e == cudaSuccess) e = cudaStreamSynchronize(st);
      if (e != cudaSuccess || !batch_encode_segment_map(s->tmXh, xh, s->n, s->d, dp, true, 256) ||
          !batch_encode_segment_map(s->tmXh_b128, xh, s->n, s->d, dp, true, 128)) {
        cudaFree(xh);
        if (e != cudaSuccess) return cuda_fail(e, "row_convert");
      } else {
        s->Xh = xh;
        s->dp = dp;
        s->x_scale = sx;
      }
    }
  }
  *half = s->Xh != nullptr && g_batch_fp16.load() != 0;
  return VS_OK;
}

// queries per chunk: the group-minima scratch stays within its budget (at least one 128-query block)
static int batch_chunk(const BatchLaunch& p, int nq) {
  const int64_t per_q = p.gm_stride * 4;
  int64_t c = g_batch_gm_bytes.load() / (per_q > 0 ? per_q : 1);
  c = (c / 128) * 128;
  if (c < 128) c = 128;
  return (int)(c < nq ? c : nq);
}
static size_t batch_scratch_need(const BatchLaunch& p, int nq) {
  const int c = batch_chunk(p, nq);
  const size_t cp = (size_t)((c + 255) / 256 * 256);
  const size_t gm = cp * p.gm_stride * 4;
  return Arena::need({gm, (size_t)(1 + 2 * (size_t)c) * 4, (size_t)c * batch_partial_keys(p, c) * 16,
                      p.half ? (size_t)c * p.dp * 2 : 0, p.half ? cp * 4 : 0});
}

// everything on `st`, no synchronisation; scratch = batch_scratch_need bytes, tickets for one chunk reserved
static int batch_run_dev(cudaStream_t st, const Segment* s, const BatchLaunch& p, bool cosine, const float* d_q, int nq,
                         int64_t* d_ids, double* d_scores, int32_t* d_counts, int64_t out_stride, void* scratch,
                         unsigned long long* d_ticket, bool direct = false) {
  const int m = cosine ? 1 : 0;
  const int chunk = batch_chunk(p, nq);
  const int64_t os = out_stride > 0 ? out_stride : p.k;
  for (int q0 = 0; q0 < nq; q0 += chunk) {
    const int c = q0 + chunk <= nq ? chunk : nq - q0;
    Arena A(scratch);
    BatchLaunch L = p;
    L.X = s->X; L.skip = s->skip; L.q = d_q + (size_t)q0 * s->d; L.nq = c;
    L.tmX = p.half ? s->tmXh : s->tmX; L.tmX128 = p.half ? s->tmXh_b128 : s->tmX_b128; L.x_scale = s->x_scale;
    L.pairs = g_batch_pairs.load() == 1 || (g_batch_pairs.load() == 2 && !p.pair_stat);
    L.xh = s->Xh; L.prefilter = g_batch_prefilter.load() != 0;
    L.direct = direct; L.reserve_sms = (int)g_scan_reserve_sms.load();
    L.coef = static_cast<const float*>(s->ab[m]); L.stats = static_cast<const SegStats*>(s->stats[m]);
    L.gm = A.take<float>((size_t)((chunk + 255) / 256 * 256) * p.gm_stride);
    L.fb = A.take<int32_t>(1 + 2 * (size_t)chunk);
    L.partial_keys = batch_partial_keys(p, c);
    L.partial = A.take<ulonglong2>((size_t)chunk * batch_partial_keys(p, chunk));
    L.qh = p.half ? A.take<char>((size_t)chunk * p.dp * 2) : nullptr;
    L.qinv = p.half ? A.take<float>((size_t)((chunk + 255) / 256 * 256)) : nullptr;
    L.ctrl = d_ticket;
    L.ids_out = d_ids + (size_t)q0 * os; L.scores_out = d_scores + (size_t)q0 * os; L.counts_out = d_counts + q0;
    L.id_base = s->id_base; L.out_stride = os; L.warp_min_q = (int)g_batch_warp_min_q.load();
    CK(launch_batch(L, st), "batched scan launch");
  }
  return VS_OK;
}

static int check_query_args(const Segment* s, const void* q, int nq, int k, int metric) {
  if (!s) return fail(VS_EHANDLE, "unknown segment handle");
  if (!q || nq <= 0) return fail(VS_EINVAL, "q must be non-null and nq positive");
  if (k <= 0 || k > TOPK_MAX_K) return fail(VS_EINVAL, "k must be in 1..%d", TOPK_MAX_K);
  if (metric != VS_METRIC_L2 && metric != VS_METRIC_COSINE) return fail(VS_EINVAL, "unknown metric %d", metric);
  return VS_OK;
}

}  // namespace vs

using namespace vs;

// =================================================================================================
// lifecycle
// =================================================================================================
extern "C" {

int32_t vs_version(void) { return 100; }
const char* vs_last_error(void) { return t_err.c_str(); }

// Binds the devices this process drives.  One device: everything runs on it.  Several (one JVM, all GPUs of the box):
// slot 0 is the primary (pair operations, host-row builds), vs_segment_upload shards rows across all of them and the
// query / build entry points fan out on one worker thread per GPU (group.cu) -- no torch, no second process.
int32_t vs_init_multi(int32_t n_gpus, const int32_t* device_ids) {
  if (n_gpus < 1 || n_gpus > VS_MAX_DEVICES || !device_ids) return fail(VS_EINVAL, "n_gpus must be in 1..%d", VS_MAX_DEVICES);
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count <= 0) {
    cudaGetLastError();
    return fail(VS_ECUDA, "no usable CUDA device (%s); libvsgpu has no CPU fallback",
                e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
  }
  int sms = 0;
  for (int i = 0; i < n_gpus; i++) {
    const int device = device_ids[i];
    if (device < 0 || device >= count) return fail(VS_EINVAL, "device %d out of range (0..%d)", device, count - 1);
    if (device >= VS_MAX_DEVICES) return fail(VS_EINVAL, "device ordinal %d is beyond the %d this build supports", device, VS_MAX_DEVICES);
    // (a device may be listed more than once: several ranks then share it -- slower, but the whole coordinator runs
    //  on a one-GPU box, which is how the tests cover it)
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, device), "cudaGetDeviceProperties");
    if (prop.major < 10) return fail(VS_ECUDA, "device %d is sm_%d%d; libvsgpu is built for sm_100a only", device, prop.major, prop.minor);
    if (i > 0 && prop.multiProcessorCount != sms) return fail(VS_EINVAL, "the devices of a group must be identical (SM counts differ)");
    sms = prop.multiProcessorCount;
  }
  // re-binding: a previous group (its workers, communicators and sharded segments) goes first
  if (group_size() > 0) RET(group_stop());
  // a device that stays bound keeps its resident segments; others must not leave segments behind
  {
    std::lock_guard<std::mutex> lk(g_mu);
    for (auto& kv : g_segments) {
      bool stays = false;
      for (int i = 0; i < n_gpus; i++) stays = stays || device_ids[i] == kv.second->device;
      if (!stays) return fail(VS_ESTATE, "device %d still holds resident segments: free them (or vs_shutdown) before re-binding", kv.second->device);
    }
  }
  CK(cudaSetDevice(device_ids[0]), "cudaSetDevice");
  g_sms = sms;
  for (int i = 0; i < n_gpus; i++) g_devs[i] = device_ids[i];
  g_ndev.store(n_gpus);
  g_device.store(device_ids[0]);
  if (n_gpus > 1) {
    int r = group_start(n_gpus, device_ids);
    if (r != VS_OK) {  // fall back to nothing: the caller sees the failure
      g_ndev.store(1);
      return r;
    }
  }
  return VS_OK;
}

int32_t vs_init(int32_t device) { return vs_init_multi(1, &device); }

int32_t vs_device_count(void) { return g_ndev.load(); }

int32_t vs_shutdown(void) {
  residency_clear();
  if (group_size() > 0) group_stop();
  std::vector<std::shared_ptr<Segment>> segs;
  {
    std::lock_guard<std::mutex> lk(g_mu);
    for (auto& kv : g_segments) segs.push_back(std::move(kv.second));
    g_segments.clear();
  }
  segs.clear();  // seg_destroy drains each segment's device first
  return VS_OK;
}

int32_t vs_set_simd_lanes(int32_t lanes_) {
  if (lanes_ != 16 && lanes_ != 8 && lanes_ != 4) return fail(VS_EINVAL, "lanes must be 16, 8 or 4");
  g_lanes.store(lanes_);
  return VS_OK;
}
int32_t vs_get_simd_lanes(void) { return g_lanes.load(); }

int32_t vs_device_info(int32_t* sm_count_, int64_t* free_bytes, int64_t* total_bytes) {
  ThreadCtx* c;
  RET(ctx_bind(&c));
  size_t f = 0, t = 0;
  CK(cudaMemGetInfo(&f, &t), "cudaMemGetInfo");
  if (sm_count_) *sm_count_ = g_sms;
  if (free_bytes) *free_bytes = (int64_t)f;
  if (total_bytes) *total_bytes = (int64_t)t;
  return VS_OK;
}

int64_t vs_kernel_launch_count(void) { return g_launches.load(); }

#ifdef VS_DEBUG_EXPORTS  // development builds only (make EXTRA=-DVS_DEBUG_EXPORTS): not part of the product ABI
int32_t vs_debug_adc_stats(uint32_t* out8) { return vs::debug_adc_stats(out8) == 0 ? VS_OK : VS_ECUDA; }
#endif

int32_t vs_set_option(const char* name, int64_t value) {
  if (!name) return fail(VS_EINVAL, "null option name");
  if (!strcmp(name, "peer_fused")) {
    g_peer_fused.store(value != 0 ? 1 : 0);
    return VS_OK;
  }
  if (!strcmp(name, "peer_spin_shared")) {
    g_peer_spin_shared.store(value != 0 ? 1 : 0);
    return VS_OK;
  }
  if (!strcmp(name, "train_exact_order")) {
    group_set_train_exact(value != 0);
    return VS_OK;
  }
  if (!strcmp(name, "pq_tc_keep_bytes")) {
    if (value < 0) return fail(VS_EINVAL, "pq_tc_keep_bytes must be >= 0");
    pq_tc_set_keep_bytes((unsigned long long)value);
    return VS_OK;
  }
  if (!strcmp(name, "pq_tensor_cores")) {
    if (value < 0 || value > 2) return fail(VS_EINVAL, "pq_tensor_cores must be 0 (FFMA), 1 (mma.sync) or 2 (tcgen05)");
    pq_set_tensor_cores((int)value);
    return VS_OK;
  }
  if (!strcmp(name, "scan_reserve_sms")) {
    if (value < 0 || value >= 128) return fail(VS_EINVAL, "scan_reserve_sms must be in 0..127");
    g_scan_reserve_sms.store(value);
    return VS_OK;
  }
  if (!strcmp(name, "pdl")) {
    g_pdl.store(value != 0);
    return VS_OK;
  }
  if (!strcmp(name, "batch_prefilter")) {
    g_batch_prefilter.store(value != 0);
    return VS_OK;
  }
  if (!strcmp(name, "batch_pairs")) {
    if (value < 0 || value > 2) return fail(VS_EINVAL, "batch_pairs must be 0 (never), 1 (always) or 2 (long vectors only)");
    g_batch_pairs.store(value);
    return VS_OK;
  }
  if (!strcmp(name, "batch_prefetch_rounds")) {  // this device only (diagnostics)
    if (value < -1 || value > 64) return fail(VS_EINVAL, "batch_prefetch_rounds must be -1 .. 64");
    CK(cudaError_t(batch_set_prefetch_rounds((int)value)), "batch_prefetch_rounds");
    return VS_OK;
  }
  if (!strcmp(name, "scan_fp16")) {
    g_scan_fp16.store(value != 0);
    return VS_OK;
  }
  if (!strcmp(name, "adc_reserve_sms")) {
    if (value < 0 || value >= 128) return fail(VS_EINVAL, "adc_reserve_sms must be in 0..127");
    g_adc_reserve_sms.store(value);
    return VS_OK;
  }
  if (!strcmp(name, "scan_half_ctas")) {
    if (value < 0 || value > 2) return fail(VS_EINVAL, "scan_half_ctas must be 0 (automatic), 1 or 2");
    g_scan_half_ctas.store(value);
    return VS_OK;
  }
  if (!strcmp(name, "batch_select_ctas")) {
    if (value < 0 || value > 64) return fail(VS_EINVAL, "batch_select_ctas must be 0 (automatic) .. 64");
    g_batch_select_ctas.store(value);
    return VS_OK;
  }
  if (!strcmp(name, "batch_group")) {
    if (value != 0 && value != 16 && value != 32 && value != 64) return fail(VS_EINVAL, "batch_group must be 0, 16, 32 or 64");
    g_batch_group.store(value);
    return VS_OK;
  }
  if (!strcmp(name, "batch_warp_min_queries")) {
    if (value < 0) return fail(VS_EINVAL, "batch_warp_min_queries must be >= 0");
    g_batch_warp_min_q.store(value);
    return VS_OK;
  }
  if (!strcmp(name, "batch_fp16")) {  // applies to segments that have not been queried in batches yet
    g_batch_fp16.store(value != 0);
    return VS_OK;
  }
  if (!strcmp(name, "batch_min_queries") || !strcmp(name, "batch_min_rows") || !strcmp(name, "batch_gm_bytes")) {
    if (value < 1) return fail(VS_EINVAL, "%s must be >= 1", name);
    (name[6] == 'm' && name[10] == 'q' ? g_batch_min_queries : (name[6] == 'm' ? g_batch_min_rows : g_batch_gm_bytes)).store(value);
    return VS_OK;
  }
  if (!strcmp(name, "adc_fast_min_rows")) {
    if (value < 0) return fail(VS_EINVAL, "adc_fast_min_rows must be >= 0");
    g_adc_fast_min_rows.store(value);
    return VS_OK;
  }
  if (!strcmp(name, "adc_fast_cap")) {
    if (value < 1 || value > (int64_t(1) << 24)) return fail(VS_EINVAL, "adc_fast_cap must be in 1..2^24");
    g_adc_fast_cap.store(value);
    return VS_OK;
  }
  return fail(VS_EINVAL, "unknown option %s", name);
}
#ifdef VS_BQ_STAMPS
int32_t vs_debug_read_stamps_batch(void* dst, int64_t bytes) { return vs::debug_read_stamps_batch(dst, (size_t)bytes); }
#endif
#ifdef VS_PHASE_STAMPS
int32_t vs_debug_read_stamps(void* dst, int64_t bytes) { return vs::debug_read_stamps(dst, (size_t)bytes); }
int32_t vs_debug_read_stamps_adc(void* dst, int64_t bytes) { return vs::debug_read_stamps_adc(dst, (size_t)bytes); }
#endif

// =================================================================================================
// pair operations
// =================================================================================================
static int32_t pair_op(int op, const float* a, const float* b, int32_t len, double* out) {
  if (!a || (!b && op != PAIR_NORM) || !out || len < 0) return fail(VS_EINVAL, "null pointer or negative length");
  ThreadCtx* c;
  RET(ctx_bind(&c));
  const size_t fb = (size_t)len * 4;
  RET(ctx_reserve_dev(c, Arena::need({fb, fb, 8})));
  RET(ctx_reserve_host(c, 8));
  Arena A(c->d_buf);
  float* da = A.take<float>(len);
  float* db = A.take<float>(len);
  double* dout = A.take<double>(1);
  if (len > 0) {
    CK(cudaMemcpyAsync(da, a, fb, cudaMemcpyHostToDevice, c->stream), "H2D a");
    if (b) CK(cudaMemcpyAsync(db, b, fb, cudaMemcpyHostToDevice, c->stream), "H2D b");
  }
  CK(launch_pair(op, da, b ? db : da, len, lanes(), dout, c->stream), "pair launch");
  CK(cudaMemcpyAsync(c->h_buf, dout, 8, cudaMemcpyDeviceToHost, c->stream), "D2H");
  CK(cudaStreamSynchronize(c->stream), "sync");
  *out = *static_cast<double*>(c->h_buf);
  return VS_OK;
}
int32_t vs_l2(const float* a, const float* b, int32_t len, double* out) { return pair_op(PAIR_L2, a, b, len, out); }
int32_t vs_l2_squared(const float* a, const float* b, int32_t len, double* out) { return pair_op(PAIR_L2SQ, a, b, len, out); }
int32_t vs_dot(const float* a, const float* b, int32_t len, double* out) { return pair_op(PAIR_DOT, a, b, len, out); }
int32_t vs_norm(const float* a, int32_t len, double* out) { return pair_op(PAIR_NORM, a, nullptr, len, out); }
int32_t vs_cosine(const float* a, const float* b, int32_t len, double* out) { return pair_op(PAIR_COSINE, a, b, len, out); }

static int check_pq_shape(int M, int K, int subDim) {
  if (M <= 0 || K <= 0 || subDim <= 0) return fail(VS_EINVAL, "Invalid PQ params (m,k,dimension)");
  return VS_OK;
}

int32_t vs_pq_encode(const float* centroids, int32_t M, int32_t K, int32_t subDim, const float* v,
                     uint8_t* codes_out) {
  return vs_pq_encode_batch(centroids, M, K, subDim, v, 0, 1, codes_out);
}

int32_t vs_pq_lut_distance(const float* lut, int32_t M, int32_t K, const uint8_t* codes, float* out) {
  if (!lut || !codes || !out) return fail(VS_EINVAL, "null pointer");
  if (M <= 0 || K <= 0) return fail(VS_EINVAL, "M and K must be positive");
  for (int m = 0; m < M; m++)
    if (codes[m] >= K) return fail(VS_EINVAL, "code %d of subspace %d is outside the LUT (K=%d)", (int)codes[m], m, K);
  ThreadCtx* c;
  RET(ctx_bind(&c));
  const size_t lb = (size_t)M * K * 4;
  RET(ctx_reserve_dev(c, Arena::need({lb, (size_t)M, 4})));
  RET(ctx_reserve_host(c, 4));
  Arena A(c->d_buf);
  float* dl = A.take<float>((size_t)M * K);
  uint8_t* dc = A.take<uint8_t>(M);
  float* dout = A.take<float>(1);
  CK(cudaMemcpyAsync(dl, lut, lb, cudaMemcpyHostToDevice, c->stream), "H2D lut");
  CK(cudaMemcpyAsync(dc, codes, M, cudaMemcpyHostToDevice, c->stream), "H2D codes");
  CK(launch_lut_distance_f32(dl, M, K, dc, dout, c->stream), "lut distance launch");
  CK(cudaMemcpyAsync(c->h_buf, dout, 4, cudaMemcpyDeviceToHost, c->stream), "D2H");
  CK(cudaStreamSynchronize(c->stream), "sync");
  *out = *static_cast<float*>(c->h_buf);
  return VS_OK;
}

int32_t vs_build_lut(const float* centroids, int32_t M, int32_t K, int32_t subDim, const float* q,
                     double* lut_out) {
  if (!centroids || !q || !lut_out) return fail(VS_EINVAL, "null pointer");
  RET(check_pq_shape(M, K, subDim));
  ThreadCtx* c;
  RET(ctx_bind(&c));
  const size_t cb = (size_t)M * K * subDim * 4, qb = (size_t)M * subDim * 4, lb = (size_t)M * K * 8;
  RET(ctx_reserve_dev(c, Arena::need({cb, qb, lb})));
  Arena A(c->d_buf);
  float* dc = A.take<float>((size_t)M * K * subDim);
  float* dq = A.take<float>((size_t)M * subDim);
  double* dl = A.take<double>((size_t)M * K);
  CK(cudaMemcpyAsync(dc, centroids, cb, cudaMemcpyHostToDevice, c->stream), "H2D centroids");
  CK(cudaMemcpyAsync(dq, q, qb, cudaMemcpyHostToDevice, c->stream), "H2D q");
  CK(launch_build_lut(dc, M, K, subDim, dq, 1, lanes(), dl, c->stream), "build_lut launch");
  CK(cudaMemcpyAsync(lut_out, dl, lb, cudaMemcpyDeviceToHost, c->stream), "D2H lut");
  CK(cudaStreamSynchronize(c->stream), "sync");
  return VS_OK;
}

int32_t vs_pq_approx_distance(const double* lut, int32_t M, int32_t K, const uint8_t* codes, int64_t n,
                              double* out) {
  if (!lut || (!codes && n > 0) || (!out && n > 0) || n < 0) return fail(VS_EINVAL, "null pointer or negative n");
  if (M <= 0 || K <= 0) return fail(VS_EINVAL, "M and K must be positive");
  if (n == 0) return VS_OK;
  ThreadCtx* c;
  RET(ctx_bind(&c));
  const size_t lb = (size_t)M * K * 8, cb = (size_t)n * M, ob = (size_t)n * 8;
  RET(ctx_reserve_dev(c, Arena::need({lb, cb, ob})));
  Arena A(c->d_buf);
  double* dl = A.take<double>((size_t)M * K);
  uint8_t* dc = A.take<uint8_t>(cb);
  double* dout = A.take<double>(n);
  CK(cudaMemcpyAsync(dl, lut, lb, cudaMemcpyHostToDevice, c->stream), "H2D lut");
  CK(cudaMemcpyAsync(dc, codes, cb, cudaMemcpyHostToDevice, c->stream), "H2D codes");
  CK(launch_approx_distance(dl, M, K, dc, n, dout, c->stream), "approx distance launch");
  CK(cudaMemcpyAsync(out, dout, ob, cudaMemcpyDeviceToHost, c->stream), "D2H");
  CK(cudaStreamSynchronize(c->stream), "sync");
  return VS_OK;
}

// =================================================================================================
// segment residency
// =================================================================================================
static int seg_new(int64_t n, int32_t d, int64_t id_base, Segment** out) {
  if (n < 0 || d <= 0) return fail(VS_EINVAL, "n must be >= 0 and d positive");
  Segment* s = new Segment();
  s->n = n;
  s->d = d;
  s->id_base = id_base;
  cudaGetDevice(&s->device);  // the calling thread's context is bound (ctx_bind)
  if (n > 0) {
    cudaError_t e = cudaMalloc(&s->X, (size_t)n * d * 4);
    if (e != cudaSuccess) {
      delete s;
      return cuda_fail(e, "cudaMalloc(rows)");
    }
  }
  *out = s;
  return VS_OK;
}

static int seg_set_skip(ThreadCtx* c, Segment* s, const uint8_t* skip_mask) {
  {  // the skip mask is folded into the batched path's row coefficients: rebuild them at next use
    std::lock_guard<std::mutex> lk(s->mu);
    if (s->ab[0] || s->ab[1]) {
      CK(cudaDeviceSynchronize(), "sync");
      batch_invalidate(s);
    }
  }
  if (!skip_mask) {
    if (s->skip) {
      CK(cudaStreamSynchronize(c->stream), "sync");
      cudaFree(s->skip);
      s->skip = nullptr;
    }
    return VS_OK;
  }
  if (s->n == 0) return VS_OK;
  if (!s->skip) CK(cudaMalloc(&s->skip, (size_t)s->n), "cudaMalloc(skip)");
  CK(cudaMemcpyAsync(s->skip, skip_mask, (size_t)s->n, cudaMemcpyHostToDevice, c->stream), "H2D skip");
  CK(cudaStreamSynchronize(c->stream), "sync");
  return VS_OK;
}

int32_t vs_segment_upload(const float* rows, int64_t n, int32_t d, const uint8_t* skip_mask,
                          int64_t id_base, uint64_t* handle_out) {
  if (!handle_out || (!rows && n > 0)) return fail(VS_EINVAL, "null pointer");
  if (group_wants_sharding()) return group_segment_upload(rows, n, d, skip_mask, id_base, handle_out);
  ThreadCtx* c;
  RET(ctx_bind(&c));
  Segment* s;
  RET(seg_new(n, d, id_base, &s));
  if (n > 0) {
    cudaError_t e = cudaMemcpyAsync(s->X, rows, (size_t)n * d * 4, cudaMemcpyHostToDevice, c->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
    if (e != cudaSuccess) {
      seg_destroy(s);
      return cuda_fail(e, "H2D rows");
    }
  }
  int r = seg_set_skip(c, s, skip_mask);
  if (r != VS_OK) {
    seg_destroy(s);
    return r;
  }
  *handle_out = seg_register(s);
  return VS_OK;
}

// Rows as they sit in the reference's storage: packed little-endian fp32 (FloatPacker.floatsToBytes,
// J/util/FloatPacker.java:21-25 -- the bytes of VectorRecord.embedding, vectorsearch.proto:114-117), `stride` bytes
// from one record's embedding to the next (>= d * 4).  The bytes ARE the device layout: one strided copy, no decode.
int32_t vs_segment_upload_strided(const uint8_t* bytes, int64_t n, int32_t d, int64_t stride, const uint8_t* skip_mask,
                                  int64_t id_base, uint64_t* handle_out) {
  if (!handle_out || (!bytes && n > 0)) return fail(VS_EINVAL, "null pointer");
  if (d <= 0 || n < 0 || stride < (int64_t)d * 4) return fail(VS_EINVAL, "need d > 0, n >= 0 and stride >= d * 4");
  if (group_wants_sharding()) return group_segment_upload_strided(bytes, n, d, stride, skip_mask, id_base, handle_out);
  ThreadCtx* c;
  RET(ctx_bind(&c));
  Segment* s;
  RET(seg_new(n, d, id_base, &s));
  if (n > 0) {
    cudaError_t e = cudaMemcpy2DAsync(s->X, (size_t)d * 4, bytes, (size_t)stride, (size_t)d * 4, (size_t)n, cudaMemcpyHostToDevice, c->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
    if (e != cudaSuccess) {
      seg_destroy(s);
      return cuda_fail(e, "H2D rows (strided)");
    }
  }
  int r = seg_set_skip(c, s, skip_mask);
  if (r != VS_OK) {
    seg_destroy(s);
    return r;
  }
  *handle_out = seg_register(s);
  return VS_OK;
}

// ---- serialized VectorRecord messages (vectorsearch.proto:108-127) ---------------------------------------------
// proto3 wire format: tag = (field << 3) | type; type 0 varint, 1 fixed64, 2 length-delimited, 5 fixed32.
namespace {
inline bool pb_varint(const uint8_t*& p, const uint8_t* end, uint64_t* v) {
  uint64_t r = 0;
  for (int shift = 0; shift < 64 && p < end; shift += 7) {
    const uint8_t b = *p++;
    r |= (uint64_t)(b & 0x7f) << shift;
    if (!(b & 0x80)) {
      *v = r;
      return true;
    }
  }
  return false;
}
struct RecordView {
  const uint8_t* emb = nullptr;
  uint64_t emb_len = 0;
  int32_t vec_id = 0;
  bool deleted = false;
};
// walks one VectorRecord: seg_id = 1, vec_id = 2, embedding = 3 (bytes), deleted = 4 (bool), payload = 5; unknown
// fields are skipped as protobuf requires; the LAST occurrence of a field wins (protobuf merge semantics)
inline bool pb_vector_record(const uint8_t* p, const uint8_t* end, RecordView* out) {
  while (p < end) {
    uint64_t tag;
    if (!pb_varint(p, end, &tag)) return false;
    const uint32_t field = (uint32_t)(tag >> 3), type = (uint32_t)(tag & 7);
    uint64_t v = 0;
    switch (type) {
      case 0:
        if (!pb_varint(p, end, &v)) return false;
        if (field == 2) out->vec_id = (int32_t)v;
        if (field == 4) out->deleted = v != 0;
        break;
      case 1:
        if (end - p < 8) return false;
        p += 8;
        break;
      case 2:
        if (!pb_varint(p, end, &v) || (uint64_t)(end - p) < v) return false;
        if (field == 3) {
          out->emb = p;
          out->emb_len = v;
        }
        p += v;
        break;
      case 5:
        if (end - p < 4) return false;
        p += 4;
        break;
      default:
        return false;  // groups are not used by this schema
    }
  }
  return true;
}
}  // namespace

// Consumes the stored VectorRecord values of a segment as they come out of the range read
// (J/fdb/FdbVectorIndex.java:676-699): record i is buf[offsets[i] .. offsets[i+1]).  Embeddings go to the device through
// pinned staging in slabs (no float[] is ever materialised), `deleted` becomes the row's skip flag (:681), vec_ids_out
// (nullable, [n]) receives each row's vec_id so that the caller can map rows back to gids.
int32_t vs_segment_upload_records(const uint8_t* buf, const int64_t* offsets, int64_t n, int32_t d, int64_t id_base,
                                  int32_t* vec_ids_out, uint64_t* handle_out) {
  if (!handle_out || ((!buf || !offsets) && n > 0)) return fail(VS_EINVAL, "null pointer");
  if (d <= 0 || n < 0) return fail(VS_EINVAL, "need d > 0 and n >= 0");
  if (group_wants_sharding()) return group_segment_upload_records(buf, offsets, n, d, id_base, vec_ids_out, handle_out);
  ThreadCtx* c;
  RET(ctx_bind(&c));
  Segment* s;
  RET(seg_new(n, d, id_base, &s));
  std::vector<uint8_t> skip((size_t)n, 0);
  bool any_deleted = false;
  const size_t row_b = (size_t)d * 4;
  int64_t slab = (int64_t)((size_t(16) << 20) / row_b);
  if (slab < 1) slab = 1;
  int r = ctx_reserve_host(c, 2 * (size_t)slab * row_b);
  cudaEvent_t ev[2] = {nullptr, nullptr};
  cudaError_t e = cudaSuccess;
  for (int i = 0; i < 2 && r == VS_OK && e == cudaSuccess; i++) e = cudaEventCreateWithFlags(&ev[i], cudaEventDisableTiming);
  for (int64_t r0 = 0, b = 0; r == VS_OK && e == cudaSuccess && r0 < n; r0 += slab, b ^= 1) {
    const int64_t cnt = std::min<int64_t>(slab, n - r0);
    uint8_t* stage = static_cast<uint8_t*>(c->h_buf) + (size_t)b * slab * row_b;
    if (r0 >= 2 * slab) e = cudaEventSynchronize(ev[b]);  // the copy that last used this half has finished
    for (int64_t i = 0; i < cnt && r == VS_OK; i++) {
      const int64_t o0 = offsets[r0 + i], o1 = offsets[r0 + i + 1];
      RecordView rv;
      if (o1 < o0 || !pb_vector_record(buf + o0, buf + o1, &rv)) {
        r = fail(VS_EINVAL, "record %lld is not a valid VectorRecord message", (long long)(r0 + i));
      } else if (rv.emb_len != row_b) {
        r = fail(VS_EINVAL, "record %lld: embedding holds %llu bytes, dimension %d needs %zu", (long long)(r0 + i),
                 (unsigned long long)rv.emb_len, d, row_b);
      } else {
        memcpy(stage + (size_t)i * row_b, rv.emb, row_b);
        skip[(size_t)(r0 + i)] = rv.deleted ? 1 : 0;
        any_deleted = any_deleted || rv.deleted;
        if (vec_ids_out) vec_ids_out[r0 + i] = rv.vec_id;
      }
    }
    if (r != VS_OK || e != cudaSuccess) break;
    e = cudaMemcpyAsync(s->X + (size_t)r0 * d, stage, (size_t)cnt * row_b, cudaMemcpyHostToDevice, c->stream);
    if (e == cudaSuccess) e = cudaEventRecord(ev[b], c->stream);
  }
  if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
  for (int i = 0; i < 2; i++)
    if (ev[i]) cudaEventDestroy(ev[i]);
  if (r == VS_OK && e != cudaSuccess) r = cuda_fail(e, "H2D rows (records)");
  if (r == VS_OK && any_deleted) r = seg_set_skip(c, s, skip.data());
  if (r != VS_OK) {
    seg_destroy(s);
    return r;
  }
  *handle_out = seg_register(s);
  return VS_OK;
}

int32_t vs_segment_generate(int64_t seed, int64_t first_row, int64_t n, int32_t d, int64_t id_base,
                            uint64_t* handle_out) {
  if (!handle_out || first_row < 0) return fail(VS_EINVAL, "null pointer or negative first_row");
  if (group_wants_sharding()) return group_segment_generate(seed, first_row, n, d, id_base, handle_out);
  ThreadCtx* c;
  RET(ctx_bind(&c));
  Segment* s;
  RET(seg_new(n, d, id_base, &s));
  cudaError_t e = launch_generate(s->X, n * d, seed, first_row * d, 0, c->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
  if (e != cudaSuccess) {
    seg_destroy(s);
    return cuda_fail(e, "generate");
  }
  *handle_out = seg_register(s);
  return VS_OK;
}

int32_t vs_segment_set_skip(uint64_t h, const uint8_t* skip_mask) {
  if (group_is_sharded(h)) return group_segment_set_skip(h, skip_mask);
  std::shared_ptr<Segment> s_ref = seg_lookup(h);
  Segment* s = s_ref.get();
  if (!s) return fail(VS_EHANDLE, "unknown segment handle");
  ThreadCtx* c;
  RET(ctx_bind_seg(&c, s));
  return seg_set_skip(c, s, skip_mask);
}

int32_t vs_segment_info(uint64_t h, int64_t* n, int32_t* d, int32_t* M, int32_t* K, int64_t* id_base) {
  if (group_is_sharded(h)) return group_segment_info(h, n, d, M, K, id_base);
  std::shared_ptr<Segment> s_ref = seg_lookup(h);
  Segment* s = s_ref.get();
  if (!s) return fail(VS_EHANDLE, "unknown segment handle");
  if (n) *n = s->n;
  if (d) *d = s->d;
  if (M) *M = s->M;
  if (K) *K = s->K;
  if (id_base) *id_base = s->id_base;
  return VS_OK;
}

int32_t vs_segment_download_rows(uint64_t h, int64_t first, int64_t count, float* rows_out) {
  if (group_is_sharded(h)) return group_segment_download_rows(h, first, count, rows_out);
  std::shared_ptr<Segment> s_ref = seg_lookup(h);
  Segment* s = s_ref.get();
  if (!s) return fail(VS_EHANDLE, "unknown segment handle");
  ThreadCtx* c;
  RET(ctx_bind_seg(&c, s));
  if (first < 0 || count < 0 || first + count > s->n || (!rows_out && count > 0)) return fail(VS_EINVAL, "row range out of bounds");
  if (count == 0) return VS_OK;
  CK(cudaMemcpyAsync(rows_out, s->X + (size_t)first * s->d, (size_t)count * s->d * 4, cudaMemcpyDeviceToHost, c->stream), "D2H rows");
  CK(cudaStreamSynchronize(c->stream), "sync");
  return VS_OK;
}

int32_t vs_segment_attach_pq(uint64_t h, const float* centroids, int32_t M, int32_t K, const uint8_t* codes) {
  if (group_is_sharded(h)) return group_segment_attach_pq(h, centroids, M, K, codes);
  std::shared_ptr<Segment> s_ref = seg_lookup(h);
  Segment* s = s_ref.get();
  if (!s) return fail(VS_EHANDLE, "unknown segment handle");
  ThreadCtx* c;
  RET(ctx_bind_seg(&c, s));
  if (!centroids) return fail(VS_EINVAL, "null centroids");
  if (M <= 0 || K <= 0 || s->d % M != 0) return fail(VS_EINVAL, "Invalid PQ params (m,k,dimension)");
  const int subDim = s->d / M;
  CK(cudaStreamSynchronize(c->stream), "sync");
  // re-sealing with the same shape (a rebuilt codebook) keeps the buffers: freeing and re-allocating gigabytes of codes
  // costs more than encoding them
  const bool same_shape = s->centroids && s->M == M && s->K == K && (s->codes || s->n == 0);
  if (!same_shape) {
    if (s->centroids) cudaFree(s->centroids);
    if (s->codes) cudaFree(s->codes);
    s->centroids = nullptr;
    s->codes = nullptr;
  }
  s->M = s->K = s->subDim = 0;
  const size_t cb = (size_t)M * K * subDim * 4;
  if (!s->centroids) CK(cudaMalloc(&s->centroids, cb), "cudaMalloc(centroids)");
  CK(cudaMemcpyAsync(s->centroids, centroids, cb, cudaMemcpyHostToDevice, c->stream), "H2D centroids");
  if (s->n > 0) {
    if (!s->codes) CK(cudaMalloc(&s->codes, (size_t)s->n * M), "cudaMalloc(codes)");
    if (codes) {
      CK(cudaMemcpyAsync(s->codes, codes, (size_t)s->n * M, cudaMemcpyHostToDevice, c->stream), "H2D codes");
    } else {
      PqAssignLaunch L{};
      L.X = s->X; L.n = s->n; L.d = s->d; L.M = M; L.K = K; L.subDim = subDim; L.centroids = s->centroids;
      L.lanes = lanes(); L.codes_u8 = s->codes; L.assign_i32 = nullptr; L.s_begin = 0; L.s_end = M;
      CK(launch_pq_assign(L, c->stream), "pq assign launch");
    }
  }
  CK(cudaStreamSynchronize(c->stream), "sync");
  s->M = M;
  s->K = K;
  s->subDim = subDim;
  return VS_OK;
}

int32_t vs_segment_download_codes(uint64_t h, int64_t first, int64_t count, uint8_t* codes_out) {
  if (group_is_sharded(h)) return group_segment_download_codes(h, first, count, codes_out);
  std::shared_ptr<Segment> s_ref = seg_lookup(h);
  Segment* s = s_ref.get();
  if (!s) return fail(VS_EHANDLE, "unknown segment handle");
  ThreadCtx* c;
  RET(ctx_bind_seg(&c, s));
  if (s->M == 0) return fail(VS_ESTATE, "segment has no PQ attached");
  if (first < 0 || count < 0 || first + count > s->n || (!codes_out && count > 0)) return fail(VS_EINVAL, "row range out of bounds");
  if (count == 0) return VS_OK;
  CK(cudaMemcpyAsync(codes_out, s->codes + (size_t)first * s->M, (size_t)count * s->M, cudaMemcpyDeviceToHost, c->stream), "D2H codes");
  CK(cudaStreamSynchronize(c->stream), "sync");
  return VS_OK;
}

int32_t vs_segment_free(uint64_t h) {
  if (group_is_sharded(h)) return group_segment_free(h);
  std::shared_ptr<Segment> sp;
  {
    std::lock_guard<std::mutex> lk(g_mu);
    auto it = g_segments.find(h);
    if (it == g_segments.end()) return fail(VS_EHANDLE, "unknown segment handle");
    sp = std::move(it->second);
    g_segments.erase(it);
  }
  // the memory goes back when the last holder lets go (seg_destroy): a query still running on another thread
  // keeps the rows alive until it returns
  sp.reset();
  return VS_OK;
}

// =================================================================================================
// query operations (host pointers; blocking)
// =================================================================================================
// queries are processed in groups so that the per-CTA partial lists stay within this budget
static const size_t PARTIAL_BUDGET = size_t(192) << 20;

int32_t vs_bruteforce_topk(uint64_t h, const float* q, int32_t nq, int32_t k, int32_t metric,
                           int64_t* ids_out, double* scores_out, int32_t* counts_out) {
  if (group_is_sharded(h)) return group_bruteforce_topk(h, q, nq, k, metric, ids_out, scores_out, counts_out);
  std::shared_ptr<Segment> s_ref = seg_lookup(h);
  Segment* s = s_ref.get();
  RET(check_query_args(s, q, nq, k, metric));
  if (!ids_out || !scores_out) return fail(VS_EINVAL, "null output pointer");
  ThreadCtx* c;
  RET(ctx_bind_seg(&c, s));
  if (s->n == 0) {
    fill_empty(ids_out, scores_out, counts_out, nq, k);
    return VS_OK;
  }
  const bool cosine = metric == VS_METRIC_COSINE;
  const int bmode = batch_wanted(s, nq, cosine);
  if (bmode != 0) {  // nomination on tensor cores (batches) or on the fp16 copy (one or two queries) + exact re-score (batch.cu)
    bool ok = false, half = false;
    RET(batch_prepare(c->stream, s, cosine, &ok, &half));
    BatchLaunch bp;
    if (ok) RET(plan_batch(s, k, cosine, half, &bp));
    if (ok && bmode == 2 && !(half && bp.sh_ok)) ok = false;
    if (ok) {
      int64_t hg = (int64_t(32) << 20) / ((int64_t)s->d * 4);  // queries per host staging group
      if (hg < 128) hg = 128;
      if (hg > nq) hg = nq;
      const int group = (int)hg;
      const size_t qb = (size_t)group * s->d * 4, ib = (size_t)group * k * 8, cb = (size_t)group * 4;
      const size_t sb = batch_scratch_need(bp, group);
      RET(ctx_reserve_dev(c, Arena::need({qb, ib, ib, cb, sb})));
      RET(ctx_reserve_host(c, Arena::need({qb, ib, ib, cb})));
      RET(ctx_reserve_ticket(c, batch_chunk(bp, group)));
      for (int q0 = 0; q0 < nq; q0 += group) {
        const int g = (q0 + group <= nq) ? group : nq - q0;
        Arena A(c->d_buf), H(c->h_buf);
        float* dq = A.take<float>((size_t)group * s->d);
        int64_t* dids = A.take<int64_t>((size_t)group * k);
        double* dsc = A.take<double>((size_t)group * k);
        int32_t* dcn = A.take<int32_t>(group);
        char* scratch = A.take<char>(sb);
        float* hq = H.take<float>((size_t)group * s->d);
        int64_t* hids = H.take<int64_t>((size_t)group * k);
        double* hsc = H.take<double>((size_t)group * k);
        int32_t* hcn = H.take<int32_t>(group);
        memcpy(hq, q + (size_t)q0 * s->d, (size_t)g * s->d * 4);
        CK(cudaMemcpyAsync(dq, hq, (size_t)g * s->d * 4, cudaMemcpyHostToDevice, c->stream), "H2D q");
        if ((size_t)g * k <= 4096) {
          // short result lists: written straight into the pinned staging buffer (device-visible under UVA)
          RET(batch_run_dev(c->stream, s, bp, cosine, dq, g, hids, hsc, hcn, 0, scratch, c->d_ticket, bmode == 2));
        } else {
          RET(batch_run_dev(c->stream, s, bp, cosine, dq, g, dids, dsc, dcn, 0, scratch, c->d_ticket, bmode == 2));
          CK(cudaMemcpyAsync(hids, dids, (size_t)g * k * 8, cudaMemcpyDeviceToHost, c->stream), "D2H ids");
          CK(cudaMemcpyAsync(hsc, dsc, (size_t)g * k * 8, cudaMemcpyDeviceToHost, c->stream), "D2H scores");
          CK(cudaMemcpyAsync(hcn, dcn, (size_t)g * 4, cudaMemcpyDeviceToHost, c->stream), "D2H counts");
        }
        CK(cudaStreamSynchronize(c->stream), "sync");
        memcpy(ids_out + (size_t)q0 * k, hids, (size_t)g * k * 8);
        memcpy(scores_out + (size_t)q0 * k, hsc, (size_t)g * k * 8);
        if (counts_out) memcpy(counts_out + q0, hcn, (size_t)g * 4);
      }
      return VS_OK;
    }
  }
  ScanLaunch p;
  RET(plan_scan(s, nq, k, cosine, &p));
  const size_t per_q_partial = (size_t)p.partial_keys * 16;
  int group = (int)(PARTIAL_BUDGET / per_q_partial);
  if (group < 1) group = 1;
  if (group > nq) group = nq;
  const size_t qb = (size_t)group * s->d * 4, ib = (size_t)group * k * 8, cb = (size_t)group * 4;
  RET(ctx_reserve_dev(c, Arena::need({qb, ib, ib, cb, per_q_partial * group})));
  RET(ctx_reserve_host(c, Arena::need({qb, ib, ib, cb})));
  RET(ctx_reserve_ticket(c, group));
  for (int q0 = 0; q0 < nq; q0 += group) {
    const int g = (q0 + group <= nq) ? group : nq - q0;
    Arena A(c->d_buf), H(c->h_buf);
    float* dq = A.take<float>((size_t)group * s->d);
    int64_t* dids = A.take<int64_t>((size_t)group * k);
    double* dsc = A.take<double>((size_t)group * k);
    int32_t* dcn = A.take<int32_t>(group);
    ulonglong2* dpart = A.take<ulonglong2>((size_t)group * p.partial_keys);
    float* hq = H.take<float>((size_t)group * s->d);
    int64_t* hids = H.take<int64_t>((size_t)group * k);
    double* hsc = H.take<double>((size_t)group * k);
    int32_t* hcn = H.take<int32_t>(group);
    memcpy(hq, q + (size_t)q0 * s->d, (size_t)g * s->d * 4);
    CK(cudaMemcpyAsync(dq, hq, (size_t)g * s->d * 4, cudaMemcpyHostToDevice, c->stream), "H2D q");
    if ((size_t)g * k <= 4096) {
      // short result lists: the last CTA writes ids / scores / counts straight into the pinned staging buffer
      // (device-visible under UVA) -- three device-to-host copies less on the latency path of a query
      RET(bruteforce_dev(c, c->stream, s, dq, g, k, metric, hids, hsc, hcn, dpart, c->d_ticket, p));
      CK(cudaStreamSynchronize(c->stream), "sync");
    } else {
      RET(bruteforce_dev(c, c->stream, s, dq, g, k, metric, dids, dsc, dcn, dpart, c->d_ticket, p));
      CK(cudaMemcpyAsync(hids, dids, (size_t)g * k * 8, cudaMemcpyDeviceToHost, c->stream), "D2H ids");
      CK(cudaMemcpyAsync(hsc, dsc, (size_t)g * k * 8, cudaMemcpyDeviceToHost, c->stream), "D2H scores");
      CK(cudaMemcpyAsync(hcn, dcn, (size_t)g * 4, cudaMemcpyDeviceToHost, c->stream), "D2H counts");
      CK(cudaStreamSynchronize(c->stream), "sync");
    }
    memcpy(ids_out + (size_t)q0 * k, hids, (size_t)g * k * 8);
    memcpy(scores_out + (size_t)q0 * k, hsc, (size_t)g * k * 8);
    if (counts_out) memcpy(counts_out + q0, hcn, (size_t)g * 4);
  }
  return VS_OK;
}

static int adc_common(uint64_t h, const float* q, int32_t nq, int32_t n_cand, bool rerank, int32_t k,
                      int32_t metric, int64_t* ids_out, double* scores_out, int32_t* counts_out) {
  std::shared_ptr<Segment> s_ref = seg_lookup(h);
  Segment* s = s_ref.get();
  RET(check_query_args(s, q, nq, n_cand, rerank ? metric : VS_METRIC_L2));
  if (rerank && (k <= 0 || k > TOPK_MAX_K)) return fail(VS_EINVAL, "k must be in 1..%d", TOPK_MAX_K);
  if (!ids_out || !scores_out) return fail(VS_EINVAL, "null output pointer");
  if (s->M == 0) return fail(VS_ESTATE, "segment has no PQ attached");
  ThreadCtx* c;
  RET(ctx_bind_seg(&c, s));
  const int kout = rerank ? k : n_cand;
  if (s->n == 0) {
    fill_empty(ids_out, scores_out, counts_out, nq, kout);
    return VS_OK;
  }
  AdcPlan p;
  RET(plan_adc(s, nq, n_cand, &p));
  const size_t per_q_partial = (size_t)p.partial_keys * 16, per_q_cand = (p.cand_entries + p.extra_words) * 8;
  int group = (int)(PARTIAL_BUDGET / (per_q_partial + per_q_cand));
  if (group < 1) group = 1;
  if (group > nq) group = nq;
  const size_t qb = (size_t)group * s->d * 4, cib = (size_t)group * n_cand * 8, ccb = (size_t)group * 4;
  const size_t ob = (size_t)group * kout * 8, lb = (size_t)group * s->M * s->K * 8;
  RET(ctx_reserve_dev(c, Arena::need({qb, cib, cib, ccb, ob, ob, ccb, lb, per_q_partial * group, per_q_cand * group})));
  RET(ctx_reserve_host(c, Arena::need({qb, ob, ob, ccb})));
  RET(ctx_reserve_ticket(c, group));
  if (p.fast) RET(ctx_reserve_fs(c, group));
  for (int q0 = 0; q0 < nq; q0 += group) {
    const int g = (q0 + group <= nq) ? group : nq - q0;
    Arena A(c->d_buf), H(c->h_buf);
    float* dq = A.take<float>((size_t)group * s->d);
    int64_t* dcid = A.take<int64_t>((size_t)group * n_cand);
    double* dcap = A.take<double>((size_t)group * n_cand);
    int32_t* dccn = A.take<int32_t>(group);
    int64_t* dids = A.take<int64_t>((size_t)group * kout);
    double* dsc = A.take<double>((size_t)group * kout);
    int32_t* dcn = A.take<int32_t>(group);
    double* dlut = A.take<double>((size_t)group * s->M * s->K);
    ulonglong2* dpart = A.take<ulonglong2>((size_t)group * p.partial_keys);
    unsigned long long* dcand = A.take<unsigned long long>((size_t)group * (p.cand_entries + p.extra_words));
    float* hq = H.take<float>((size_t)group * s->d);
    int64_t* hids = H.take<int64_t>((size_t)group * kout);
    double* hsc = H.take<double>((size_t)group * kout);
    int32_t* hcn = H.take<int32_t>(group);
    memcpy(hq, q + (size_t)q0 * s->d, (size_t)g * s->d * 4);
    CK(cudaMemcpyAsync(dq, hq, (size_t)g * s->d * 4, cudaMemcpyHostToDevice, c->stream), "H2D q");
    RET(adc_dev(c->stream, s, dq, g, n_cand, dlut, dcid, dcap, dccn, dpart, c->d_ticket, c->d_fs, dcand, p));
    const int64_t* src_ids = dcid;
    const double* src_sc = dcap;
    const int32_t* src_cn = dccn;
    bool direct = false;
    if (rerank) {
      // short lists: the re-rank kernel writes straight into the pinned staging buffer (device-visible under UVA)
      direct = (size_t)g * kout <= 4096;
      RET(rerank_dev(c->stream, s, dq, g, dcid, n_cand, k, metric, direct ? hids : dids, direct ? hsc : dsc, direct ? hcn : dcn));
      src_ids = dids;
      src_sc = dsc;
      src_cn = dcn;
    }
    if (!direct) {
      CK(cudaMemcpyAsync(hids, src_ids, (size_t)g * kout * 8, cudaMemcpyDeviceToHost, c->stream), "D2H ids");
      CK(cudaMemcpyAsync(hsc, src_sc, (size_t)g * kout * 8, cudaMemcpyDeviceToHost, c->stream), "D2H scores");
      CK(cudaMemcpyAsync(hcn, src_cn, (size_t)g * 4, cudaMemcpyDeviceToHost, c->stream), "D2H counts");
    }
    CK(cudaStreamSynchronize(c->stream), "sync");
    memcpy(ids_out + (size_t)q0 * kout, hids, (size_t)g * kout * 8);
    memcpy(scores_out + (size_t)q0 * kout, hsc, (size_t)g * kout * 8);
    if (counts_out) memcpy(counts_out + q0, hcn, (size_t)g * 4);
  }
  return VS_OK;
}

int32_t vs_adc_topk(uint64_t h, const float* q, int32_t nq, int32_t n_cand, int64_t* ids_out,
                    double* approx_out, int32_t* counts_out) {
  if (group_is_sharded(h)) return group_adc_topk(h, q, nq, n_cand, ids_out, approx_out, counts_out);
  return adc_common(h, q, nq, n_cand, false, 0, VS_METRIC_L2, ids_out, approx_out, counts_out);
}

int32_t vs_adc_rerank_topk(uint64_t h, const float* q, int32_t nq, int32_t n_cand, int32_t k, int32_t metric,
                           int32_t normalize_on_read, int64_t* ids_out, double* scores_out,
                           int32_t* counts_out) {
  if (group_is_sharded(h))
    return group_adc_rerank_topk(h, q, nq, n_cand, k, metric, normalize_on_read, ids_out, scores_out, counts_out);
  (void)normalize_on_read;  // same arithmetic either way (J/fdb/FdbVectorIndex.java:1006-1012)
  return adc_common(h, q, nq, n_cand, true, k, metric, ids_out, scores_out, counts_out);
}

int32_t vs_rerank_topk(uint64_t h, const float* q, const int64_t* cand_ids, int32_t n_cand, int32_t k,
                       int32_t metric, int32_t normalize_on_read, int64_t* ids_out, double* scores_out,
                       int32_t* count_out) {
  if (group_is_sharded(h))
    return group_rerank_topk(h, q, cand_ids, n_cand, k, metric, normalize_on_read, ids_out, scores_out, count_out);
  (void)normalize_on_read;
  std::shared_ptr<Segment> s_ref = seg_lookup(h);
  Segment* s = s_ref.get();
  if (!s) return fail(VS_EHANDLE, "unknown segment handle");
  if (!q || !ids_out || !scores_out || (!cand_ids && n_cand > 0)) return fail(VS_EINVAL, "null pointer");
  if (n_cand < 0 || n_cand > 8192) return fail(VS_EINVAL, "n_cand must be in 0..8192");
  if (k <= 0 || k > TOPK_MAX_K) return fail(VS_EINVAL, "k must be in 1..%d", TOPK_MAX_K);
  if (metric != VS_METRIC_L2 && metric != VS_METRIC_COSINE) return fail(VS_EINVAL, "unknown metric %d", metric);
  ThreadCtx* c;
  RET(ctx_bind_seg(&c, s));
  if (n_cand == 0 || s->n == 0) {
    fill_empty(ids_out, scores_out, count_out, 1, k);
    return VS_OK;
  }
  const size_t qb = (size_t)s->d * 4, cb = (size_t)n_cand * 8, ob = (size_t)k * 8;
  RET(ctx_reserve_dev(c, Arena::need({qb, cb, ob, ob, 4})));
  RET(ctx_reserve_host(c, Arena::need({qb, cb, ob, ob, 4})));
  Arena A(c->d_buf), H(c->h_buf);
  float* dq = A.take<float>(s->d);
  int64_t* dcand = A.take<int64_t>(n_cand);
  int64_t* dids = A.take<int64_t>(k);
  double* dsc = A.take<double>(k);
  int32_t* dcn = A.take<int32_t>(1);
  float* hq = H.take<float>(s->d);
  int64_t* hcand = H.take<int64_t>(n_cand);
  int64_t* hids = H.take<int64_t>(k);
  double* hsc = H.take<double>(k);
  int32_t* hcn = H.take<int32_t>(1);
  memcpy(hq, q, qb);
  memcpy(hcand, cand_ids, cb);
  CK(cudaMemcpyAsync(dq, hq, qb, cudaMemcpyHostToDevice, c->stream), "H2D q");
  CK(cudaMemcpyAsync(dcand, hcand, cb, cudaMemcpyHostToDevice, c->stream), "H2D cand");
  RET(rerank_dev(c->stream, s, dq, 1, dcand, n_cand, k, metric, dids, dsc, dcn));
  CK(cudaMemcpyAsync(hids, dids, ob, cudaMemcpyDeviceToHost, c->stream), "D2H ids");
  CK(cudaMemcpyAsync(hsc, dsc, ob, cudaMemcpyDeviceToHost, c->stream), "D2H scores");
  CK(cudaMemcpyAsync(hcn, dcn, 4, cudaMemcpyDeviceToHost, c->stream), "D2H count");
  CK(cudaStreamSynchronize(c->stream), "sync");
  memcpy(ids_out, hids, ob);
  memcpy(scores_out, hsc, ob);
  if (count_out) *count_out = *hcn;
  return VS_OK;
}

int32_t vs_merge_topk(const int64_t* ids, const double* scores, int64_t total, int32_t k, int64_t* ids_out,
                      double* scores_out, int32_t* count_out) {
  if ((!ids || !scores) && total > 0) return fail(VS_EINVAL, "null pointer");
  if (!ids_out || !scores_out || total < 0) return fail(VS_EINVAL, "null output pointer or negative total");
  if (k <= 0 || k > TOPK_MAX_K) return fail(VS_EINVAL, "k must be in 1..%d", TOPK_MAX_K);
  ThreadCtx* c;
  RET(ctx_bind(&c));
  if (total == 0) {
    fill_empty(ids_out, scores_out, count_out, 1, k);
    return VS_OK;
  }
  const size_t tb = (size_t)total * 8, ob = (size_t)k * 8;
  RET(ctx_reserve_dev(c, Arena::need({tb, tb, ob, ob, 4})));
  RET(ctx_reserve_host(c, Arena::need({ob, ob, 4})));
  Arena A(c->d_buf), H(c->h_buf);
  int64_t* din = A.take<int64_t>(total);
  double* dsin = A.take<double>(total);
  int64_t* dids = A.take<int64_t>(k);
  double* dsc = A.take<double>(k);
  int32_t* dcn = A.take<int32_t>(1);
  int64_t* hids = H.take<int64_t>(k);
  double* hsc = H.take<double>(k);
  int32_t* hcn = H.take<int32_t>(1);
  CK(cudaMemcpyAsync(din, ids, tb, cudaMemcpyHostToDevice, c->stream), "H2D ids");
  CK(cudaMemcpyAsync(dsin, scores, tb, cudaMemcpyHostToDevice, c->stream), "H2D scores");
  CK(launch_merge(din, dsin, total, k, dids, dsc, dcn, c->stream), "merge launch");
  CK(cudaMemcpyAsync(hids, dids, ob, cudaMemcpyDeviceToHost, c->stream), "D2H ids");
  CK(cudaMemcpyAsync(hsc, dsc, ob, cudaMemcpyDeviceToHost, c->stream), "D2H scores");
  CK(cudaMemcpyAsync(hcn, dcn, 4, cudaMemcpyDeviceToHost, c->stream), "D2H count");
  CK(cudaStreamSynchronize(c->stream), "sync");
  memcpy(ids_out, hids, ob);
  memcpy(scores_out, hsc, ob);
  if (count_out) *count_out = *hcn;
  return VS_OK;
}

// ---- BEST_FIRST expansion scoring: approximate distances of listed ids against the resident codes -------------
// The reference builds the LUT once per (query, sealed segment) (J/fdb/FdbVectorIndex.java:741), reads ALL codes of
// the segment into a HashMap (:746-759) and then scores frontier neighbours one at a time (:950-963).  Here the codes
// are already resident: vs_adc_query_begin builds the LUT on the device and keeps it, vs_adc_query_gather scores an
// id list per expansion step, vs_adc_query_end releases the LUT.
namespace {
struct AdcQuery {
  std::shared_ptr<Segment> seg;      // plain handle: the LUT lives on the segment's device
  double* d_lut = nullptr;           // [M][K]
  std::vector<uint64_t> shard_q;     // sharded handle: one context per shard
  uint64_t sharded = 0;
};
std::mutex g_aq_mu;
std::unordered_map<uint64_t, std::shared_ptr<AdcQuery>> g_aq;
uint64_t g_aq_next = 0x4151000000000001ull;  // "AQ"
std::shared_ptr<AdcQuery> aq_lookup(uint64_t qh) {
  std::lock_guard<std::mutex> g(g_aq_mu);
  auto it = g_aq.find(qh);
  return it == g_aq.end() ? nullptr : it->second;
}
}  // namespace

int32_t vs_adc_query_begin(uint64_t h, const float* q, uint64_t* query_out) {
  if (!q || !query_out) return fail(VS_EINVAL, "null pointer");
  auto aq = std::make_shared<AdcQuery>();
  if (group_is_sharded(h)) {
    RET(group_adc_query_begin(h, q, &aq->shard_q));
    aq->sharded = h;
  } else {
    aq->seg = seg_lookup(h);
    Segment* s = aq->seg.get();
    if (!s) return fail(VS_EHANDLE, "unknown segment handle");
    if (s->M == 0) return fail(VS_ESTATE, "segment has no PQ attached");
    ThreadCtx* c;
    RET(ctx_bind_seg(&c, s));
    const size_t qb = (size_t)s->d * 4;
    RET(ctx_reserve_dev(c, Arena::need({qb})));
    RET(ctx_reserve_host(c, qb));
    CK(cudaMalloc(&aq->d_lut, (size_t)s->M * s->K * 8), "cudaMalloc(lut)");
    float* dq = static_cast<float*>(c->d_buf);
    memcpy(c->h_buf, q, qb);
    cudaError_t e = cudaMemcpyAsync(dq, c->h_buf, qb, cudaMemcpyHostToDevice, c->stream);
    if (e == cudaSuccess) e = launch_build_lut(s->centroids, s->M, s->K, s->subDim, dq, 1, lanes(), aq->d_lut, c->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
    if (e != cudaSuccess) {
      cudaFree(aq->d_lut);
      return cuda_fail(e, "build_lut");
    }
  }
  std::lock_guard<std::mutex> g(g_aq_mu);
  *query_out = g_aq_next++;
  g_aq[*query_out] = aq;
  return VS_OK;
}

int32_t vs_adc_query_gather(uint64_t query, const int64_t* ids, int64_t n_ids, double* out, uint8_t* valid_out) {
  std::shared_ptr<AdcQuery> aq = aq_lookup(query);
  if (!aq) return fail(VS_EHANDLE, "unknown ADC query context");
  if (n_ids < 0 || ((!ids || !out) && n_ids > 0)) return fail(VS_EINVAL, "null pointer or negative count");
  if (n_ids == 0) return VS_OK;
  if (aq->sharded) return group_adc_query_gather(aq->sharded, aq->shard_q, ids, n_ids, out, valid_out);
  Segment* s = aq->seg.get();
  ThreadCtx* c;
  RET(ctx_bind_seg(&c, s));
  const size_t ib = (size_t)n_ids * 8, ob = (size_t)n_ids * 8, vb = (size_t)n_ids;
  RET(ctx_reserve_dev(c, Arena::need({ib, ob, vb})));
  RET(ctx_reserve_host(c, Arena::need({ib, ob, vb})));
  Arena A(c->d_buf), H(c->h_buf);
  int64_t* dids = A.take<int64_t>(n_ids);
  double* dout = A.take<double>(n_ids);
  uint8_t* dval = A.take<uint8_t>(n_ids);
  int64_t* hids = H.take<int64_t>(n_ids);
  double* hout = H.take<double>(n_ids);
  uint8_t* hval = H.take<uint8_t>(n_ids);
  memcpy(hids, ids, ib);
  CK(cudaMemcpyAsync(dids, hids, ib, cudaMemcpyHostToDevice, c->stream), "H2D ids");
  // short lists: the kernel writes the pinned staging buffer itself (device-visible under UVA)
  const bool direct = n_ids <= 4096;
  CK(launch_adc_gather(aq->d_lut, s->M, s->K, s->codes, s->n, s->id_base, dids, n_ids, direct ? hout : dout,
                       direct ? hval : dval, c->stream), "adc gather launch");
  if (!direct) {
    CK(cudaMemcpyAsync(hout, dout, ob, cudaMemcpyDeviceToHost, c->stream), "D2H distances");
    CK(cudaMemcpyAsync(hval, dval, vb, cudaMemcpyDeviceToHost, c->stream), "D2H valid");
  }
  CK(cudaStreamSynchronize(c->stream), "sync");
  memcpy(out, hout, ob);
  if (valid_out) memcpy(valid_out, hval, vb);
  return VS_OK;
}

int32_t vs_adc_query_end(uint64_t query) {
  std::shared_ptr<AdcQuery> aq;
  {
    std::lock_guard<std::mutex> g(g_aq_mu);
    auto it = g_aq.find(query);
    if (it == g_aq.end()) return fail(VS_EHANDLE, "unknown ADC query context");
    aq = it->second;
    g_aq.erase(it);
  }
  if (aq->sharded) return group_adc_query_end(aq->shard_q);
  if (aq->d_lut && device_registered(aq->seg->device)) {
    cudaSetDevice(aq->seg->device);
    cudaFree(aq->d_lut);  // (implicitly waits for the device)
  }
  return VS_OK;
}

// one-shot form: LUT, gather, release
int32_t vs_adc_gather(uint64_t h, const float* q, const int64_t* ids, int64_t n_ids, double* out, uint8_t* valid_out) {
  uint64_t qh = 0;
  RET(vs_adc_query_begin(h, q, &qh));
  const int r = vs_adc_query_gather(qh, ids, n_ids, out, valid_out);
  std::string msg = r != VS_OK ? t_err : std::string();
  vs_adc_query_end(qh);
  if (r != VS_OK) t_err = msg;
  return r;
}

// ---- graph construction distances (GraphBuilder.buildL2Neighbors / buildPrunedNeighbors) ------------------------
// The O(n^2) all-pairs part of sealing: every row of the segment queries the segment (the batched tensor-core
// nomination + exact re-score of vs_bruteforce_topk, device to device, rows never leave HBM), then knn_finalize
// re-orders each node's candidates by (l2Squared, j) -- the reference's Comparator.comparingDouble(l2Squared) over a
// stable sort (J/graph/GraphBuilder.java:50, :83-87) -- drops the node itself and, for the pruned variant, applies
// the greedy alpha rule (:92-106).  l_build <= 0: buildL2Neighbors(vectors, degree); else buildPrunedNeighbors(vectors,
// degree, l_build, alpha).  neighbors_out is int[n][degree] (row indices, -1 padded), counts_out[n] the list lengths.
int32_t vs_knn_graph(uint64_t h, int32_t degree, int32_t l_build, double alpha, int32_t* neighbors_out, int32_t* counts_out) {
  if (group_is_sharded(h)) return fail(VS_ESTATE, "graph construction runs on one device: upload the segment with vs_init(device)");
  std::shared_ptr<Segment> s_ref = seg_lookup(h);
  Segment* s = s_ref.get();
  if (!s) return fail(VS_EHANDLE, "unknown segment handle");
  if (!neighbors_out || degree <= 0 || degree > 512) return fail(VS_EINVAL, "need an output and degree in 1..512");
  if (s->skip) return fail(VS_ESTATE, "the segment has a skip mask: GraphBuilder takes every vector of the segment");
  if (s->n > 0x7fffffffLL) return fail(VS_EINVAL, "a Java array cannot hold more than 2^31-1 vectors");
  ThreadCtx* c;
  RET(ctx_bind_seg(&c, s));
  const int64_t n = s->n;
  if (n == 0) return VS_OK;
  const bool prune_mode = l_build > 0;
  const bool prune = prune_mode && alpha > 1.0;                       // :79 boolean prune = alpha > 1.0
  int64_t limit = prune_mode ? std::min<int64_t>(std::max(0, l_build), n - 1) : std::min<int64_t>(degree, n - 1);   // :88 / :51
  const int out_deg = prune_mode ? (int)std::min<int64_t>(degree, limit) : degree;  // :89 selected.length
  if (limit + 1 > TOPK_MAX_K) return fail(VS_EINVAL, "candidate lists of more than %d rows are not supported", TOPK_MAX_K - 1);
  if (n == 1 || limit == 0) {
    for (int64_t i = 0; i < n * degree; i++) neighbors_out[i] = -1;
    if (counts_out) for (int64_t i = 0; i < n; i++) counts_out[i] = 0;
    return VS_OK;
  }
  // the node itself is (usually) its own nearest row: one extra slot for it, a few more to close groups of equal distance
  const int kq = (int)std::min<int64_t>(std::min<int64_t>(limit + 1 + 8, TOPK_MAX_K), n);
  // rows per pass: the per-query partial lists of the scan path (one list per CTA) stay within half a gigabyte
  int64_t chunk64 = (int64_t(1) << 29) / ((int64_t)g_sms * kq * 16);
  chunk64 = std::max<int64_t>(16, std::min<int64_t>(2048, chunk64));
  const int chunk = (int)std::min<int64_t>(n, chunk64);
  const int keep = prune_mode ? out_deg : (int)std::min<int64_t>(degree, limit);
  cudaStream_t st = c->stream;
  struct Tmp {
    void* p = nullptr;
    ~Tmp() { if (p) cudaFree(p); }
  } t_ids, t_sc, t_cn, t_nb, t_ct, t_fl, t_nodes;
  CK(cudaMalloc(&t_ids.p, (size_t)chunk * kq * 8), "cudaMalloc(knn ids)");
  CK(cudaMalloc(&t_sc.p, (size_t)chunk * kq * 8), "cudaMalloc(knn scores)");
  CK(cudaMalloc(&t_cn.p, (size_t)chunk * 4), "cudaMalloc(knn counts)");
  CK(cudaMalloc(&t_nb.p, (size_t)n * degree * 4), "cudaMalloc(neighbors)");
  CK(cudaMalloc(&t_ct.p, (size_t)n * 4), "cudaMalloc(neighbor counts)");
  CK(cudaMalloc(&t_fl.p, (size_t)chunk * 4), "cudaMalloc(flags)");
  CK(cudaMalloc(&t_nodes.p, (size_t)chunk * 4), "cudaMalloc(nodes)");
  int64_t* d_ids = static_cast<int64_t*>(t_ids.p);
  double* d_sc = static_cast<double*>(t_sc.p);
  int32_t* d_cn = static_cast<int32_t*>(t_cn.p);
  int32_t* d_nb = static_cast<int32_t*>(t_nb.p);
  int32_t* d_ct = static_cast<int32_t*>(t_ct.p);
  int32_t* d_fl = static_cast<int32_t*>(t_fl.p);
  int32_t* d_nodes = static_cast<int32_t*>(t_nodes.p);
  std::vector<int32_t> h_fl(chunk), h_nodes;
  for (int64_t r0 = 0; r0 < n; r0 += chunk) {
    const int cnt = (int)std::min<int64_t>(chunk, n - r0);
    // queries = the segment's own rows, already on the device
    RET(vs_bruteforce_topk_dev(h, s->X + (size_t)r0 * s->d, cnt, kq, VS_METRIC_L2, d_ids, d_sc, d_cn, st));
    CK(launch_knn_finalize(s->X, n, s->d, lanes(), r0, cnt, d_ids, d_sc, d_cn, kq, (int)limit, degree, keep, true,
                           alpha, prune, s->id_base, d_nb, d_ct, d_fl, st), "knn finalize launch");
    CK(cudaMemcpyAsync(h_fl.data(), d_fl, (size_t)cnt * 4, cudaMemcpyDeviceToHost, st), "D2H flags");
    CK(cudaStreamSynchronize(st), "sync");
    h_nodes.clear();
    for (int q = 0; q < cnt; q++)
      if (h_fl[q]) h_nodes.push_back((int32_t)(r0 + q));
    // nodes whose list may have cut a run of equal distances short: exact candidates, same finalisation (no closure check)
    for (size_t f0 = 0; f0 < h_nodes.size(); f0 += chunk) {
      const int fc = (int)std::min<size_t>(chunk, h_nodes.size() - f0);
      const int kx = (int)std::min<int64_t>(limit, n - 1);
      CK(cudaMemcpyAsync(d_nodes, h_nodes.data() + f0, (size_t)fc * 4, cudaMemcpyHostToDevice, st), "H2D nodes");
      CK(launch_knn_exact(s->X, n, s->d, lanes(), d_nodes, fc, kx, s->id_base, d_ids, d_sc, d_cn, st), "knn exact launch");
      // the exact lists are per listed node, not per consecutive row: finalise them one node at a time
      for (int q = 0; q < fc; q++)
        CK(launch_knn_finalize(s->X, n, s->d, lanes(), h_nodes[f0 + q], 1, d_ids + (size_t)q * kx, d_sc + (size_t)q * kx, d_cn + q, kx,
                               (int)limit, degree, keep, false, alpha, prune, s->id_base, d_nb, d_ct, nullptr, st),
           "knn finalize launch");
      CK(cudaStreamSynchronize(st), "sync");
    }
  }
  CK(cudaMemcpyAsync(neighbors_out, d_nb, (size_t)n * degree * 4, cudaMemcpyDeviceToHost, st), "D2H neighbors");
  if (counts_out) CK(cudaMemcpyAsync(counts_out, d_ct, (size_t)n * 4, cudaMemcpyDeviceToHost, st), "D2H counts");
  CK(cudaStreamSynchronize(st), "sync");
  return VS_OK;
}

// =================================================================================================
// build operations
// =================================================================================================
// PqTrainer.train over a corpus sharded by ascending row range: this process holds rows
// [row_lo, row_lo + segment rows) of n_total.  See include/vsgpu.h.
int32_t vs_pq_train_sharded(uint64_t h, int64_t n_total, int64_t row_lo, int32_t rank, int32_t world, int32_t exact_order,
                            int32_t M, int32_t K, int32_t iterations, int64_t seed, float* d_comm_f32, int32_t* d_comm_i32,
                            vs_allreduce_fn allreduce, void* user, float* centroids_out) {
  std::shared_ptr<Segment> s_ref = seg_lookup(h);
  Segment* s = s_ref.get();
  if (!s) return fail(VS_EHANDLE, "unknown segment handle");
  const int d = s->d;
  if (M <= 0 || K <= 0) return fail(VS_EINVAL, "Invalid PQ params (m,k,dimension)");
  if (d % M != 0) return fail(VS_EINVAL, "dimension must be divisible by m");
  if (!centroids_out || !d_comm_f32 || !d_comm_i32 || !allreduce) return fail(VS_EINVAL, "null pointer");
  if (n_total <= 0) return fail(VS_EEMPTY, "empty training set (the reference throws IndexOutOfBoundsException)");
  if (n_total > 0x7fffffffLL) return fail(VS_EINVAL, "a Java List cannot hold more than 2^31-1 vectors");
  if (row_lo < 0 || row_lo + s->n > n_total) return fail(VS_EINVAL, "row range outside the corpus");
  if (s->n == 0) return fail(VS_EINVAL, "a rank must own at least one row");
  if (world <= 0 || rank < 0 || rank >= world) return fail(VS_EINVAL, "rank must be in [0, world)");
  ThreadCtx* c;
  RET(ctx_bind_seg(&c, s));
  TrainComm comm{row_lo, n_total, rank, world, exact_order != 0 ? 1 : 0, user, allreduce, d_comm_f32, d_comm_i32};
  return pq_train_device(c->stream, s->X, s->n, d, M, K, iterations, seed, lanes(), centroids_out, &comm);
}

int32_t vs_pq_encode_batch(const float* centroids, int32_t M, int32_t K, int32_t subDim, const float* rows,
                           uint64_t h, int64_t n, uint8_t* codes_out) {
  if (!centroids || (!codes_out && n > 0) || n < 0) return fail(VS_EINVAL, "null pointer or negative n");
  RET(check_pq_shape(M, K, subDim));
  if (!rows && group_is_sharded(h)) return group_pq_encode(h, centroids, M, K, subDim, n, codes_out);
  const int d = M * subDim;
  const float* dX = nullptr;
  std::shared_ptr<Segment> s_ref;
  Segment* s = nullptr;
  if (!rows && n > 0) {
    s_ref = seg_lookup(h);
    s = s_ref.get();
    if (!s) return fail(VS_EHANDLE, "rows == NULL needs a valid segment handle");
    if (s->d != d) return fail(VS_EINVAL, "segment dimension %d != M*subDim %d", s->d, d);
    if (n > s->n) return fail(VS_EINVAL, "n exceeds the segment's row count");
    dX = s->X;
  }
  ThreadCtx* c;
  if (s) RET(ctx_bind_seg(&c, s));
  else RET(ctx_bind(&c));
  if (n == 0) return VS_OK;
  const size_t cb = (size_t)M * K * subDim * 4;
  // host rows are streamed through the scratch in slabs
  const int64_t slab = rows ? std::min<int64_t>(n, (int64_t)((size_t(256) << 20) / ((size_t)d * 4)) + 1) : n;
  const size_t xb = rows ? (size_t)slab * d * 4 : 0, ob = (size_t)slab * M;
  RET(ctx_reserve_dev(c, Arena::need({cb, xb, ob})));
  Arena A(c->d_buf);
  float* dc = A.take<float>((size_t)M * K * subDim);
  float* dx = A.take<float>(rows ? (size_t)slab * d : 0);
  uint8_t* dcodes = A.take<uint8_t>(ob);
  CK(cudaMemcpyAsync(dc, centroids, cb, cudaMemcpyHostToDevice, c->stream), "H2D centroids");
  for (int64_t r0 = 0; r0 < n; r0 += slab) {
    const int64_t cnt = std::min<int64_t>(slab, n - r0);
    const float* src = dX ? dX + (size_t)r0 * d : dx;
    if (rows) CK(cudaMemcpyAsync(dx, rows + (size_t)r0 * d, (size_t)cnt * d * 4, cudaMemcpyHostToDevice, c->stream), "H2D rows");
    PqAssignLaunch L{};
    L.X = src; L.n = cnt; L.d = d; L.M = M; L.K = K; L.subDim = subDim; L.centroids = dc; L.lanes = lanes();
    L.codes_u8 = dcodes; L.assign_i32 = nullptr; L.s_begin = 0; L.s_end = M;
    CK(launch_pq_assign(L, c->stream), "pq assign launch");
    CK(cudaMemcpyAsync(codes_out + (size_t)r0 * M, dcodes, (size_t)cnt * M, cudaMemcpyDeviceToHost, c->stream), "D2H codes");
    CK(cudaStreamSynchronize(c->stream), "sync");
  }
  return VS_OK;
}

int32_t vs_pq_train(const float* rows, uint64_t h, int64_t n, int32_t d, int32_t M, int32_t K,
                    int32_t iterations, int64_t seed, float* centroids_out) {
  // PqTrainer.java:29-34
  if (M <= 0 || K <= 0 || d <= 0) return fail(VS_EINVAL, "Invalid PQ params (m,k,dimension)");
  if (d % M != 0) return fail(VS_EINVAL, "dimension must be divisible by m");
  if (!centroids_out) return fail(VS_EINVAL, "null output pointer");
  if (n <= 0) return fail(VS_EEMPTY, "empty training set (the reference throws IndexOutOfBoundsException)");
  if (n > 0x7fffffffLL) return fail(VS_EINVAL, "a Java List cannot hold more than 2^31-1 vectors");
  std::shared_ptr<Segment> s_ref;
  if (!rows) {
    if (group_is_sharded(h)) return group_pq_train(h, n, d, M, K, iterations, seed, centroids_out);
    s_ref = seg_lookup(h);
    if (!s_ref) return fail(VS_EHANDLE, "rows == NULL needs a valid segment handle");
  }
  ThreadCtx* c;
  if (s_ref) RET(ctx_bind_seg(&c, s_ref.get()));
  else RET(ctx_bind(&c));
  const float* dX = nullptr;
  float* owned = nullptr;
  if (rows) {
    CK(cudaMalloc(&owned, (size_t)n * d * 4), "cudaMalloc(train rows)");
    cudaError_t e = cudaMemcpyAsync(owned, rows, (size_t)n * d * 4, cudaMemcpyHostToDevice, c->stream);
    if (e != cudaSuccess) {
      cudaFree(owned);
      return cuda_fail(e, "H2D rows");
    }
    dX = owned;
  } else {
    Segment* s = s_ref.get();
    if (s->d != d) return fail(VS_EINVAL, "segment dimension %d != d %d", s->d, d);
    if (n > s->n) return fail(VS_EINVAL, "n exceeds the segment's row count");
    dX = s->X;
  }
  int r = pq_train_device(c->stream, dX, n, d, M, K, iterations, seed, lanes(), centroids_out);
  if (owned) {
    cudaStreamSynchronize(c->stream);
    cudaFree(owned);
  }
  return r;
}

// =================================================================================================
// device-pointer (stream) variants
// =================================================================================================
// Scratch for these comes from the calling thread's context and is only valid until that thread's
// next libvsgpu call; the caller orders work on `stream` (the multi-GPU coordinator uses one
// thread and one stream per process).
// the stream variants' batched route: *done = true when the work was enqueued here
static int batch_try_dev(ThreadCtx* c, cudaStream_t st, Segment* s, const float* d_q, int nq, int k, int metric,
                         int64_t* d_ids, double* d_scores, int32_t* d_counts, int64_t out_stride, bool* done) {
  *done = false;
  const bool cosine = metric == VS_METRIC_COSINE;
  const int mode = batch_wanted(s, nq, cosine);
  if (mode == 0) return VS_OK;
  bool ok = false, half = false;
  RET(batch_prepare(st, s, cosine, &ok, &half));
  if (!ok) return VS_OK;
  BatchLaunch bp;
  RET(plan_batch(s, k, cosine, half, &bp));
  if (mode == 2 && !(half && bp.sh_ok)) return VS_OK;
  const size_t sb = batch_scratch_need(bp, nq);
  const int chunk = batch_chunk(bp, nq);
  const bool grow = Arena::need({sb}) > c->d_cap || (size_t)chunk > c->ticket_cap;
  RET(ctx_reserve_dev(c, Arena::need({sb})));
  RET(ctx_reserve_ticket(c, chunk));
  if (grow) CK(cudaStreamSynchronize(c->stream), "sync");  // ticket memset ran on c->stream
  Arena A(c->d_buf);
  RET(batch_run_dev(st, s, bp, cosine, d_q, nq, d_ids, d_scores, d_counts, out_stride, A.take<char>(sb), c->d_ticket, mode == 2));
  *done = true;
  return VS_OK;
}

// Diagnostics for the batched path's nomination bound: runs the tensor-core stage only and returns, per query,
// the group minima of a(q, x) (L2: |x|^2 - 2<q,x>; COSINE: -<q,x>/|x|) and the slack the selection adds.
int32_t vs_debug_batch_groupmins(uint64_t h, const float* q, int32_t nq, int32_t metric, float* gm_out,
                                 int64_t gm_capacity, int64_t* ngroups_out, int32_t* group_out, double* slack_out) {
  std::shared_ptr<Segment> s_ref = seg_lookup(h);
  Segment* s = s_ref.get();
  RET(check_query_args(s, q, nq, 1, metric));
  if (!ngroups_out || !group_out) return fail(VS_EINVAL, "null output pointer");
  ThreadCtx* c;
  RET(ctx_bind_seg(&c, s));
  const bool cosine = metric == VS_METRIC_COSINE;
  if (s->n == 0 || !batch_supported(s->d, lanes(), cosine, s->n)) return fail(VS_ESTATE, "segment shape not eligible for the batched path");
  bool ok = false, half = false;
  RET(batch_prepare(c->stream, s, cosine, &ok, &half));
  if (!ok) return fail(VS_ESTATE, "segment holds non-finite rows: the batched path is off");
  BatchLaunch bp;
  RET(plan_batch(s, 1, cosine, half, &bp));
  *ngroups_out = bp.ngroups;
  *group_out = bp.group;
  if (!gm_out) return VS_OK;
  if (nq > batch_chunk(bp, nq)) return fail(VS_EINVAL, "too many queries for one chunk");
  if (gm_capacity < (int64_t)nq * bp.ngroups) return fail(VS_EINVAL, "gm_out too small");
  const size_t qb = (size_t)nq * s->d * 4, sb = batch_scratch_need(bp, nq);
  RET(ctx_reserve_dev(c, Arena::need({qb, sb})));
  Arena A(c->d_buf);
  float* dq = A.take<float>((size_t)nq * s->d);
  char* scratch = A.take<char>(sb);
  CK(cudaMemcpyAsync(dq, q, qb, cudaMemcpyHostToDevice, c->stream), "H2D q");
  Arena B(scratch);
  BatchLaunch L = bp;
  const int m = cosine ? 1 : 0;
  L.X = s->X; L.skip = s->skip; L.q = dq; L.nq = nq; L.tmX = half ? s->tmXh : s->tmX; L.x_scale = s->x_scale;
  L.tmX128 = half ? s->tmXh_b128 : s->tmX_b128; L.pairs = g_batch_pairs.load() == 1 || (g_batch_pairs.load() == 2 && !bp.pair_stat);
  L.coef = static_cast<const float*>(s->ab[m]); L.stats = static_cast<const SegStats*>(s->stats[m]);
  const int chunk = batch_chunk(bp, nq);
  L.gm = B.take<float>((size_t)((chunk + 255) / 256 * 256) * bp.gm_stride);
  L.fb = B.take<int32_t>(1 + 2 * (size_t)chunk);
  L.partial = B.take<ulonglong2>((size_t)chunk * batch_partial_keys(bp, chunk));
  L.qh = half ? B.take<char>((size_t)chunk * bp.dp * 2) : nullptr;
  L.qinv = half ? B.take<float>((size_t)((chunk + 255) / 256 * 256)) : nullptr;
  L.gemm_only = true;
  CK(launch_batch(L, c->stream), "batched scan launch");
  CK(cudaMemcpy2DAsync(gm_out, (size_t)bp.ngroups * 4, L.gm, (size_t)bp.gm_stride * 4, (size_t)bp.ngroups * 4, nq,
                       cudaMemcpyDeviceToHost, c->stream), "D2H group minima");
  CK(cudaStreamSynchronize(c->stream), "sync");
  if (slack_out) {
    float xmax2;
    memcpy(&xmax2, &s->xmax2, 4);
    for (int i = 0; i < nq; i++) {
      double qq = 0.0;
      for (int j = 0; j < s->d; j++) qq += (double)q[(size_t)i * s->d + j] * q[(size_t)i * s->d + j];
      slack_out[i] = batch_slack_host(cosine, half, s->d, sqrt((double)xmax2), sqrt(qq));
    }
  }
  return VS_OK;
}

int32_t vs_bruteforce_topk_dev(uint64_t h, const float* d_q, int32_t nq, int32_t k, int32_t metric,
                               int64_t* d_ids, double* d_scores, int32_t* d_counts, void* stream) {
  std::shared_ptr<Segment> s_ref = seg_lookup(h);
  Segment* s = s_ref.get();
  RET(check_query_args(s, d_q, nq, k, metric));
  if (!d_ids || !d_scores || !d_counts) return fail(VS_EINVAL, "null output pointer");
  ThreadCtx* c;
  RET(ctx_bind_seg(&c, s));
  RET(ctx_use_stream(c, stream));
  if (s->n == 0) return fail(VS_EINVAL, "empty segment: use the host variant");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  bool done = false;
  RET(batch_try_dev(c, st, s, d_q, nq, k, metric, d_ids, d_scores, d_counts, 0, &done));
  if (done) return VS_OK;
  ScanLaunch p;
  RET(plan_scan(s, nq, k, metric == VS_METRIC_COSINE, &p));
  const size_t pb = (size_t)nq * p.partial_keys * 16;
  if (pb > (size_t(1) << 31)) return fail(VS_EINVAL, "query batch too large for the stream variant");
  const bool grow = Arena::need({pb}) > c->d_cap || (size_t)nq > c->ticket_cap;
  RET(ctx_reserve_dev(c, Arena::need({pb})));
  RET(ctx_reserve_ticket(c, nq));
  // ticket memset (if any) ran on c->stream: order it before the caller's stream
  if (grow) CK(cudaStreamSynchronize(c->stream), "sync");
  Arena A(c->d_buf);
  ulonglong2* dpart = A.take<ulonglong2>((size_t)nq * p.partial_keys);
  return bruteforce_dev(c, st, s, d_q, nq, k, metric, d_ids, d_scores, d_counts, dpart, c->d_ticket, p);
}

int32_t vs_bruteforce_topk_packed_dev(uint64_t h, const float* d_q, int32_t nq, int32_t k, int32_t metric,
                                      int64_t* d_pack, int32_t* d_counts, void* stream) {
  std::shared_ptr<Segment> s_ref = seg_lookup(h);
  Segment* s = s_ref.get();
  RET(check_query_args(s, d_q, nq, k, metric));
  if (!d_pack || !d_counts) return fail(VS_EINVAL, "null output pointer");
  ThreadCtx* c;
  RET(ctx_bind_seg(&c, s));
  RET(ctx_use_stream(c, stream));
  if (s->n == 0) {  // an empty shard contributes an all-empty list to the exchange (same collective path as its peers)
    CK(launch_fill_pack(d_pack, nq, k, 0, d_counts, static_cast<cudaStream_t>(stream)), "fill launch");
    return VS_OK;
  }
  bool done = false;
  RET(batch_try_dev(c, static_cast<cudaStream_t>(stream), s, d_q, nq, k, metric, d_pack, reinterpret_cast<double*>(d_pack + k),
                    d_counts, 2 * (int64_t)k, &done));
  if (done) return VS_OK;
  ScanLaunch p;
  RET(plan_scan(s, nq, k, metric == VS_METRIC_COSINE, &p));
  const size_t pb = (size_t)nq * p.partial_keys * 16;
  if (pb > (size_t(1) << 31)) return fail(VS_EINVAL, "query batch too large for the stream variant");
  const bool grow = Arena::need({pb}) > c->d_cap || (size_t)nq > c->ticket_cap;
  RET(ctx_reserve_dev(c, Arena::need({pb})));
  RET(ctx_reserve_ticket(c, nq));
  if (grow) CK(cudaStreamSynchronize(c->stream), "sync");
  Arena A(c->d_buf);
  ulonglong2* dpart = A.take<ulonglong2>((size_t)nq * p.partial_keys);
  return bruteforce_dev(c, static_cast<cudaStream_t>(stream), s, d_q, nq, k, metric, d_pack,
                        reinterpret_cast<double*>(d_pack + k), d_counts, dpart, c->d_ticket, p, 2 * (int64_t)k);
}

int32_t vs_merge_packed_dev(const int64_t* d_gath, int32_t world, int32_t nq, int32_t k, int32_t descending,
                            int64_t* d_ids_out, double* d_scores_out, int32_t* d_counts_out, void* stream) {
  if (!d_gath || !d_ids_out || !d_scores_out || !d_counts_out) return fail(VS_EINVAL, "null pointer");
  if (world <= 0 || nq <= 0) return fail(VS_EINVAL, "world and nq must be positive");
  if (k <= 0 || k > TOPK_MAX_K) return fail(VS_EINVAL, "k must be in 1..%d", TOPK_MAX_K);
  ThreadCtx* c;
  RET(ctx_bind(&c));
  CK(launch_merge_packed(d_gath, world, nq, k, descending != 0, d_ids_out, d_scores_out, d_counts_out,
                         static_cast<cudaStream_t>(stream)), "merge launch");
  return VS_OK;
}

// ---- peer exchange: all-gather + merge over NVLink peer memory (rank.cu) ------------------------------------
namespace {
struct PeerComm {
  int rank = 0, world = 1, depth = 0, device = 0;
  size_t slot_bytes = 0, flags_off = 0, ticket_off = 0, total = 0;
  unsigned char* bases[VS_PEER_MAX_WORLD] = {};
  bool connected = false, by_ptr = false;
  std::mutex mu;
  std::vector<cudaStream_t> streams;  // ring r belongs to streams[r]
  std::vector<uint64_t> ring_seq;
  // The host-buffer entry points (vs_*_exchange) run on a stream the COMMUNICATOR owns, serialised by xmu: whichever
  // request thread makes the call, every rank's host-path exchanges use the same ring in the same order (a ring keyed
  // on the calling thread's stream would burn a ring per thread and could differ between ranks).
  std::mutex xmu;
  cudaStream_t xstream = nullptr;
  unsigned char *x_dev = nullptr, *x_host = nullptr;
  size_t x_dev_cap = 0, x_host_cap = 0;
  std::vector<unsigned char*> x_host_parked;
  cudaStream_t pstream = nullptr;            // flag polling by the host (ranks sharing a device)
  unsigned long long* h_flags = nullptr;     // pinned [VS_PEER_MAX_WORLD]
  bool shared_device = false;  // several ranks on one GPU: the fused publish-then-wait kernel could stall (rank.cu)
  struct RingPack {
    cudaStream_t stream;
    int64_t* p;
    size_t cap;
  };
  std::vector<RingPack> ring_packs;  // send buffers of the one-call stream entry points
};
// communicators live in a registry: a stale or foreign handle is an error, never a wild pointer
std::mutex g_peer_mu;
std::unordered_map<uint64_t, PeerComm*> g_peers;
uint64_t g_peer_next_id = 0x7065657200000001ull;  // "peer" + counter
PeerComm* peer_lookup(uint64_t comm) {
  std::lock_guard<std::mutex> g(g_peer_mu);
  auto it = g_peers.find(comm);
  return it == g_peers.end() ? nullptr : it->second;
}
struct PeerSlot {
  size_t data_off, flag_off;
  const int64_t* gath;
  const unsigned long long* flags;
  unsigned int* ticket;
  uint64_t seq;
  int ring = -1;
};
// Slots are handed out per STREAM: each stream a communicator sees gets its own ring of PEER_RING slots, in order
// of first use (every rank issues the same exchanges on corresponding streams in the same order, as with any
// collective).  On one stream a rank publishes exchange t only after its own merge of t-1, which needed everybody's
// publish of t-1 and hence everybody's merge of t-2: the slot of t-PEER_RING is free on every rank, however far
// other streams have run ahead.
constexpr int PEER_RING = 4;
bool peer_next(PeerComm* pc, size_t payload, cudaStream_t st, PeerSlot* out) {
  std::lock_guard<std::mutex> g(pc->mu);
  int ring = -1;
  for (size_t i = 0; i < pc->streams.size(); i++)
    if (pc->streams[i] == st) ring = (int)i;
  if (ring < 0) {
    // a ring whose stream was released (vs_peer_release_stream) is free for the next newcomer
    for (size_t i = 0; i < pc->streams.size() && ring < 0; i++)
      if (pc->streams[i] == nullptr && i != 0) ring = (int)i;
    if (ring >= 0) {
      pc->streams[ring] = st;
    } else {
      if ((int)(pc->streams.size() + 1) * PEER_RING > pc->depth) return false;
      ring = (int)pc->streams.size();
      pc->streams.push_back(st);
      pc->ring_seq.push_back(0);
    }
  }
  PeerSlot& ps = *out;
  ps.seq = ++pc->ring_seq[ring];
  ps.ring = ring;
  const size_t slot = (size_t)ring * PEER_RING + (size_t)((ps.seq - 1) % PEER_RING);
  unsigned char* own = pc->bases[pc->rank];
  const size_t slot_base = slot * (size_t)pc->world * pc->slot_bytes;
  ps.data_off = slot_base + (size_t)pc->rank * payload;
  ps.flag_off = pc->flags_off + slot * VS_PEER_MAX_WORLD * 8;
  ps.gath = reinterpret_cast<const int64_t*>(own + slot_base);
  ps.flags = reinterpret_cast<const unsigned long long*>(own + ps.flag_off);
  ps.ticket = reinterpret_cast<unsigned int*>(own + pc->ticket_off) + slot;
  return true;
}
// Ranks that share a GPU (several communicators on one device: the one-GPU shape of vs_init_multi and of the tests)
// cannot wait ON THE DEVICE at all.  A CTA polling inside a kernel holds its SM's resources, and a persistent
// one-CTA-per-SM kernel of the very peer it waits for (the scan, the nomination GEMM: all of an SM's shared memory)
// may then never become fully resident.  A stream memory operation (cuStreamWaitValue64) occupies no SM but blocks
// its hardware channel, which the peer's stream may be multiplexed onto.  Either way: deadlock until a time-out.
// So for such communicators the HOST waits: it drains its own stream (its publish is out), polls the arrival flags
// with small copies on a stream of their own, and only then launches the consuming kernel, which has nothing left to
// poll.  The exchange is then blocking for the calling thread -- a property of this test shape only; ranks on
// separate GPUs keep the fully asynchronous in-kernel wait.  src < 0: the flags of every rank.
int peer_host_wait(PeerComm* pc, cudaStream_t st, const unsigned long long* flags, unsigned long long seq, int src) {
  CK(cudaStreamSynchronize(st), "sync (exchange, shared device)");
  if (!pc->pstream) {
    CK(cudaStreamCreateWithFlags(&pc->pstream, cudaStreamNonBlocking), "cudaStreamCreate (flag polling)");
    CK(cudaMallocHost(reinterpret_cast<void**>(&pc->h_flags), VS_PEER_MAX_WORLD * 8), "cudaMallocHost (flag polling)");
  }
  const auto t0 = std::chrono::steady_clock::now();
  for (int spins = 0;; spins++) {
    CK(cudaMemcpyAsync(pc->h_flags, flags, (size_t)pc->world * 8, cudaMemcpyDeviceToHost, pc->pstream), "D2H flags");
    CK(cudaStreamSynchronize(pc->pstream), "sync (flag polling)");
    bool all = true;
    for (int r = 0; r < pc->world; r++)
      if ((src < 0 || r == src) && pc->h_flags[r] < seq) all = false;
    if (all) return VS_OK;
    if (std::chrono::steady_clock::now() - t0 > std::chrono::seconds(20))
      return fail(VS_ESTATE, "peer exchange %llu: a rank of this communicator did not arrive within 20 s", seq);
    if (spins > 50) std::this_thread::sleep_for(std::chrono::microseconds(50));
  }
}
// a launch failed after peer_next on this rank: hand the sequence number back so the ring stays in step with the
// peers (they time out on this exchange, but the communicator is not left off by one for good)
void peer_rollback(PeerComm* pc, const PeerSlot& ps) {
  std::lock_guard<std::mutex> g(pc->mu);
  if (ps.ring >= 0 && ps.ring < (int)pc->ring_seq.size() && pc->ring_seq[ps.ring] == ps.seq) pc->ring_seq[ps.ring]--;
}
// the communicator's own stream (host entry points); caller holds pc->xmu
int peer_xstream(PeerComm* pc, cudaStream_t* out) {
  if (!pc->xstream) {
    cudaError_t e = cudaStreamCreateWithFlags(&pc->xstream, cudaStreamNonBlocking);
    if (e != cudaSuccess) return cuda_fail(e, "cudaStreamCreate (communicator stream)");
    std::lock_guard<std::mutex> g(pc->mu);
    pc->streams[0] = pc->xstream;  // ring 0 on every rank, whatever the order in which other streams show up
  }
  *out = pc->xstream;
  return VS_OK;
}
}  // namespace

int32_t vs_peer_create(int32_t rank, int32_t world, int64_t slot_bytes, int32_t depth, uint64_t* comm_out,
                       uint8_t* handle_out) {
  if (!comm_out || !handle_out) return fail(VS_EINVAL, "null pointer");
  if (world < 1 || world > VS_PEER_MAX_WORLD || rank < 0 || rank >= world)
    return fail(VS_EINVAL, "rank / world out of range (world <= %d)", VS_PEER_MAX_WORLD);
  if (slot_bytes < 16 || (slot_bytes & 15) != 0 || depth < PEER_RING || depth > 64 || depth % PEER_RING != 0)
    return fail(VS_EINVAL, "slot_bytes must be a positive multiple of 16 and depth a multiple of %d in %d..64", PEER_RING,
                PEER_RING);
  static_assert(sizeof(cudaIpcMemHandle_t) == VS_PEER_HANDLE_BYTES, "handle size");
  ThreadCtx* c;
  RET(ctx_bind(&c));
  PeerComm* pc = new PeerComm();
  pc->rank = rank;
  pc->world = world;
  pc->depth = depth;
  pc->slot_bytes = (size_t)slot_bytes;
  pc->flags_off = (size_t)depth * world * pc->slot_bytes;
  pc->ticket_off = pc->flags_off + (size_t)depth * VS_PEER_MAX_WORLD * 8;
  pc->total = pc->ticket_off + (size_t)depth * 4 + 256;
  cudaGetDevice(&pc->device);
  // ring 0 belongs to the communicator's own stream (the host entry points)
  pc->streams.push_back(nullptr);
  pc->ring_seq.push_back(0);
  cudaError_t e = cudaMalloc(&pc->bases[rank], pc->total);  // plain cudaMalloc: cudaIpc cannot export pool memory
  if (e == cudaSuccess) e = cudaMemset(pc->bases[rank], 0, pc->total);
  if (e == cudaSuccess) e = cudaDeviceSynchronize();
  cudaIpcMemHandle_t hd;
  if (e == cudaSuccess) e = cudaIpcGetMemHandle(&hd, pc->bases[rank]);
  if (e != cudaSuccess) {
    if (pc->bases[rank]) cudaFree(pc->bases[rank]);
    delete pc;
    return cuda_fail(e, "peer buffer");
  }
  memcpy(handle_out, &hd, sizeof(hd));
  {
    std::lock_guard<std::mutex> g(g_peer_mu);
    *comm_out = g_peer_next_id++;
    g_peers[*comm_out] = pc;
  }
  return VS_OK;
}

int32_t vs_peer_connect(uint64_t comm, const uint8_t* handles) {
  PeerComm* pc = peer_lookup(comm);
  if (!pc || !handles) return fail(VS_EINVAL, "bad peer communicator");
  if (pc->connected) return fail(VS_ESTATE, "peer communicator is already connected");
  ThreadCtx* c;
  RET(ctx_bind_dev(&c, pc->device));
  for (int p = 0; p < pc->world; p++) {
    if (p == pc->rank) continue;
    cudaIpcMemHandle_t hd;
    memcpy(&hd, handles + (size_t)p * VS_PEER_HANDLE_BYTES, sizeof(hd));
    void* ptr = nullptr;
    cudaError_t e = cudaIpcOpenMemHandle(&ptr, hd, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) return cuda_fail(e, "cudaIpcOpenMemHandle (peer buffer; needs P2P access between the GPUs)");
    pc->bases[p] = static_cast<unsigned char*>(ptr);
  }
  pc->connected = true;
  return VS_OK;
}

// One process driving several GPUs (or several communicators on one GPU, as the tests do): no IPC needed, the
// caller passes the base addresses (vs_peer_base) of all ranks' buffers.  With different devices the caller has
// enabled peer access between them (cudaDeviceEnablePeerAccess; vs_init_multi does).
int32_t vs_peer_base(uint64_t comm, uint64_t* base_out) {
  PeerComm* pc = peer_lookup(comm);
  if (!pc || !base_out) return fail(VS_EINVAL, "bad peer communicator");
  *base_out = static_cast<uint64_t>(reinterpret_cast<uintptr_t>(pc->bases[pc->rank]));
  return VS_OK;
}

int32_t vs_peer_connect_ptrs(uint64_t comm, const uint64_t* bases) {
  PeerComm* pc = peer_lookup(comm);
  if (!pc || !bases) return fail(VS_EINVAL, "bad peer communicator");
  if (pc->connected) return fail(VS_ESTATE, "peer communicator is already connected");
  int devs[VS_PEER_MAX_WORLD];
  for (int p = 0; p < pc->world; p++) {
    if (p != pc->rank) {
      if (!bases[p]) return fail(VS_EINVAL, "null base address for rank %d", p);
      pc->bases[p] = reinterpret_cast<unsigned char*>(static_cast<uintptr_t>(bases[p]));
    }
    cudaPointerAttributes at{};
    devs[p] = -1 - p;
    if (cudaPointerGetAttributes(&at, pc->bases[p]) == cudaSuccess && at.type == cudaMemoryTypeDevice) devs[p] = at.device;
    else cudaGetLastError();
    // two ranks on one GPU: the fused one-query kernel (publish, then spin on the peers) needs every rank's kernel
    // resident at once, which one device does not guarantee -> separate publish kernel for this communicator
    for (int o = 0; o < p; o++)
      if (devs[o] == devs[p]) pc->shared_device = true;
  }
  pc->connected = true;
  pc->by_ptr = true;
  return VS_OK;
}

// A caller stream that will not be used with this communicator again gives its ring back (every rank releases the
// corresponding stream at the same point of its exchange sequence, as with any collective).
int32_t vs_peer_release_stream(uint64_t comm, void* stream) {
  PeerComm* pc = peer_lookup(comm);
  if (!pc) return fail(VS_EINVAL, "bad peer communicator");
  std::lock_guard<std::mutex> g(pc->mu);
  for (size_t i = 1; i < pc->streams.size(); i++)
    if (pc->streams[i] == static_cast<cudaStream_t>(stream)) {
      pc->streams[i] = nullptr;  // ring_seq stays: the slot numbering continues where every rank left it
      return VS_OK;
    }
  return fail(VS_EINVAL, "the communicator has no ring for this stream");
}

int32_t vs_peer_destroy(uint64_t comm) {
  PeerComm* pc = nullptr;
  {
    std::lock_guard<std::mutex> g(g_peer_mu);
    auto it = g_peers.find(comm);
    if (it != g_peers.end()) {
      pc = it->second;
      g_peers.erase(it);
    }
  }
  if (!pc) return fail(VS_EINVAL, "bad peer communicator");
  cudaSetDevice(pc->device);
  cudaDeviceSynchronize();
  for (int p = 0; p < pc->world; p++) {
    if (!pc->bases[p]) continue;
    if (p == pc->rank) cudaFree(pc->bases[p]);
    else if (!pc->by_ptr) cudaIpcCloseMemHandle(pc->bases[p]);
  }
  if (pc->x_dev) cudaFree(pc->x_dev);
  if (pc->x_host) cudaFreeHost(pc->x_host);
  for (unsigned char* hp : pc->x_host_parked) cudaFreeHost(hp);
  if (pc->xstream) {
    ctx_forget_stream(pc->device, pc->xstream);
    cudaStreamDestroy(pc->xstream);
  }
  if (pc->pstream) {
    ctx_forget_stream(pc->device, pc->pstream);
    cudaStreamDestroy(pc->pstream);
  }
  if (pc->h_flags) cudaFreeHost(pc->h_flags);
  for (auto& rp : pc->ring_packs) cudaFree(rp.p);
  delete pc;
  return VS_OK;
}

// kind 0: [nq][2k] lists merged by score (descending != 0) or distance; kind 1: ADC + re-rank packs [nq][4][nc] -> k.
// Everything that can fail for one rank alone is checked BEFORE the slot is taken; a launch failure hands it back.
static int exchange_merge(PeerComm* pc, cudaStream_t st, int kind, const int64_t* d_pack, int nq, int nc, int k, int descending,
                          int64_t* d_ids_out, double* d_scores_out, int32_t* d_counts_out) {
  if (!pc || !pc->connected) return fail(VS_ESTATE, "peer communicator is not connected");
  if (!d_pack || !d_ids_out || !d_scores_out || !d_counts_out) return fail(VS_EINVAL, "null pointer");
  if (nq <= 0) return fail(VS_EINVAL, "nq must be positive");
  if (k <= 0 || k > TOPK_MAX_K) return fail(VS_EINVAL, "k must be in 1..%d", TOPK_MAX_K);
  if (kind == 1) {
    if (nc <= 0 || k > nc) return fail(VS_EINVAL, "need 0 < k <= n_cand");
    if ((int64_t)pc->world * nc > 8192) return fail(VS_EINVAL, "world * n_cand must be <= 8192");
  }
  const size_t payload = kind == 0 ? (size_t)nq * 2 * k * 8 : (size_t)nq * 4 * nc * 8;
  if (payload > pc->slot_bytes) return fail(VS_EINVAL, "packed lists (%zu bytes) exceed the communicator's slot", payload);
  PeerSlot ps;
  if (!peer_next(pc, payload, st, &ps))
    return fail(VS_ESTATE, "communicator of depth %d serves %d streams; this is one more", pc->depth, pc->depth / PEER_RING);
  const PeerPublish pub{pc->bases, d_pack, payload, ps.data_off, ps.flag_off, pc->rank};
  const bool on_stream = pc->shared_device && g_peer_spin_shared.load() == 0;  // the host waits, not a kernel (peer_host_wait)
  const bool fused = nq == 1 && g_peer_fused.load() != 0 && !on_stream;  // one query: the merge kernel publishes, then waits
  cudaError_t e = cudaSuccess;
  if (!fused) e = launch_peer_publish(pc->bases, pc->world, pc->rank, d_pack, payload, ps.data_off, ps.flag_off, ps.seq, ps.ticket, st);
  const unsigned long long* poll = ps.flags;
  if (e == cudaSuccess && on_stream) {  // the host waits for the flags; the merge kernel has nothing to poll
    const int wr = peer_host_wait(pc, st, ps.flags, ps.seq, -1);
    if (wr != VS_OK) return wr;
    poll = nullptr;
  }
  if (e == cudaSuccess) {
    e = kind == 0 ? launch_merge_packed(ps.gath, pc->world, nq, k, descending != 0, d_ids_out, d_scores_out, d_counts_out, st,
                                        poll, ps.seq, fused ? &pub : nullptr)
                  : launch_merge_adc_rerank(ps.gath, pc->world, nq, nc, k, d_ids_out, d_scores_out, d_counts_out, st, poll,
                                            ps.seq, fused ? &pub : nullptr);
    // (a separate publish that went out has used the slot: only the fused form can still hand it back)
    if (e != cudaSuccess && fused) peer_rollback(pc, ps);
  } else {
    peer_rollback(pc, ps);
  }
  if (e != cudaSuccess) return cuda_fail(e, "exchange launch");
  return VS_OK;
}

int32_t vs_exchange_merge_packed_dev(uint64_t comm, const int64_t* d_pack, int32_t nq, int32_t k, int32_t descending,
                                     int64_t* d_ids_out, double* d_scores_out, int32_t* d_counts_out, void* stream) {
  PeerComm* pc = peer_lookup(comm);
  if (!pc) return fail(VS_ESTATE, "peer communicator is not connected");
  ThreadCtx* c;
  RET(ctx_bind_dev(&c, pc->device));
  return exchange_merge(pc, static_cast<cudaStream_t>(stream), 0, d_pack, nq, 0, k, descending, d_ids_out, d_scores_out, d_counts_out);
}

int32_t vs_exchange_merge_adc_rerank_packed_dev(uint64_t comm, const int64_t* d_pack, int32_t nq, int32_t n_cand, int32_t k,
                                                int64_t* d_ids_out, double* d_scores_out, int32_t* d_counts_out,
                                                void* stream) {
  PeerComm* pc = peer_lookup(comm);
  if (!pc) return fail(VS_ESTATE, "peer communicator is not connected");
  ThreadCtx* c;
  RET(ctx_bind_dev(&c, pc->device));
  return exchange_merge(pc, static_cast<cudaStream_t>(stream), 1, d_pack, nq, n_cand, k, 1, d_ids_out, d_scores_out, d_counts_out);
}

// staging of the host-buffer exchange calls (caller holds pc->xmu)
static int peer_staging(PeerComm* pc, cudaStream_t st, size_t need_d, size_t need_h) {
  if (need_d > pc->x_dev_cap) {  // (no device-wide synchronisation here: see scratch_alloc)
    CK(cudaStreamSynchronize(st), "sync");
    scratch_free(pc->x_dev, st);
    pc->x_dev = nullptr;
    pc->x_dev_cap = 0;
    void* xp = nullptr;
    CK(scratch_alloc(&xp, need_d + need_d / 2, st), "cudaMallocAsync(exchange staging)");
    pc->x_dev = static_cast<unsigned char*>(xp);
    pc->x_dev_cap = need_d + need_d / 2;
  }
  if (need_h > pc->x_host_cap) {
    CK(cudaStreamSynchronize(st), "sync");
    if (pc->x_host) pc->x_host_parked.push_back(pc->x_host);
    pc->x_host = nullptr;
    pc->x_host_cap = 0;
    CK(cudaMallocHost(&pc->x_host, need_h), "cudaMallocHost(exchange staging)");
    pc->x_host_cap = need_h;
  }
  return VS_OK;
}

// per-stream packed send buffers of the one-call stream entry points (vs_*_exchange_dev): one per ring, since on one
// stream the publish of exchange t has read the buffer before the scan of exchange t+1 writes it
static int peer_ring_pack(PeerComm* pc, cudaStream_t st, size_t bytes, int64_t** pack_out) {
  std::lock_guard<std::mutex> g(pc->mu);
  for (auto& rp : pc->ring_packs)
    if (rp.stream == st && rp.cap >= bytes) {
      *pack_out = rp.p;
      return VS_OK;
    }
  PeerComm::RingPack rp{};
  rp.stream = st;
  rp.cap = bytes < (size_t(64) << 10) ? (size_t(64) << 10) : bytes;
  CK(cudaMalloc(&rp.p, rp.cap), "cudaMalloc(exchange send buffer)");
  pc->ring_packs.push_back(rp);  // an outgrown buffer of the same stream stays until the communicator goes (work may be in flight)
  *pack_out = rp.p;
  return VS_OK;
}

// The sharded query as ONE stream call: local scan into this stream's packed send buffer, peer exchange, merge.  Outputs
// are device pointers (or pinned host memory, which the merge kernel writes directly); nothing synchronises.
int32_t vs_bruteforce_topk_exchange_dev(uint64_t h, uint64_t comm, const float* d_q, int32_t nq, int32_t k, int32_t metric,
                                        int64_t* d_ids, double* d_scores, int32_t* d_counts, void* stream) {
  PeerComm* pc = peer_lookup(comm);
  if (!pc || !pc->connected) return fail(VS_ESTATE, "peer communicator is not connected");
  if (k <= 0 || k > TOPK_MAX_K || nq <= 0) return fail(VS_EINVAL, "nq must be positive and k in 1..%d", TOPK_MAX_K);
  const size_t payload = (size_t)nq * 2 * k * 8;
  if (payload > pc->slot_bytes) return fail(VS_EINVAL, "packed lists (%zu bytes) exceed the communicator's slot", payload);
  ThreadCtx* c;
  RET(ctx_bind_dev(&c, pc->device));
  int64_t* d_pack = nullptr;
  RET(peer_ring_pack(pc, static_cast<cudaStream_t>(stream), payload + (size_t)nq * 4 + 256, &d_pack));
  int32_t* d_cn_local = reinterpret_cast<int32_t*>(reinterpret_cast<char*>(d_pack) + ((payload + 255) & ~size_t(255)));
  RET(vs_bruteforce_topk_packed_dev(h, d_q, nq, k, metric, d_pack, d_cn_local, stream));
  return exchange_merge(pc, static_cast<cudaStream_t>(stream), 0, d_pack, nq, 0, k, 1, d_ids, d_scores, d_counts);
}

int32_t vs_adc_rerank_topk_exchange_dev(uint64_t h, uint64_t comm, const float* d_q, int32_t nq, int32_t n_cand, int32_t k,
                                        int32_t metric, int32_t normalize_on_read, int64_t* d_ids, double* d_scores,
                                        int32_t* d_counts, void* stream) {
  PeerComm* pc = peer_lookup(comm);
  if (!pc || !pc->connected) return fail(VS_ESTATE, "peer communicator is not connected");
  if (nq <= 0 || n_cand <= 0 || k <= 0 || k > n_cand || k > TOPK_MAX_K) return fail(VS_EINVAL, "need nq > 0 and 0 < k <= n_cand");
  if ((int64_t)pc->world * n_cand > 8192) return fail(VS_EINVAL, "world * n_cand must be <= 8192");
  const size_t payload = (size_t)nq * 4 * n_cand * 8;
  if (payload > pc->slot_bytes) return fail(VS_EINVAL, "packed candidates (%zu bytes) exceed the communicator's slot", payload);
  ThreadCtx* c;
  RET(ctx_bind_dev(&c, pc->device));
  int64_t* d_pack = nullptr;
  RET(peer_ring_pack(pc, static_cast<cudaStream_t>(stream), payload, &d_pack));
  RET(vs_adc_rerank_packed_dev(h, d_q, nq, n_cand, metric, normalize_on_read, d_pack, stream));
  return exchange_merge(pc, static_cast<cudaStream_t>(stream), 1, d_pack, nq, n_cand, k, 1, d_ids, d_scores, d_counts);
}

// The host-buffer forms (what a rank's request thread calls): pinned staging in, the stream call above on the
// communicator's own stream, results back (short lists are written by the merge kernel straight into pinned host
// memory), one synchronisation.  Collective: every rank calls with the same queries.
//   kind 0: brute force; 1: ADC + re-rank; 2: ADC lists (ascending approximate distance); 3: re-rank of caller candidates
static int exchange_host(int kind, uint64_t h, uint64_t comm, const float* q, int nq, const int64_t* cand, int n_cand, int k,
                         int metric, int nor, int64_t* ids_out, double* scores_out, int32_t* counts_out) {
  std::shared_ptr<Segment> s_ref = seg_lookup(h);
  Segment* s = s_ref.get();
  const int kout = kind == 2 ? n_cand : k;
  RET(check_query_args(s, q, nq, kout, kind == 2 ? VS_METRIC_L2 : metric));
  if (!ids_out || !scores_out) return fail(VS_EINVAL, "null output pointer");
  if (kind != 0 && (n_cand <= 0 || n_cand > 8192)) return fail(VS_EINVAL, "n_cand must be in 1..8192");
  if ((kind == 1 || kind == 3) && k > n_cand) return fail(VS_EINVAL, "k must be in 1..n_cand");
  if (kind == 2 && n_cand > TOPK_MAX_K) return fail(VS_EINVAL, "n_cand must be in 1..%d", TOPK_MAX_K);
  if (kind == 3 && (!cand || nq != 1)) return fail(VS_EINVAL, "re-rank takes one query and a candidate list");
  if ((kind == 1 || kind == 2) && s->M == 0 && s->n > 0) return fail(VS_ESTATE, "segment has no PQ attached");
  PeerComm* pc = peer_lookup(comm);
  if (!pc || !pc->connected) return fail(VS_ESTATE, "peer communicator is not connected");
  if (pc->device != s->device) return fail(VS_EINVAL, "segment and communicator live on different devices");
  const size_t payload = (kind == 0 || kind == 2) ? (size_t)nq * 2 * kout * 8 : (size_t)nq * 4 * n_cand * 8;
  if (payload > pc->slot_bytes) return fail(VS_EINVAL, "packed lists exceed the communicator's slot");
  if ((kind == 1 || kind == 3) && (int64_t)pc->world * n_cand > 8192) return fail(VS_EINVAL, "world * n_cand must be <= 8192");
  ThreadCtx* c;
  RET(ctx_bind_seg(&c, s));
  auto up = [](size_t b) { return (b + 255) & ~size_t(255); };
  const size_t qb = up((size_t)nq * s->d * 4), pb = up(payload), ib = up((size_t)nq * kout * 8), cb = up((size_t)nq * 4);
  const size_t cdb = kind == 3 ? up((size_t)n_cand * 8) : 0;
  std::lock_guard<std::mutex> g(pc->xmu);
  cudaStream_t st;
  RET(peer_xstream(pc, &st));
  RET(peer_staging(pc, st, qb + pb + cb + 2 * ib + cb + cdb, qb + 2 * ib + cb + cdb));
  unsigned char* d = pc->x_dev;
  float* dq = reinterpret_cast<float*>(d);
  int64_t* dpack = reinterpret_cast<int64_t*>(d + qb);
  int32_t* dcn_local = reinterpret_cast<int32_t*>(d + qb + pb);
  int64_t* dids = reinterpret_cast<int64_t*>(d + qb + pb + cb);
  double* dsc = reinterpret_cast<double*>(d + qb + pb + cb + ib);
  int32_t* dcn = reinterpret_cast<int32_t*>(d + qb + pb + cb + 2 * ib);
  int64_t* dcand = reinterpret_cast<int64_t*>(d + qb + pb + cb + 2 * ib + cb);
  unsigned char* hh = pc->x_host;
  float* hq = reinterpret_cast<float*>(hh);
  int64_t* hids = reinterpret_cast<int64_t*>(hh + qb);
  double* hsc = reinterpret_cast<double*>(hh + qb + ib);
  int32_t* hcn = reinterpret_cast<int32_t*>(hh + qb + 2 * ib);
  int64_t* hcand = reinterpret_cast<int64_t*>(hh + qb + 2 * ib + cb);
  memcpy(hq, q, (size_t)nq * s->d * 4);
  CK(cudaMemcpyAsync(dq, hq, (size_t)nq * s->d * 4, cudaMemcpyHostToDevice, st), "H2D q");
  if (kind == 3) {
    memcpy(hcand, cand, (size_t)n_cand * 8);
    CK(cudaMemcpyAsync(dcand, hcand, (size_t)n_cand * 8, cudaMemcpyHostToDevice, st), "H2D candidates");
  }
  if (kind == 0) RET(vs_bruteforce_topk_packed_dev(h, dq, nq, k, metric, dpack, dcn_local, st));
  else if (kind == 1) RET(vs_adc_rerank_packed_dev(h, dq, nq, n_cand, metric, nor, dpack, st));
  else if (kind == 2) RET(vs_adc_topk_packed_dev(h, dq, nq, n_cand, dpack, dcn_local, st));
  else RET(vs_rerank_packed_dev(h, dq, dcand, n_cand, metric, nor, dpack, st));
  const bool direct = (size_t)nq * kout <= 4096;  // short lists: the merge writes pinned host memory (UVA) itself
  RET(exchange_merge(pc, st, (kind == 0 || kind == 2) ? 0 : 1, dpack, nq, n_cand, kout, kind == 2 ? 0 : 1, direct ? hids : dids,
                     direct ? hsc : dsc, direct ? hcn : dcn));
  if (!direct) {
    CK(cudaMemcpyAsync(hids, dids, (size_t)nq * kout * 8, cudaMemcpyDeviceToHost, st), "D2H ids");
    CK(cudaMemcpyAsync(hsc, dsc, (size_t)nq * kout * 8, cudaMemcpyDeviceToHost, st), "D2H scores");
    CK(cudaMemcpyAsync(hcn, dcn, (size_t)nq * 4, cudaMemcpyDeviceToHost, st), "D2H counts");
  }
  CK(cudaStreamSynchronize(st), "sync");
  memcpy(ids_out, hids, (size_t)nq * kout * 8);
  memcpy(scores_out, hsc, (size_t)nq * kout * 8);
  if (counts_out) memcpy(counts_out, hcn, (size_t)nq * 4);
  return VS_OK;
}

int32_t vs_bruteforce_topk_exchange(uint64_t h, uint64_t comm, const float* q, int32_t nq, int32_t k, int32_t metric,
                                    int64_t* ids_out, double* scores_out, int32_t* counts_out) {
  return exchange_host(0, h, comm, q, nq, nullptr, 0, k, metric, 0, ids_out, scores_out, counts_out);
}
// Sealed segments across shards: ADC candidates with their exact scores, peer exchange, global re-rank merge
// (FdbVectorIndex.java:769,820-828,997-1043 with shards in the role of row ranges of one segment).
int32_t vs_adc_rerank_topk_exchange(uint64_t h, uint64_t comm, const float* q, int32_t nq, int32_t n_cand, int32_t k,
                                    int32_t metric, int32_t normalize_on_read, int64_t* ids_out, double* scores_out,
                                    int32_t* counts_out) {
  return exchange_host(1, h, comm, q, nq, nullptr, n_cand, k, metric, normalize_on_read, ids_out, scores_out, counts_out);
}
int32_t vs_adc_topk_exchange(uint64_t h, uint64_t comm, const float* q, int32_t nq, int32_t n_cand, int64_t* ids_out,
                             double* approx_out, int32_t* counts_out) {
  return exchange_host(2, h, comm, q, nq, nullptr, n_cand, n_cand, VS_METRIC_L2, 0, ids_out, approx_out, counts_out);
}
int32_t vs_rerank_topk_exchange(uint64_t h, uint64_t comm, const float* q, const int64_t* cand_ids, int32_t n_cand, int32_t k,
                                int32_t metric, int32_t normalize_on_read, int64_t* ids_out, double* scores_out,
                                int32_t* count_out) {
  return exchange_host(3, h, comm, q, 1, cand_ids, n_cand, k, metric, normalize_on_read, ids_out, scores_out, count_out);
}

// ADC top n_cand of this shard with the exact score of every candidate, packed for the cross-shard merge
int32_t vs_adc_rerank_packed_dev(uint64_t h, const float* d_q, int32_t nq, int32_t n_cand, int32_t metric,
                                 int32_t normalize_on_read, int64_t* d_pack, void* stream) {
  (void)normalize_on_read;  // same expression with norm(q) hoisted (J/fdb/FdbVectorIndex.java:1006-1010)
  std::shared_ptr<Segment> s_ref = seg_lookup(h);
  Segment* s = s_ref.get();
  RET(check_query_args(s, d_q, nq, n_cand, metric));
  if (!d_pack) return fail(VS_EINVAL, "null output pointer");
  if (s->M == 0 && s->n > 0) return fail(VS_ESTATE, "segment has no PQ attached");
  ThreadCtx* c;
  RET(ctx_bind_seg(&c, s));
  RET(ctx_use_stream(c, stream));
  if (s->n == 0) {  // an empty shard: every slot empty (state -1)
    CK(launch_fill_pack(d_pack, nq, n_cand, 1, nullptr, static_cast<cudaStream_t>(stream)), "fill launch");
    return VS_OK;
  }
  AdcPlan p;
  RET(plan_adc(s, nq, n_cand, &p));
  const size_t pb = (size_t)nq * p.partial_keys * 16, lb = (size_t)nq * s->M * s->K * 8;
  const size_t cib = (size_t)nq * n_cand * 8, ccb = (size_t)nq * 4, cdb = (size_t)nq * (p.cand_entries + p.extra_words) * 8;
  if (pb + cdb > (size_t(1) << 31)) return fail(VS_EINVAL, "query batch too large for the stream variant");
  const bool grow = Arena::need({pb, lb, cib, cib, ccb, cdb}) > c->d_cap || (size_t)nq > c->ticket_cap ||
                    (p.fast && (size_t)nq > c->fs_cap);
  RET(ctx_reserve_dev(c, Arena::need({pb, lb, cib, cib, ccb, cdb})));
  RET(ctx_reserve_ticket(c, nq));
  if (p.fast) RET(ctx_reserve_fs(c, nq));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (grow) CK(cudaStreamSynchronize(c->stream), "sync");
  Arena A(c->d_buf);
  ulonglong2* dpart = A.take<ulonglong2>((size_t)nq * p.partial_keys);
  double* dlut = A.take<double>((size_t)nq * s->M * s->K);
  int64_t* dcid = A.take<int64_t>((size_t)nq * n_cand);
  double* dcap = A.take<double>((size_t)nq * n_cand);
  int32_t* dccn = A.take<int32_t>(nq);
  unsigned long long* dcand = A.take<unsigned long long>((size_t)nq * (p.cand_entries + p.extra_words));
  RET(adc_dev(st, s, d_q, nq, n_cand, dlut, dcid, dcap, dccn, dpart, c->d_ticket, c->d_fs, dcand, p));
  RankLaunch L{};
  L.X = s->X; L.n = s->n; L.d = s->d; L.skip = s->skip; L.lanes = lanes(); L.q = d_q; L.nq = nq;
  L.cand_ids = dcid; L.nc = n_cand; L.k = n_cand; L.metric = metric; L.id_base = s->id_base;
  CK(launch_score_pack(L, dcap, dccn, d_pack, st), "score_pack launch");
  return VS_OK;
}

// ADC lists of this shard packed for the cross-shard merge by ascending approximate distance: [nq][2 n_cand] =
// ids | approximate distance bits (the layout of vs_bruteforce_topk_packed_dev)
int32_t vs_adc_topk_packed_dev(uint64_t h, const float* d_q, int32_t nq, int32_t n_cand, int64_t* d_pack, int32_t* d_counts,
                               void* stream) {
  std::shared_ptr<Segment> s_ref = seg_lookup(h);
  Segment* s = s_ref.get();
  RET(check_query_args(s, d_q, nq, n_cand, VS_METRIC_L2));
  if (!d_pack || !d_counts) return fail(VS_EINVAL, "null output pointer");
  if (s->M == 0 && s->n > 0) return fail(VS_ESTATE, "segment has no PQ attached");
  ThreadCtx* c;
  RET(ctx_bind_seg(&c, s));
  RET(ctx_use_stream(c, stream));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (s->n == 0) {
    CK(launch_fill_pack(d_pack, nq, n_cand, 0, d_counts, st), "fill launch");
    return VS_OK;
  }
  AdcPlan p;
  RET(plan_adc(s, nq, n_cand, &p));
  const size_t pb = (size_t)nq * p.partial_keys * 16, lb = (size_t)nq * s->M * s->K * 8;
  const size_t cdb = (size_t)nq * (p.cand_entries + p.extra_words) * 8;
  if (pb + cdb > (size_t(1) << 31)) return fail(VS_EINVAL, "query batch too large for the stream variant");
  const bool grow = Arena::need({pb, lb, cdb}) > c->d_cap || (size_t)nq > c->ticket_cap || (p.fast && (size_t)nq > c->fs_cap);
  RET(ctx_reserve_dev(c, Arena::need({pb, lb, cdb})));
  RET(ctx_reserve_ticket(c, nq));
  if (p.fast) RET(ctx_reserve_fs(c, nq));
  if (grow) CK(cudaStreamSynchronize(c->stream), "sync");
  Arena A(c->d_buf);
  ulonglong2* dpart = A.take<ulonglong2>((size_t)nq * p.partial_keys);
  double* dlut = A.take<double>((size_t)nq * s->M * s->K);
  unsigned long long* dcand = A.take<unsigned long long>((size_t)nq * (p.cand_entries + p.extra_words));
  return adc_dev(st, s, d_q, nq, n_cand, dlut, d_pack, reinterpret_cast<double*>(d_pack + n_cand), d_counts, dpart, c->d_ticket,
                 c->d_fs, dcand, p, 2 * (int64_t)n_cand);
}

// fetchExactAndScore over caller-supplied candidates when the rows are sharded: every shard scores the candidates it
// owns; the pack is the ADC + re-rank layout with the candidate's POSITION as its approximate key, so the cross-shard
// merge (vs_exchange_merge_adc_rerank_packed_dev with k results) keeps ties in candidate order (:1031).
int32_t vs_rerank_packed_dev(uint64_t h, const float* d_q, const int64_t* d_cand, int32_t n_cand, int32_t metric,
                             int32_t normalize_on_read, int64_t* d_pack, void* stream) {
  (void)normalize_on_read;
  std::shared_ptr<Segment> s_ref = seg_lookup(h);
  Segment* s = s_ref.get();
  RET(check_query_args(s, d_q, 1, 1, metric));
  if (!d_cand || !d_pack || n_cand <= 0 || n_cand > 8192) return fail(VS_EINVAL, "need candidates, an output and n_cand in 1..8192");
  ThreadCtx* c;
  RET(ctx_bind_seg(&c, s));
  RankLaunch L{};
  L.X = s->X; L.n = s->n; L.d = s->d; L.skip = s->skip; L.lanes = lanes(); L.q = d_q; L.nq = 1;
  L.cand_ids = d_cand; L.nc = n_cand; L.k = n_cand; L.metric = metric; L.id_base = s->id_base;
  CK(launch_score_pack(L, nullptr, nullptr, d_pack, static_cast<cudaStream_t>(stream), true), "score_pack launch");
  return VS_OK;
}

// ---- sharded PqTrainer.train over the library's own exchange ---------------------------------------------------
static int train_peer_begin(void* pcv, cudaStream_t st, size_t bytes, PeerXchg* x) {
  PeerComm* pc = static_cast<PeerComm*>(pcv);
  if (bytes > pc->slot_bytes)
    return fail(VS_EINVAL, "the training exchange needs %zu bytes per rank; the communicator's slots hold %zu", bytes, pc->slot_bytes);
  PeerSlot ps;
  if (!peer_next(pc, bytes, st, &ps))
    return fail(VS_ESTATE, "communicator of depth %d serves %d streams; this is one more", pc->depth, pc->depth / PEER_RING);
  x->gath = reinterpret_cast<const unsigned char*>(ps.gath);
  x->stride = bytes;
  x->flags = ps.flags;
  x->seq = ps.seq;
  x->data_off = ps.data_off;
  x->flag_off = ps.flag_off;
  x->ticket = ps.ticket;
  return VS_OK;
}
// orders the consumer behind the publish of rank src (src < 0: of every rank).  *enqueued = 1 when the wait has already
// happened (ranks share a device: the host waited, the consuming kernel must not poll), 0 when the kernel polls the flags.
static int train_peer_wait(void* pcv, cudaStream_t st, const PeerXchg* x, int src, int* enqueued) {
  PeerComm* pc = static_cast<PeerComm*>(pcv);
  *enqueued = 0;
  if (!pc->shared_device || g_peer_spin_shared.load() != 0) return VS_OK;
  RET(peer_host_wait(pc, st, x->flags, x->seq, src));
  *enqueued = 1;
  return VS_OK;
}
static int train_peer_publish(void* pcv, cudaStream_t st, const void* payload, const PeerXchg* x) {
  PeerComm* pc = static_cast<PeerComm*>(pcv);
  CK(launch_peer_publish(pc->bases, pc->world, pc->rank, payload, x->stride, x->data_off, x->flag_off, x->seq, x->ticket, st),
     "peer publish launch");
  return VS_OK;
}

// PqTrainer.train over row shards with the per-iteration all-reduce of cluster sums and counts done by libvsgpu over
// the peer buffers (K10): no callback, no collective library, no host synchronisation per reduction.  Collective:
// every rank of `comm` calls it with its own shard.  exact_order as in vs_pq_train_sharded.
int32_t vs_pq_train_sharded_peer(uint64_t h, uint64_t comm, int64_t n_total, int64_t row_lo, int32_t exact_order, int32_t M,
                                 int32_t K, int32_t iterations, int64_t seed, float* centroids_out) {
  std::shared_ptr<Segment> s_ref = seg_lookup(h);
  Segment* s = s_ref.get();
  if (!s) return fail(VS_EHANDLE, "unknown segment handle");
  PeerComm* pc = peer_lookup(comm);
  if (!pc || !pc->connected) return fail(VS_ESTATE, "peer communicator is not connected");
  const int d = s->d;
  if (M <= 0 || K <= 0) return fail(VS_EINVAL, "Invalid PQ params (m,k,dimension)");
  if (d % M != 0) return fail(VS_EINVAL, "dimension must be divisible by m");
  if (!centroids_out) return fail(VS_EINVAL, "null pointer");
  if (n_total <= 0) return fail(VS_EEMPTY, "empty training set (the reference throws IndexOutOfBoundsException)");
  if (n_total > 0x7fffffffLL) return fail(VS_EINVAL, "a Java List cannot hold more than 2^31-1 vectors");
  if (row_lo < 0 || row_lo + s->n > n_total) return fail(VS_EINVAL, "row range outside the corpus");
  if (pc->device != s->device) return fail(VS_EINVAL, "segment and communicator live on different devices");
  const size_t need = ((size_t)K * d + 3) / 4 * 16 + (size_t)M * K * 4 + 16;
  if (need > pc->slot_bytes) return fail(VS_EINVAL, "the training exchange needs %zu bytes per rank; the communicator's slots hold %zu", need, pc->slot_bytes);
  ThreadCtx* c;
  RET(ctx_bind_seg(&c, s));
  std::lock_guard<std::mutex> g(pc->xmu);  // the communicator's own stream: one collective at a time, same ring on every rank
  cudaStream_t st;
  RET(peer_xstream(pc, &st));
  TrainComm tc{row_lo, n_total, pc->rank, pc->world, exact_order != 0 ? 1 : 0, nullptr, nullptr, nullptr, nullptr};
  tc.peer = pc;
  tc.peer_begin = train_peer_begin;
  tc.peer_publish = train_peer_publish;
  tc.peer_wait = train_peer_wait;
  return pq_train_device(st, s->X, s->n, d, M, K, iterations, seed, lanes(), centroids_out, &tc);
}

int32_t vs_merge_adc_rerank_packed_dev(const int64_t* d_gath, int32_t world, int32_t nq, int32_t n_cand, int32_t k,
                                       int64_t* d_ids_out, double* d_scores_out, int32_t* d_counts_out, void* stream) {
  if (!d_gath || !d_ids_out || !d_scores_out || !d_counts_out) return fail(VS_EINVAL, "null pointer");
  if (world <= 0 || nq <= 0 || n_cand <= 0) return fail(VS_EINVAL, "world, nq and n_cand must be positive");
  if (k <= 0 || k > TOPK_MAX_K) return fail(VS_EINVAL, "k must be in 1..%d", TOPK_MAX_K);
  if ((int64_t)world * n_cand > 8192) return fail(VS_EINVAL, "world * n_cand must be <= 8192");
  ThreadCtx* c;
  RET(ctx_bind(&c));
  CK(launch_merge_adc_rerank(d_gath, world, nq, n_cand, k, d_ids_out, d_scores_out, d_counts_out,
                             static_cast<cudaStream_t>(stream)), "merge launch");
  return VS_OK;
}

int32_t vs_merge_topk_dev(const int64_t* d_ids, const double* d_scores, int64_t total, int32_t k,
                          int64_t* d_ids_out, double* d_scores_out, int32_t* d_count_out, void* stream) {
  if (!d_ids || !d_scores || !d_ids_out || !d_scores_out || !d_count_out || total <= 0) return fail(VS_EINVAL, "null pointer or empty input");
  if (k <= 0 || k > TOPK_MAX_K) return fail(VS_EINVAL, "k must be in 1..%d", TOPK_MAX_K);
  ThreadCtx* c;
  RET(ctx_bind(&c));
  CK(launch_merge(d_ids, d_scores, total, k, d_ids_out, d_scores_out, d_count_out, static_cast<cudaStream_t>(stream)), "merge launch");
  return VS_OK;
}

int32_t vs_adc_topk_dev(uint64_t h, const float* d_q, int32_t nq, int32_t n_cand, int64_t* d_ids,
                        double* d_approx, int32_t* d_counts, void* stream) {
  std::shared_ptr<Segment> s_ref = seg_lookup(h);
  Segment* s = s_ref.get();
  RET(check_query_args(s, d_q, nq, n_cand, VS_METRIC_L2));
  if (!d_ids || !d_approx || !d_counts) return fail(VS_EINVAL, "null output pointer");
  if (s->M == 0) return fail(VS_ESTATE, "segment has no PQ attached");
  ThreadCtx* c;
  RET(ctx_bind_seg(&c, s));
  RET(ctx_use_stream(c, stream));
  if (s->n == 0) return fail(VS_EINVAL, "empty segment: use the host variant");
  AdcPlan p;
  RET(plan_adc(s, nq, n_cand, &p));
  const size_t pb = (size_t)nq * p.partial_keys * 16, lb = (size_t)nq * s->M * s->K * 8;
  const size_t cdb = (size_t)nq * (p.cand_entries + p.extra_words) * 8;
  if (pb + cdb > (size_t(1) << 31)) return fail(VS_EINVAL, "query batch too large for the stream variant");
  const bool grow = Arena::need({pb, lb, cdb}) > c->d_cap || (size_t)nq > c->ticket_cap || (p.fast && (size_t)nq > c->fs_cap);
  RET(ctx_reserve_dev(c, Arena::need({pb, lb, cdb})));
  RET(ctx_reserve_ticket(c, nq));
  if (p.fast) RET(ctx_reserve_fs(c, nq));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (grow) CK(cudaStreamSynchronize(c->stream), "sync");
  Arena A(c->d_buf);
  ulonglong2* dpart = A.take<ulonglong2>((size_t)nq * p.partial_keys);
  double* dlut = A.take<double>((size_t)nq * s->M * s->K);
  unsigned long long* dcand = A.take<unsigned long long>((size_t)nq * (p.cand_entries + p.extra_words));
  return adc_dev(st, s, d_q, nq, n_cand, dlut, d_ids, d_approx, d_counts, dpart, c->d_ticket, c->d_fs, dcand, p);
}

int32_t vs_adc_rerank_topk_dev(uint64_t h, const float* d_q, int32_t nq, int32_t n_cand, int32_t k,
                               int32_t metric, int32_t normalize_on_read, int64_t* d_ids, double* d_scores,
                               int32_t* d_counts, void* stream) {
  (void)normalize_on_read;
  std::shared_ptr<Segment> s_ref = seg_lookup(h);
  Segment* s = s_ref.get();
  RET(check_query_args(s, d_q, nq, n_cand, metric));
  if (k <= 0 || k > TOPK_MAX_K) return fail(VS_EINVAL, "k must be in 1..%d", TOPK_MAX_K);
  if (!d_ids || !d_scores || !d_counts) return fail(VS_EINVAL, "null output pointer");
  if (s->M == 0) return fail(VS_ESTATE, "segment has no PQ attached");
  ThreadCtx* c;
  RET(ctx_bind_seg(&c, s));
  RET(ctx_use_stream(c, stream));
  if (s->n == 0) return fail(VS_EINVAL, "empty segment: use the host variant");
  AdcPlan p;
  RET(plan_adc(s, nq, n_cand, &p));
  const size_t pb = (size_t)nq * p.partial_keys * 16, lb = (size_t)nq * s->M * s->K * 8;
  const size_t cib = (size_t)nq * n_cand * 8, ccb = (size_t)nq * 4, cdb = (size_t)nq * (p.cand_entries + p.extra_words) * 8;
  if (pb + cdb > (size_t(1) << 31)) return fail(VS_EINVAL, "query batch too large for the stream variant");
  const bool grow = Arena::need({pb, lb, cib, cib, ccb, cdb}) > c->d_cap || (size_t)nq > c->ticket_cap ||
                    (p.fast && (size_t)nq > c->fs_cap);
  RET(ctx_reserve_dev(c, Arena::need({pb, lb, cib, cib, ccb, cdb})));
  RET(ctx_reserve_ticket(c, nq));
  if (p.fast) RET(ctx_reserve_fs(c, nq));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (grow) CK(cudaStreamSynchronize(c->stream), "sync");
  Arena A(c->d_buf);
  ulonglong2* dpart = A.take<ulonglong2>((size_t)nq * p.partial_keys);
  double* dlut = A.take<double>((size_t)nq * s->M * s->K);
  int64_t* dcid = A.take<int64_t>((size_t)nq * n_cand);
  double* dcap = A.take<double>((size_t)nq * n_cand);
  int32_t* dccn = A.take<int32_t>(nq);
  unsigned long long* dcand = A.take<unsigned long long>((size_t)nq * (p.cand_entries + p.extra_words));
  RET(adc_dev(st, s, d_q, nq, n_cand, dlut, dcid, dcap, dccn, dpart, c->d_ticket, c->d_fs, dcand, p));
  return rerank_dev(st, s, d_q, nq, dcid, n_cand, k, metric, d_ids, d_scores, d_counts);
}

}  // extern "C"
