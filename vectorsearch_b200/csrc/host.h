// host.h -- host-side declarations shared by the translation units of libvsgpu.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include <memory>
#include <mutex>
#include <vector>

namespace vs {

constexpr int VS_MAX_DEVICES = 16;  // CUDA ordinals 0..15 (one box)

// A row range of vectors resident in HBM (+ PQ codebook and codes once sealed).
struct Segment {
  float* X = nullptr;        // [n][d] fp32, row-major (bit-identical to FloatPacker bytes)
  uint8_t* skip = nullptr;   // nullable [n]: deleted or gid missing
  int64_t n = 0;
  int d = 0;
  int64_t id_base = 0;
  int device = 0;            // CUDA ordinal the rows live on
  float* centroids = nullptr;  // [M][K][subDim]
  uint8_t* codes = nullptr;    // [n][M]
  int M = 0, K = 0, subDim = 0;
  // batched-query nomination state (batch.cu), built at first use per metric under `mu`
  std::mutex mu;
  void* ab[2] = {nullptr, nullptr};     // float[n] nomination coefficients per metric
  void* stats[2] = {nullptr, nullptr};  // SegStats per metric (device)
  int nonfinite[2] = {0, 0};            // host copy of SegStats::nonfinite
  unsigned int xmax2 = 0;               // host copy of SegStats::xmax2_bits (same for both metrics)
  bool tm_ok = false;
  alignas(64) unsigned char tmX[128];   // CUtensorMap over X (fp32 rows, consumed as tf32)
  void* Xh = nullptr;                   // fp16 operand copy [n][dp], scaled by x_scale (a power of two)
  int dp = 0;
  float x_scale = 1.0f;
  bool xh_tried = false;                // the copy was attempted (it is optional: tf32 operands otherwise)
  alignas(64) unsigned char tmXh[128];  // CUtensorMap over Xh
  alignas(64) unsigned char tmX_b128[128], tmXh_b128[128];  // the same two with 128-row boxes (CTA pairs)
};

int fail(int code, const char* fmt, ...);
int cuda_fail(cudaError_t e, const char* what);
int lanes();
int sm_count();
// Entry points hold a reference for the duration of a call: vs_segment_free on another thread only drops the handle's.
std::shared_ptr<Segment> seg_lookup(uint64_t h);
// A group worker thread (group.cu) drives one GPU: its default device for calls that name no segment.
void set_thread_device(int cuda_dev);
int primary_device();
void set_error(const char* msg);  // thread-local message of vs_last_error()
void residency_clear();            // wire.cu: forget the residency table (vs_shutdown frees the handles wholesale)
int64_t nomination_aux_bytes(int64_t n, int d);  // api.cu: what a resident n x d segment may grow by at its first queries

// ---- group.cu: one process, several GPUs (vs_init_multi) -----------------------------------------------------
// A sharded segment is a row range split over the group's devices by ascending row range; its handle lives in a
// registry of its own.  The public entry points of api.cu forward such handles here.
bool group_is_sharded(uint64_t h);
bool group_wants_sharding();  // a group is up and the caller is not one of its workers: new segments are sharded
int group_size();  // devices bound by vs_init_multi when more than one, else 0
int group_start(int n, const int* cuda_devs);
int group_stop();
int group_segment_upload(const float* rows, int64_t n, int d, const uint8_t* skip, int64_t id_base, uint64_t* handle_out);
int group_segment_generate(int64_t seed, int64_t first_row, int64_t n, int d, int64_t id_base, uint64_t* handle_out);
int group_segment_set_skip(uint64_t h, const uint8_t* skip);
int group_segment_info(uint64_t h, int64_t* n, int32_t* d, int32_t* M, int32_t* K, int64_t* id_base);
int group_segment_download_rows(uint64_t h, int64_t first, int64_t count, float* out);
int group_segment_attach_pq(uint64_t h, const float* centroids, int M, int K, const uint8_t* codes);
int group_segment_download_codes(uint64_t h, int64_t first, int64_t count, uint8_t* out);
int group_segment_free(uint64_t h);
int group_bruteforce_topk(uint64_t h, const float* q, int nq, int k, int metric, int64_t* ids, double* scores, int32_t* counts);
int group_adc_topk(uint64_t h, const float* q, int nq, int n_cand, int64_t* ids, double* approx, int32_t* counts);
int group_adc_rerank_topk(uint64_t h, const float* q, int nq, int n_cand, int k, int metric, int nor, int64_t* ids,
                          double* scores, int32_t* counts);
int group_rerank_topk(uint64_t h, const float* q, const int64_t* cand, int n_cand, int k, int metric, int nor, int64_t* ids,
                      double* scores, int32_t* count);
int group_pq_train(uint64_t h, int64_t n, int d, int M, int K, int iterations, int64_t seed, float* centroids_out);
int group_pq_encode(uint64_t h, const float* centroids, int M, int K, int subDim, int64_t n, uint8_t* codes_out);
int group_adc_query_begin(uint64_t h, const float* q, std::vector<uint64_t>* shard_q);
int group_adc_query_gather(uint64_t h, const std::vector<uint64_t>& shard_q, const int64_t* ids, int64_t n_ids, double* out,
                           uint8_t* valid);
int group_adc_query_end(const std::vector<uint64_t>& shard_q);
int group_segment_upload_strided(const uint8_t* bytes, int64_t n, int d, int64_t stride, const uint8_t* skip, int64_t id_base,
                                 uint64_t* handle_out);
int group_segment_upload_records(const uint8_t* buf, const int64_t* offsets, int64_t n, int d, int64_t id_base,
                                 int32_t* vec_ids_out, uint64_t* handle_out);
void group_set_train_exact(int v);

// Row-sharded training: the caller's collective hook and the device buffers it reduces (sum, in place)
// one exchange over the peer buffers (api.cu): where the `world` published copies land in THIS rank's buffer
struct PeerXchg {
  const unsigned char* gath;  // copy of rank r at gath + r * stride
  size_t stride;
  const unsigned long long* flags;  // arrival flag of rank r reaches seq
  unsigned long long seq;
  size_t data_off, flag_off;  // internal: where this rank's copy goes in every peer's buffer
  unsigned int* ticket;
};
struct TrainComm {
  int64_t row_lo, n_total;
  int rank, world;
  int exact_order;  // 1: sums continue rank after rank in row order (bit-exact); 0: one all-reduce per iteration
  // callback transport (vs_pq_train_sharded): the caller's collective
  void* user;
  int32_t (*allreduce)(void* user, int32_t kind, int64_t count);  // kind 0: d_f32[0..count), 1: d_i32[0..count)
  float* d_f32;    // >= M * K * subDim floats
  int32_t* d_i32;  // >= M * K ints
  // peer transport (vs_pq_train_sharded_peer): libvsgpu's own exchange; peer != nullptr selects it
  void* peer = nullptr;
  int (*peer_begin)(void* peer, cudaStream_t st, size_t bytes, PeerXchg* x) = nullptr;    // reserves the next slot
  int (*peer_publish)(void* peer, cudaStream_t st, const void* payload, const PeerXchg* x) = nullptr;
  // orders st behind rank src's publish (src < 0: all); *enqueued = 1: done on the stream, the kernel must not poll
  int (*peer_wait)(void* peer, cudaStream_t st, const PeerXchg* x, int src, int* enqueued) = nullptr;
};
// PqTrainer.train on device-resident rows; centroids_out is HOST memory [M][K][d/M]  (pqtrain.cu)
int pq_train_device(cudaStream_t st, const float* dX, int64_t n, int d, int M, int K, int iterations,
                    int64_t seed, int lanes, float* centroids_out, const TrainComm* comm = nullptr);

#ifdef VS_BQ_STAMPS
int debug_read_stamps_batch(void* dst, size_t bytes);  // development only (batch.cu)
#endif
#ifdef VS_PHASE_STAMPS
int debug_read_stamps(void* dst, size_t bytes);  // development only (scan.cu)
int debug_read_stamps_adc(void* dst, size_t bytes);  // development only (adc_fast.cu)
#endif

}  // namespace vs
