// host.h -- host-side declarations shared by the translation units of libvsgpu.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include <mutex>

namespace vs {

// A row range of vectors resident in HBM (+ PQ codebook and codes once sealed).
struct Segment {
  float* X = nullptr;        // [n][d] fp32, row-major (bit-identical to FloatPacker bytes)
  uint8_t* skip = nullptr;   // nullable [n]: deleted or gid missing
  int64_t n = 0;
  int d = 0;
  int64_t id_base = 0;
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
Segment* seg_lookup(uint64_t h);

// Row-sharded training: the caller's collective hook and the device buffers it reduces (sum, in place)
struct TrainComm {
  int64_t row_lo, n_total;
  int rank, world;
  int exact_order;  // 1: sums continue rank after rank in row order (bit-exact); 0: one all-reduce per iteration
  void* user;
  int32_t (*allreduce)(void* user, int32_t kind, int64_t count);  // kind 0: d_f32[0..count), 1: d_i32[0..count)
  float* d_f32;    // >= M * K * subDim floats
  int32_t* d_i32;  // >= M * K ints
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
