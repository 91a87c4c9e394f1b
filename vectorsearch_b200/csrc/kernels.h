// kernels.h -- host-side launch interface between api.cu and the kernel translation units.
#pragma once

#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

namespace vs {

constexpr int SCAN_THREADS = 256;

void count_launch();  // bumps the process-wide kernel launch counter (api.cu)

// Launch with programmatic stream serialisation: the kernel may be set up (and its CTAs placed, resources
// permitting) while its predecessor in the stream is still running.  Every kernel launched this way executes
// pdl_wait() before it touches anything the predecessor wrote; pdl_trigger() at the top of a kernel lets ITS
// successor start early.  Both are no-ops without the attribute.
bool pdl_enabled();  // option "pdl" (api.cu)
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl_if(bool want, void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = (want && pdl_enabled()) ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args... args) {
  return launch_pdl_if(true, kern, grid, block, smem, st, args...);
}

// ---- scan.cu: exact brute-force scan + top-k, one launch --------------------------------------
struct ScanLaunch {
  const float* X;        // [n][d] resident rows
  int64_t n;
  int d;
  const float* q;        // [nq][d] device
  int nq;
  const uint8_t* skip;   // nullable [n]
  int lanes;             // modelled SIMD lane count (16 / 8 / 4)
  bool cosine;
  int k, kp;             // kp = topk_pad(k)
  ulonglong2* partial;   // scratch [nq][grid][k]
  unsigned long long* ctrl;  // scratch [nq][4], zero between launches
  int64_t partial_keys;      // keys of `partial` per query (filled by scan_configure)
  int64_t* ids_out;      // [nq][k]
  double* scores_out;    // [nq][k]
  int32_t* counts_out;   // [nq]
  int64_t id_base;
  int64_t out_stride;    // elements between queries in ids_out / scores_out (0 = k)
  // filled by scan_configure:
  int variant;           // ScanVariant
  int grid, threads;
  size_t smem_bytes;
  int TR, NS;            // TMA ring: rows per stage, stages per warp
};
enum ScanVariant { SCAN_TMA = 0, SCAN_LDG = 1, SCAN_ROWTHREAD = 2 };
bool scan_is_streaming(int d, int lanes, bool cosine);
// chooses the kernel variant and launch shape from (n, d, nq, lanes, cosine, k); false = not launchable
bool scan_configure(ScanLaunch& L, int sms);
cudaError_t launch_scan(const ScanLaunch& L, cudaStream_t st);

// ---- batch.cu: batched-query brute force -- tensor-core nomination + exact re-score ----------------
struct SegStats {
  unsigned int xmax2_bits;  // bit pattern of max |x|^2 over live rows
  int nonfinite;            // a live row has a non-finite (or absurdly large) norm
};
// per-row nomination coefficient for one metric + SegStats; one pass over the segment
cudaError_t launch_row_prep(const float* X, int64_t n, int d, const uint8_t* skip, bool cosine, float* coef,
                            SegStats* stats, int sms, cudaStream_t st);
struct BatchLaunch {
  const float* X;
  int64_t n;
  int d;
  const uint8_t* skip;
  int lanes;
  bool cosine;
  const float* q;        // [nq][d] device
  int nq;
  int k, kp;
  const void* tmX;       // host copy of the CUtensorMap (128 bytes) over the operand rows (fp32 rows or their fp16 copy)
  const void* tmX128;    // the same with a 128-row box (CTA pairs load half a row tile each), nullable
  bool pairs;            // allow the cta_group::2 kernel for batches of more than 128 queries
  bool half;             // operands are fp16 copies (kind::f16) instead of the fp32 data read as tf32
  int dp;                // fp16 row pitch in elements (d rounded up to 8)
  float x_scale;         // power of two the fp16 row copy was scaled by
  const void* xh;        // [n][dp] the fp16 operand copy itself (half only)
  bool prefilter;        // candidate groups: fp16 pre-filter before the exact scores (half only)
  void* qh;              // [nq][dp] fp16 scratch for the scaled queries (half only)
  float* qinv;           // [round_up(nq, 256)] 1 / (x_scale * query scale) (half only)
  const float* coef;     // [n] per-row nomination coefficient of this metric
  const SegStats* stats; // device
  float* gm;             // [round_up(nq, 256)][gm_stride] group minima
  int32_t* fb;           // [1 + 2 nq]: fallback count, fallback query list, per-query "listed" flags
  ulonglong2* partial;   // [nq][partial_keys]
  int64_t partial_keys;
  unsigned long long* ctrl;  // [nq][4], zero between launches
  int64_t* ids_out;
  double* scores_out;
  int32_t* counts_out;
  int64_t id_base;
  int64_t out_stride;
  int select_ctas_override;  // 0 = choose the selection CTAs per query from the batch size
  int sh_override;           // scan_half_kernel CTAs per SM: 0 = automatic, 1 = one, 2 = two where they fit (diagnostics)
  int group_override;    // 0 = choose the rows per nomination group (16 / 32 / 64) from the shape
  bool direct;           // one or two queries: scan_half_kernel instead of the tensor-core nomination (needs sh_ok)
  int reserve_sms;       // ... on this many SMs fewer (room for a neighbouring stream's small kernels)
  bool gemm_only;        // diagnostics: stop after the group minima
  int warp_min_q;        // chunks of at least this many queries select with one warp per query (k <= 32)
  // filled by batch_configure:
  int64_t tiles, ngroups, gm_stride;
  int group;             // rows per nomination group
  int cap, sms, fb_gx, fb_threads, fbd_gx, fbd_threads;  // (fbd_*: the fallback check behind scan_half_kernel)
  size_t gemm_smem, select_smem, selw_smem, fb_smem, fbd_smem;
  bool gemm_stat;        // query block resident in shared memory (short vectors)
  int gemm_stages;
  int pair_stages;       // stages of the cta_group::2 kernel
  bool pair_stat;        // ... with the query block resident per CTA (else both operands stream)
  size_t pair_smem;
  bool sh_ok;            // scan_half_kernel fits this shape (fp16 copy, k <= 32)
  int sh_TR, sh_NS, sh_kk, sh_cpl, sh_grid, sh_threads, sh_per_sm;
  size_t sh_smem;
};
bool batch_supported(int d, int lanes, bool cosine, int64_t n);
double batch_slack_host(bool cosine, bool half, int d, double xmax, double qn);
bool batch_encode_segment_map(void* tm128, const void* rows, int64_t n, int d, int64_t pitch, bool half, int box_rows);
cudaError_t launch_row_convert(const float* X, int64_t n, int d, int dp, float sx, void* Xh, int sms, cudaStream_t st);
bool batch_configure(BatchLaunch& L, int sms);
int batch_set_prefetch_rounds(int rounds);  // diagnostics: rounds of 8 rows the pre-filter prefetches ahead (-1: none)          // from (n, d, lanes, cosine, k, half)
int64_t batch_partial_keys(const BatchLaunch& L, int nq);  // keys of `partial` per query for a chunk of nq
cudaError_t launch_batch(const BatchLaunch& L, cudaStream_t st);  // one chunk of L.nq queries

// ---- rank.cu: exact scoring + ordering of caller-supplied candidates (re-rank), merge ----------
struct RankLaunch {
  const float* X;
  int64_t n;
  int d;
  const uint8_t* skip;
  int lanes;
  const float* q;           // [nq][d]
  int nq;
  const int64_t* cand_ids;  // [nq][nc] global ids (id_base + row), anything else = missing record
  int nc;
  int k;
  int metric;               // VS_METRIC_*
  int64_t id_base;
  int64_t* ids_out;         // [nq][k]
  double* scores_out;
  int32_t* counts_out;
};
cudaError_t launch_rank(const RankLaunch& L, cudaStream_t st);
// cross-shard ADC + re-rank: per query int64[4][nc] = ids | approx bits | exact score bits | state
// cand_approx == nullptr: the candidate's position is its approximate key (caller-supplied candidates re-ranked across
// shards); cand_counts == nullptr: all nc are valid; foreign_empty: ids outside [id_base, id_base + n) are another
// shard's (state -1) instead of missing records (state 0)
cudaError_t launch_score_pack(const RankLaunch& L, const double* cand_approx, const int32_t* cand_counts, int64_t* pack,
                              cudaStream_t st, bool foreign_empty = false);
// an empty shard's contribution: kind 0 = [nq][2k] lists (ids -1, NaN), kind 1 = [nq][4][k] ADC + re-rank packs (state -1)
cudaError_t launch_fill_pack(int64_t* pack, int nq, int k, int kind, int32_t* counts, cudaStream_t st);
// wait_flags != nullptr: the gathered lists arrive through the peer exchange; the kernel first waits until
// wait_flags[0..w) have reached wait_seq (launch_peer_publish of every rank).  pub != nullptr and nq == 1: the
// kernel publishes this rank's list itself first (no separate launch_peer_publish for that exchange).
struct PeerPublish {
  unsigned char* const* bases;
  const void* payload;
  size_t bytes, data_off, flag_off;
  int rank;
};
cudaError_t launch_merge_adc_rerank(const int64_t* gath, int w, int nq, int nc, int k, int64_t* ids_out,
                                    double* scores_out, int32_t* counts_out, cudaStream_t st,
                                    const unsigned long long* wait_flags = nullptr, unsigned long long wait_seq = 0,
                                    const PeerPublish* pub = nullptr);
// Peer exchange (rank.cu): copy `bytes` of payload to bases[p] + data_off of every peer p, then set
// flag word `rank` at bases[p] + flag_off to seq.  ticket: one zeroed word of this rank per concurrent exchange.
constexpr int VS_PEER_MAX_WORLD = 16;
constexpr unsigned long long VS_PEER_TIMEOUT_NS = 20ull * 1000 * 1000 * 1000;
cudaError_t launch_peer_publish(unsigned char* const* bases, int w, int rank, const void* payload, size_t bytes,
                                size_t data_off, size_t flag_off, unsigned long long seq, unsigned int* ticket,
                                cudaStream_t st);
// All-reduce over the peer buffers (K10): the `w` published copies of [nf floats | ni ints] (each `stride` bytes apart,
// ints at the 16-byte-rounded end of the floats) are combined after the arrival flags reach seq.  mode 0: fp32 sum in
// ascending rank order, 1: bitwise OR (exact broadcast of rows owned by one rank), 2: rank w-1's copy; ints: sum.
cudaError_t launch_peer_reduce(const void* gath, int w, size_t stride, int64_t nf, int64_t ni, int mode, float* out_f,
                               int32_t* out_i, const unsigned long long* flags, unsigned long long seq, cudaStream_t st);
// orders the stream behind the publish of ONE source rank
cudaError_t launch_peer_wait_one(const unsigned long long* flags, int src, unsigned long long seq, cudaStream_t st);
// Graph construction (GraphBuilder.buildL2Neighbors / buildPrunedNeighbors): node row0 + q's nominated candidates
// (brute-force order, the node itself possibly among them) -> the reference's neighbour list in (l2Squared, j) order,
// optionally pruned.  flags[q] = 1: the candidate list may have cut a group of equal-distance rows: redo exactly.
cudaError_t launch_knn_finalize(const float* X, int64_t n, int d, int lanes, int64_t row0, int nrows, const int64_t* cand_ids,
                                const double* cand_scores, const int32_t* cand_counts, int kq, int take, int degree, int keep,
                                bool check_closure, double alpha, bool prune, int64_t id_base, int32_t* neighbors,
                                int32_t* counts, int32_t* flags, cudaStream_t st);
// exact candidates of listed nodes: the kq smallest (l2Squared, j), j != node, written like a brute-force result
cudaError_t launch_knn_exact(const float* X, int64_t n, int d, int lanes, const int32_t* nodes, int nnodes, int kq, int64_t id_base,
                             int64_t* cand_ids, double* cand_scores, int32_t* cand_counts, cudaStream_t st);
// stable sort by score descending of `total` (id, score) pairs, first k
cudaError_t launch_merge(const int64_t* ids, const double* scores, int64_t total, int k,
                         int64_t* ids_out, double* scores_out, int32_t* count_out, cudaStream_t st);
// gathered [w][nq][2k] packed lists (k ids then k score bit patterns; id < 0 = empty slot):
// per query, concatenate in rank order, stable sort (descending scores or ascending distances), first k
cudaError_t launch_merge_packed(const int64_t* gath, int w, int nq, int k, bool descending,
                                int64_t* ids_out, double* scores_out, int32_t* counts_out, cudaStream_t st,
                                const unsigned long long* wait_flags = nullptr, unsigned long long wait_seq = 0,
                                const PeerPublish* pub = nullptr);

// ---- misc.cu ----------------------------------------------------------------------------------
// element e of the output = draw (first + e) of new java.util.Random(seed); kind 0: nextFloat()*2f-1f,
// 1: nextFloat(), 2: nextFloat()*10f  (B/DistanceAndPqBenchmark.java:66-85,127-133)
cudaError_t launch_generate(float* out, int64_t count, int64_t seed, int64_t first, int kind,
                            cudaStream_t st);
enum PairOp { PAIR_L2 = 0, PAIR_L2SQ = 1, PAIR_DOT = 2, PAIR_NORM = 3, PAIR_COSINE = 4 };
cudaError_t launch_pair(int op, const float* a, const float* b, int len, int lanes, double* out,
                        cudaStream_t st);
cudaError_t launch_lut_distance_f32(const float* lut, int M, int K, const uint8_t* codes, float* out,
                                    cudaStream_t st);

// ---- adc.cu -----------------------------------------------------------------------------------
// lut64[nq][M][K] in reference arithmetic (buildLut, J/fdb/FdbVectorIndex.java:1067-1079)
cudaError_t launch_build_lut(const float* centroids, int M, int K, int subDim, const float* q, int nq,
                             int lanes, double* lut64, cudaStream_t st);
cudaError_t launch_approx_distance(const double* lut, int M, int K, const uint8_t* codes, int64_t n,
                                   double* out, cudaStream_t st);
// pqApproxDistance of listed ids (global: id_base + row) against resident codes; ids without a code: valid 0, NaN
cudaError_t launch_adc_gather(const double* lut, int M, int K, const uint8_t* codes, int64_t n, int64_t id_base,
                              const int64_t* ids, int64_t n_ids, double* out, uint8_t* valid, cudaStream_t st);
struct AdcScanLaunch {
  const uint8_t* codes;  // [n][M]
  int64_t n;
  int M, K;
  const double* lut64;   // [nq][M][K]
  int nq;
  int k, kp;
  ulonglong2* partial;
  unsigned long long* ctrl;
  int64_t partial_keys;
  int64_t* ids_out;
  double* approx_out;
  int32_t* counts_out;
  int64_t id_base;
  int64_t out_stride;
  int grid, threads;
  size_t smem_bytes;
};
// fills kp / threads / grid / smem_bytes from (n, M, K, nq, k); false = not launchable
bool adc_configure(AdcScanLaunch& L, int sms);
cudaError_t launch_adc_scan(const AdcScanLaunch& L, cudaStream_t st);

// ---- adc_fast.cu: conflict-free byte-LUT scan + exact candidate ranking (M = 8 or 16, K <= 256) ----
constexpr int FS_THREADS = 1024;
constexpr int FS_U = 8;                     // rows in flight per thread (128 KB per SM: the HBM latency x bandwidth product)
constexpr int FS_WU = 4;                    // rows per thread of the warm-up sample
#ifndef VS_FS_PF
#define VS_FS_PF 3
#endif
constexpr int FS_PF = VS_FS_PF;              // batches requested into L2 ahead of the loads (-DVS_FS_PF=.. for experiments)
constexpr int FS_BINS = 4096;               // histogram bins of the integer row sum (<= 255 * 16)
constexpr int FS_CTRL = 8;                  // control words: ticket, ~T, fallback flag
constexpr int FS_TICKET = 0, FS_TINV = 1, FS_FLAG = 2, FS_NEXT = 3;  // FS_NEXT: dynamic batch counter
constexpr int FS_MAX_GRID = 256;            // scan CTAs per query (one candidate list each)
constexpr int FS_WORDS = FS_BINS + FS_CTRL + FS_MAX_GRID;  // per-query persistent scratch (uint32)
constexpr unsigned int FS_T_INF = 0xfffffffeu;  // "no threshold yet"; 0xffffffff marks rows that do not exist
struct AdcFastLaunch {
  const uint8_t* codes;
  int64_t n;
  int M, K;
  const double* lut64;       // [nq][M][K]
  const unsigned long long* mm;  // [nq][M][2] per-subspace {min, max} images of the LUT (build_lut_mm)
  int nq;
  int k, kp;
  unsigned int* fs;          // [nq][FS_WORDS], zero between launches
  unsigned long long* cand;  // [nq][grid][cap] per-CTA candidate lists (S << 48 | row)
  unsigned int cap;          // entries per CTA list
  ulonglong2* partial;       // final kernel: per-CTA lists
  unsigned long long* ctrl;
  int64_t partial_keys;
  int64_t* ids_out;
  double* approx_out;
  int32_t* counts_out;
  int64_t id_base;
  int64_t out_stride;
  int grid;                  // scan CTAs per query
  int reserve_sms;           // ... of which this many are not launched (room for a neighbouring stream's small kernels)
  size_t smem_bytes;
  int final_grid, final_threads;
  size_t final_smem;
};
bool adc_fast_supported(int M, int K);
bool adc_fast_configure(AdcFastLaunch& L, int sms);
// buildLut for nq queries plus the per-subspace extremes the fast scan quantises with (M <= 16)
cudaError_t launch_build_lut_mm(const float* centroids, int M, int K, int subDim, const float* q, int nq, int lanes,
                                double* lut64, unsigned long long* mm, cudaStream_t st);
cudaError_t launch_adc_fast(const AdcFastLaunch& L, cudaStream_t st);
cudaError_t launch_adc_fallback(const AdcFastLaunch& L, cudaStream_t st);
int debug_adc_stats(unsigned int* out8);  // development: {candidates, T_final, full-scan flag} of the last query 0

// ---- pq.cu ------------------------------------------------------------------------------------
// PqEncoder.encode / the PqTrainer assignment step over n rows (J/pq/PqEncoder.java:18-37,
// J/pq/PqTrainer.java:56-68): strict-< argmin over K centroids per subspace, reference arithmetic.
struct PqAssignLaunch {
  const float* X;          // [n][d]
  int64_t n;
  int d, M, K, subDim;
  const float* centroids;  // [M][K][subDim]
  int lanes;
  uint8_t* codes_u8;       // nullable: [n][M] (best & 0xFF)
  int32_t* assign_i32;     // nullable: [M][n]
  int s_begin, s_end;      // subspaces to process
};
// subDim 8 nomination: 0 = FFMA kernel, 1 = mma.sync 3xTF32, 2 = tcgen05 (pq_tc.cu)
void pq_set_tensor_cores(int mode);
bool pq_tc_supported(const PqAssignLaunch& L);
cudaError_t launch_pq_assign_tc(const PqAssignLaunch& L, int sms, cudaStream_t st);
// Between begin and end (same host thread) the rows' operand image is built once and reused by every
// launch_pq_assign_tc over the same rows (the Lloyd iterations of PqTrainer.train).
void pq_tc_scope_begin();
void pq_tc_scope_end();
void pq_tc_set_keep_bytes(unsigned long long bytes);  // scratch the pool keeps cached between calls
// stream-ordered scratch from that pool (released with cudaFreeAsync on the same stream)
cudaError_t pq_pool_alloc(void** p, size_t bytes, cudaStream_t st);
cudaError_t launch_pq_assign(const PqAssignLaunch& L, cudaStream_t st);

}  // namespace vs
