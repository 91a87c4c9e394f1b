/*
 * vsgpu.h -- C ABI of libvsgpu, the B200-native (sm_100a) scoring path for VectorSearch.
 *
 * The reference (panghy/vectorsearch) has no plugin boundary for this path: the hot functions
 * are public static Java methods.  This header is the boundary a Java maintainer binds through
 * the FFM API (java.lang.foreign) or JNI; every entry point names the reference code it
 * replaces (J/ = src/main/java/io/github/panghy/vectorsearch/).  See INTEGRATION.md for the
 * Java-side stubs.
 *
 * Conventions
 *   - every function returns an int32 status: VS_OK or a negative VS_E* code;
 *     vs_last_error() returns a thread-local message for the last failure on this thread.
 *     VS_EINVAL maps to IllegalArgumentException, VS_EEMPTY to IndexOutOfBoundsException,
 *     everything else to IllegalStateException.
 *   - pointer arguments are caller-owned HOST memory, read or written only during the call,
 *     unless the function name ends in _dev (then they are DEVICE pointers and the work is
 *     enqueued on the given CUDA stream without synchronising).
 *   - device memory is owned by the library behind opaque uint64 handles.
 *   - layouts are row-major and contiguous: vectors fp32 little-endian (bit-identical to
 *     FloatPacker bytes, J/util/FloatPacker.java:21-39); codes uint8[n][M]; centroids
 *     float[M][K][subDim] (the flattened Java float[][][], equal to the PQCodebook blobs
 *     concatenated, J/tasks/SegmentBuildService.java:325-338); ids are int64 row indices
 *     (segment id_base + local row); scores are IEEE doubles carrying the reference's
 *     arithmetic (score = -l2 for L2, = similarity for COSINE, J/fdb/FdbVectorIndex.java:687-693).
 *   - one process drives one GPU (vs_init(device)); multi-GPU runs use one process per GPU
 *     and shard rows by range (see vectorsearch_b200/sharded.py).
 *   - thread-safe and re-entrant: each calling thread gets its own stream and scratch.
 *   - there is NO CPU fallback: every entry point fails with VS_ECUDA when no device is usable.
 */
#ifndef VSGPU_H
#define VSGPU_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VS_OK 0
#define VS_EINVAL (-1)   /* bad argument (IllegalArgumentException) */
#define VS_ENOMEM (-2)   /* host or device allocation failed */
#define VS_ECUDA (-3)    /* CUDA runtime error / no usable device */
#define VS_EHANDLE (-4)  /* unknown or freed handle */
#define VS_ESTATE (-5)   /* operation needs something the segment does not have (e.g. PQ) */
#define VS_EEMPTY (-6)   /* empty input where the reference throws IndexOutOfBounds */

#define VS_METRIC_L2 0      /* J/config/VectorIndexConfig.Metric.L2 */
#define VS_METRIC_COSINE 1  /* J/config/VectorIndexConfig.Metric.COSINE */

/* ---- lifecycle ------------------------------------------------------------------------- */
int32_t vs_version(void);
const char* vs_last_error(void);
/* Binds the calling process to one CUDA device.  Fails with VS_ECUDA if there is none. */
int32_t vs_init(int32_t device);
int32_t vs_shutdown(void);
/* Lane count of the JVM's FloatVector.SPECIES_PREFERRED being reproduced (J/util/Distances.java:15):
 * 16 (AVX-512, default), 8 (AVX2), 4 (NEON/SSE).  Affects only the low-order bits of scores. */
int32_t vs_set_simd_lanes(int32_t lanes);
int32_t vs_get_simd_lanes(void);
/* SM count, and bytes of free / total device memory */
int32_t vs_device_info(int32_t* sm_count, int64_t* free_bytes, int64_t* total_bytes);

/* ---- pair operations: the DistanceAndPqBenchmark surface -------------------------------------
 * Each call is one tiny kernel launch; they exist for API/known-answer parity, the throughput
 * path is the segment API below. */
int32_t vs_l2(const float* a, const float* b, int32_t len, double* out);          /* Distances.java:31-33 */
int32_t vs_l2_squared(const float* a, const float* b, int32_t len, double* out);  /* :48-64, :77-94 */
int32_t vs_dot(const float* a, const float* b, int32_t len, double* out);         /* :103-118 */
int32_t vs_norm(const float* a, int32_t len, double* out);                        /* :126-140 */
int32_t vs_cosine(const float* a, const float* b, int32_t len, double* out);      /* :149-153 */
/* PqEncoder.encode, J/pq/PqEncoder.java:18-37 */
int32_t vs_pq_encode(const float* centroids, int32_t M, int32_t K, int32_t subDim, const float* v,
                     uint8_t* codes_out);
/* pqLutDistance of the JMH suite (float LUT, float sum), B/DistanceAndPqBenchmark.java:116-123 */
int32_t vs_pq_lut_distance(const float* lut, int32_t M, int32_t K, const uint8_t* codes, float* out);
/* buildLut, J/fdb/FdbVectorIndex.java:1067-1079 -> double[M][K] */
int32_t vs_build_lut(const float* centroids, int32_t M, int32_t K, int32_t subDim, const float* q,
                     double* lut_out);
/* pqApproxDistance over n code rows, J/fdb/FdbVectorIndex.java:1057-1065 */
int32_t vs_pq_approx_distance(const double* lut, int32_t M, int32_t K, const uint8_t* codes,
                              int64_t n, double* out);

/* ---- segment residency ---------------------------------------------------------------------
 * A segment is a row range of vectors resident in HBM (plus, once sealed, its PQ codebook and
 * codes).  skip_mask (nullable, one byte per row, non-zero = skip) carries "deleted or gid
 * missing" (J/fdb/FdbVectorIndex.java:681,696,1000,1022). */
int32_t vs_segment_upload(const float* rows, int64_t n, int32_t d, const uint8_t* skip_mask,
                          int64_t id_base, uint64_t* handle_out);
/* Synthetic rows generated on the device: element (r, c) is draw (first_row + r) * d + c of
 * new java.util.Random(seed), mapped as nextFloat()*2f-1f (B/DistanceAndPqBenchmark.java:127-133). */
int32_t vs_segment_generate(int64_t seed, int64_t first_row, int64_t n, int32_t d, int64_t id_base,
                            uint64_t* handle_out);
int32_t vs_segment_set_skip(uint64_t h, const uint8_t* skip_mask /* nullable clears it */);
int32_t vs_segment_info(uint64_t h, int64_t* n, int32_t* d, int32_t* M, int32_t* K, int64_t* id_base);
int32_t vs_segment_download_rows(uint64_t h, int64_t first, int64_t count, float* rows_out);
/* Attach a codebook and codes; codes == NULL encodes the resident rows on the device. */
int32_t vs_segment_attach_pq(uint64_t h, const float* centroids, int32_t M, int32_t K,
                             const uint8_t* codes);
int32_t vs_segment_download_codes(uint64_t h, int64_t first, int64_t count, uint8_t* codes_out);
int32_t vs_segment_free(uint64_t h);

/* ---- query operations --------------------------------------------------------------------
 * Outputs are [nq][k] (or [nq][n_cand]); counts_out[i] entries of row i are valid, the rest
 * are id -1 / score NaN. */
/* searchBruteForceSegment scoring + stable sort + subList(0,k), J/fdb/FdbVectorIndex.java:676-721 */
int32_t vs_bruteforce_topk(uint64_t h, const float* q, int32_t nq, int32_t k, int32_t metric,
                           int64_t* ids_out, double* scores_out, int32_t* counts_out);
/* buildLut + full ADC scan + stable ascending sort + first n_cand, :741,:754-769,:820-822 */
int32_t vs_adc_topk(uint64_t h, const float* q, int32_t nq, int32_t n_cand, int64_t* ids_out,
                    double* approx_out, int32_t* counts_out);
/* fetchExactAndScore: candidates scored in the given order, ties keep it, :997-1043 */
int32_t vs_rerank_topk(uint64_t h, const float* q, const int64_t* cand_ids, int32_t n_cand, int32_t k,
                       int32_t metric, int32_t normalize_on_read, int64_t* ids_out,
                       double* scores_out, int32_t* count_out);
/* ADC top n_cand followed by exact re-rank to k in one call (config C4) */
int32_t vs_adc_rerank_topk(uint64_t h, const float* q, int32_t nq, int32_t n_cand, int32_t k,
                           int32_t metric, int32_t normalize_on_read, int64_t* ids_out,
                           double* scores_out, int32_t* counts_out);
/* cross-segment merge: stable sort by score descending of lists concatenated in segment order,
 * first k, J/fdb/FdbVectorIndex.java:432-437 */
int32_t vs_merge_topk(const int64_t* ids, const double* scores, int64_t total, int32_t k,
                      int64_t* ids_out, double* scores_out, int32_t* count_out);

/* ---- build operations ----------------------------------------------------------------------
 * Rows come from host memory (rows != NULL) or from a resident segment (rows == NULL, h != 0). */
/* PqTrainer.train, J/pq/PqTrainer.java:28-91 (production call: iterations 5, seed 42,
 * J/tasks/SegmentBuildService.java:180) */
int32_t vs_pq_train(const float* rows, uint64_t h, int64_t n, int32_t d, int32_t M, int32_t K,
                    int32_t iterations, int64_t seed, float* centroids_out);
/* PqTrainer.train over a corpus sharded by ascending row range over `world` processes (one per GPU): the
 * segment h holds rows [row_lo, row_lo + its row count) of n_total and this process is `rank`.  Every Lloyd
 * iteration the per-cluster fp32 sums and int counts of all shards are combined through the caller's
 * collective: the library fills the caller-owned DEVICE buffers d_comm_f32 (>= M*K*d/M floats) / d_comm_i32
 * (>= M*K ints), synchronises its stream and calls allreduce(user, kind, count) -- kind 0: sum
 * d_comm_f32[0..count) over all ranks in place, kind 1: the same for d_comm_i32 -- which must return 0 once
 * the reduced values are visible to other streams.  Rows another rank owns (initial centroids, re-initialised
 * empty clusters) travel through the same hook as zero-padded sums.  Every rank returns the same centroids.
 *   exact_order != 0: the sums continue rank after rank in ascending row order (world reductions per
 *     iteration, one contributing rank each): centroids are bit-identical to the reference / vs_pq_train.
 *   exact_order == 0: ONE all-reduce of the sums per iteration (north_star's scheme).  fp32 additions are
 *     re-associated across shards; k-means amplifies that last-bit difference as soon as one row changes
 *     cluster, so this is a different but statistically equivalent Lloyd trajectory, not the reference's. */
typedef int32_t (*vs_allreduce_fn)(void* user, int32_t kind, int64_t count);
int32_t vs_pq_train_sharded(uint64_t h, int64_t n_total, int64_t row_lo, int32_t rank, int32_t world, int32_t exact_order,
                            int32_t M, int32_t K, int32_t iterations, int64_t seed, float* d_comm_f32, int32_t* d_comm_i32,
                            vs_allreduce_fn allreduce, void* user, float* centroids_out);
/* PqEncoder.encode over n rows, J/pq/PqEncoder.java:18-37, J/tasks/SegmentBuildService.java:301 */
int32_t vs_pq_encode_batch(const float* centroids, int32_t M, int32_t K, int32_t subDim,
                           const float* rows, uint64_t h, int64_t n, uint8_t* codes_out);

/* ---- device-side (stream) variants -------------------------------------------------------------
 * Same semantics; q / outputs are DEVICE pointers, work is enqueued on `stream` (a cudaStream_t)
 * and NOT synchronised.  Used by the multi-GPU coordinator and for kernel-only timing. */
int32_t vs_bruteforce_topk_dev(uint64_t h, const float* d_q, int32_t nq, int32_t k, int32_t metric,
                               int64_t* d_ids, double* d_scores, int32_t* d_counts, void* stream);
int32_t vs_adc_topk_dev(uint64_t h, const float* d_q, int32_t nq, int32_t n_cand, int64_t* d_ids,
                        double* d_approx, int32_t* d_counts, void* stream);
int32_t vs_adc_rerank_topk_dev(uint64_t h, const float* d_q, int32_t nq, int32_t n_cand, int32_t k,
                               int32_t metric, int32_t normalize_on_read, int64_t* d_ids,
                               double* d_scores, int32_t* d_counts, void* stream);
/* Packed variants for the multi-GPU merge (one collective per query batch): d_pack is [nq][2k]
 * int64 -- k ids (id_base + row, -1 = empty slot) followed by the k score bit patterns. */
int32_t vs_bruteforce_topk_packed_dev(uint64_t h, const float* d_q, int32_t nq, int32_t k, int32_t metric,
                                      int64_t* d_pack, int32_t* d_counts, void* stream);
/* d_gath is the all-gathered [world][nq][2k] buffer: per query, lists concatenated in rank order,
 * stable sort by score descending (descending != 0) or distance ascending, first k
 * (J/fdb/FdbVectorIndex.java:432-437 with shards in the role of segments). */
int32_t vs_merge_packed_dev(const int64_t* d_gath, int32_t world, int32_t nq, int32_t k, int32_t descending,
                            int64_t* d_ids_out, double* d_scores_out, int32_t* d_counts_out, void* stream);
/* Cross-shard ADC + re-rank (config C4 on several GPUs).  The reference re-ranks the GLOBAL first n_cand rows
 * by approximate distance (J/fdb/FdbVectorIndex.java:769,820-828), so a shard ships its own n_cand candidates
 * WITH their exact scores: d_pack is [nq][4][n_cand] int64 = ids | approximate distance bits | exact score
 * bits | state (1 scored, 0 dropped by the re-rank: deleted or gid missing, -1 empty slot). */
int32_t vs_adc_rerank_packed_dev(uint64_t h, const float* d_q, int32_t nq, int32_t n_cand, int32_t metric,
                                 int32_t normalize_on_read, int64_t* d_pack, void* stream);
/* d_gath is the all-gathered [world][nq][4][n_cand] buffer.  Per query: the global first n_cand by
 * (approximate distance, rank, position) -- shards are ascending row ranges, so that is the reference's
 * stable order -- and of those the scored ones by exact score descending, ties in approximate order; first k. */
int32_t vs_merge_adc_rerank_packed_dev(const int64_t* d_gath, int32_t world, int32_t nq, int32_t n_cand, int32_t k,
                                       int64_t* d_ids_out, double* d_scores_out, int32_t* d_counts_out, void* stream);
/* ---- peer exchange: the all-gather and the merge as two kernels over NVLink peer memory -----------------
 * One process per GPU of one node.  vs_peer_create allocates this rank's communication buffer (`depth` slots of
 * [world][slot_bytes] plus arrival flags) and returns its 64-byte cudaIpc handle; the host exchanges the handles
 * of all ranks by any means (the coordinator uses one torch.distributed all-gather) and passes the rank-ordered
 * [world][64] array to vs_peer_connect.  vs_exchange_merge_* then replace "collective all-gather of d_pack +
 * vs_merge_*_packed_dev": the rank pushes its packed lists into every peer's buffer with plain stores and raises
 * a flag there; the merge kernel waits for the `world` flags of its own buffer and reads local memory only.
 * Results are those of the NCCL path, bit for bit.  Every rank must issue the same exchanges in the same order on
 * corresponding streams (as with any collective).  Each stream a communicator sees gets its own ring of 4 slots,
 * in order of first use, so exchanges on different streams may interleave freely; depth (a multiple of 4) = 4 x the
 * number of streams the communicator will serve -- one more stream is refused with VS_ESTATE.  A peer that never
 * arrives traps the waiting kernel after 20 s instead of hanging the GPU. */
#define VS_PEER_HANDLE_BYTES 64
int32_t vs_peer_create(int32_t rank, int32_t world, int64_t slot_bytes, int32_t depth, uint64_t* comm_out,
                       uint8_t* handle_out /* [VS_PEER_HANDLE_BYTES] */);
int32_t vs_peer_connect(uint64_t comm, const uint8_t* handles /* [world][VS_PEER_HANDLE_BYTES] */);
/* One process driving several GPUs (a JVM with one context per device; the tests with several communicators on one
 * GPU): no IPC -- pass the base addresses (vs_peer_base) of all ranks' buffers, peer access already enabled. */
int32_t vs_peer_base(uint64_t comm, uint64_t* base_out);
int32_t vs_peer_connect_ptrs(uint64_t comm, const uint64_t* bases /* [world] */);
int32_t vs_peer_destroy(uint64_t comm);
int32_t vs_exchange_merge_packed_dev(uint64_t comm, const int64_t* d_pack, int32_t nq, int32_t k, int32_t descending,
                                     int64_t* d_ids_out, double* d_scores_out, int32_t* d_counts_out, void* stream);
/* The whole sharded query as one HOST-buffer call (what a rank's request thread makes): H2D of the queries, local
 * scan, peer exchange, merge, results back (short lists are written by the merge kernel straight into pinned host
 * memory), one synchronisation.  Collective: every rank calls it with the same queries; results are identical on
 * every rank and equal to vs_bruteforce_topk over the concatenated shards. */
int32_t vs_bruteforce_topk_exchange(uint64_t h, uint64_t comm, const float* q, int32_t nq, int32_t k, int32_t metric,
                                    int64_t* ids_out, double* scores_out, int32_t* counts_out);
int32_t vs_adc_rerank_topk_exchange(uint64_t h, uint64_t comm, const float* q, int32_t nq, int32_t n_cand, int32_t k,
                                    int32_t metric, int32_t normalize_on_read, int64_t* ids_out, double* scores_out,
                                    int32_t* counts_out);
int32_t vs_exchange_merge_adc_rerank_packed_dev(uint64_t comm, const int64_t* d_pack, int32_t nq, int32_t n_cand,
                                                int32_t k, int64_t* d_ids_out, double* d_scores_out,
                                                int32_t* d_counts_out, void* stream);
int32_t vs_merge_topk_dev(const int64_t* d_ids, const double* d_scores, int64_t total, int32_t k,
                          int64_t* d_ids_out, double* d_scores_out, int32_t* d_count_out, void* stream);
/* number of kernels this library has launched so far in this process (for gpu_launches) */
int64_t vs_kernel_launch_count(void);
/* Tuning knobs (process-wide).  "adc_fast_min_rows": segments with fewer code rows use the generic
 * ADC kernel (default 16384); "adc_fast_cap": candidate-list entries per scan CTA of the fast ADC
 * scan (default 4096; on overflow the query is evaluated exactly over every row).
 * Batched brute force (query batches are nominated on the tensor cores and re-scored exactly, batch.cu):
 * "batch_min_queries" (default 3) and "batch_min_rows" (default 16384): smaller batches / segments use
 * the per-query scan; "batch_fp16" (default 1): nominate on an fp16 operand copy of the rows (n * d * 2
 * bytes of HBM, made at the first batched query) instead of the fp32 rows read as tf32; "batch_group"
 * (0 = automatic, 16 / 32 / 64): rows per nomination group; "batch_gm_bytes": scratch per query chunk;
 * "batch_warp_min_queries" (0 = automatic); "batch_pairs" (default 0): batches of more than 128 queries
 * nominate on CTA pairs (tcgen05 cta_group::2).  "scan_reserve_sms": SMs the one-query scan leaves free so that
 * work of another stream (the next query's prologue, a collective's CTAs) runs beside it (default 0).
 * "peer_fused" (default 1): a one-query peer exchange publishes inside the merge kernel (one launch) instead
 * of a publishing kernel followed by the merge.
 * "pdl" (default 1): the kernels of one call are chained with programmatic dependent launch (the next
 * kernel is set up while its predecessor runs and waits, in the kernel, for its results).
 * "pq_tensor_cores": how PQ assignment with 8-float sub-vectors nominates -- 2 (default) tcgen05 on fp16
 * hi/lo operand pairs (pq_tc.cu; needs n * M * 64 bytes of scratch, built slab by slab), 0 the FFMA kernel,
 * 1 mma.sync 3xTF32 (slower on B200, kept as a measured reference point).  "pq_tc_keep_bytes" (default 16 GiB):
 * how much of that scratch a private stream-ordered pool keeps cached between calls (0 = give everything back at
 * the next synchronisation).  Results never depend on any of them. */
int32_t vs_set_option(const char* name, int64_t value);

/* ---- diagnostics -----------------------------------------------------------------------------------
 * The batched path nominates rows with a(q, x) computed on the tensor cores (L2: |x|^2 - 2<q,x>,
 * COSINE: -<q,x>/|x|) and keeps every row whose value can be within `slack` of the exact one.  This
 * returns what the tensor-core stage produced -- per query the minimum of a over each group of
 * *group_out consecutive rows, gm_out[nq][*ngroups_out] -- and the slack per query, so that a test can
 * check |tensor-core value - exact value| <= slack.  gm_out == NULL only reports the two sizes. */
int32_t vs_debug_batch_groupmins(uint64_t h, const float* q, int32_t nq, int32_t metric, float* gm_out,
                                 int64_t gm_capacity, int64_t* ngroups_out, int32_t* group_out, double* slack_out);

#ifdef __cplusplus
}
#endif
#endif
