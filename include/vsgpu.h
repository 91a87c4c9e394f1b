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
 *   - devices: vs_init(device) binds one GPU; vs_init_multi(n, devices) binds several to ONE process (the JVM):
 *     segments uploaded afterwards are sharded by ascending row range over all of them and every query / build
 *     entry point fans out inside the library (worker thread per GPU, cross-shard exchange over NVLink peer
 *     memory) -- same results, bit for bit, as on one GPU.  One process per GPU (torchrun, MPI) is the other
 *     supported shape: vs_peer_* + the *_exchange entry points (see vectorsearch_b200/sharded.py).
 *   - thread-safe and re-entrant: each calling thread gets its own stream and scratch per device; a segment stays
 *     alive until every call that uses it has returned, whatever vs_segment_free does on another thread.
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
/* One process, several GPUs of one box (the reference's caller is one JVM fanning out with CompletableFuture.allOf,
 * J/fdb/FdbVectorIndex.java:418-437).  device_ids[0] is the primary device (pair operations, host-row builds).  With
 * n_gpus > 1 the devices must be identical and peer-accessible; vs_segment_upload / _generate / _upload_strided /
 * _upload_records then return SHARDED handles (rows split by ascending range), which every segment, query and build
 * entry point of this header accepts: vs_bruteforce_topk, vs_adc_topk, vs_adc_rerank_topk, vs_rerank_topk,
 * vs_pq_train (rows == NULL), vs_pq_encode_batch (rows == NULL), vs_segment_*, vs_adc_query_*, vs_adc_gather.
 * A device may be listed more than once (several ranks share it; this is how a one-GPU box exercises the
 * coordinator).  Re-binding (another vs_init / vs_init_multi) frees the previous group's sharded segments. */
int32_t vs_init_multi(int32_t n_gpus, const int32_t* device_ids);
int32_t vs_device_count(void); /* devices bound by the last vs_init / vs_init_multi */
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
/* The same from the reference's stored bytes: packed little-endian fp32 (FloatPacker.floatsToBytes,
 * J/util/FloatPacker.java:21-25 = VectorRecord.embedding, vectorsearch.proto:114-117), `stride` bytes from one
 * record's embedding to the next (>= d * 4).  One strided copy; nothing is decoded on the host. */
int32_t vs_segment_upload_strided(const uint8_t* bytes, int64_t n, int32_t d, int64_t stride, const uint8_t* skip_mask,
                                  int64_t id_base, uint64_t* handle_out);
/* ... and from serialized VectorRecord messages as they come out of the segment's range read
 * (J/fdb/FdbVectorIndex.java:676-699): record i is buf[offsets[i] .. offsets[i + 1]).  The embedding field goes to the
 * device through pinned staging, `deleted` becomes the row's skip flag (:681), vec_ids_out (nullable, [n]) receives each
 * row's vec_id.  VS_EINVAL for a malformed message or an embedding that is not d * 4 bytes long. */
int32_t vs_segment_upload_records(const uint8_t* buf, const int64_t* offsets, int64_t n, int32_t d, int64_t id_base,
                                  int32_t* vec_ids_out, uint64_t* handle_out);
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

/* ---- wire formats and residency (the steps either side of the scoring path) ---------------------------------
 * PQCodebook (vectorsearch.proto:135-142: int32 m = 1; int32 k = 2; repeated bytes centroids = 3, entry s = K * subDim
 * little-endian fp32) as SegmentBuildService.buildCodebookBytes writes it (J/tasks/SegmentBuildService.java:325-338)
 * and SegmentCaches.decodeCodebook reads it (J/cache/SegmentCaches.java:141-162).  vs_codebook_encode emits the bytes
 * protobuf-java's toByteArray() produces; out == NULL only reports the length.  vs_codebook_decode: centroids_out ==
 * NULL only reports M, K and subDim. */
int32_t vs_codebook_encode(const float* centroids, int32_t M, int32_t K, int32_t subDim, uint8_t* out, int64_t capacity,
                           int64_t* len_out);
int32_t vs_codebook_decode(const uint8_t* bytes, int64_t len, float* centroids_out, int64_t capacity_floats, int32_t* M_out,
                           int32_t* K_out, int32_t* subDim_out);
/* vs_segment_attach_pq from the stored PQCodebook message (codes == NULL encodes the resident rows on the device) */
int32_t vs_segment_attach_pq_codebook(uint64_t h, const uint8_t* codebook, int64_t len, const uint8_t* codes);
/* Residency table keyed by segment id and SegmentMeta.State (vectorsearch.proto:84: 0 ACTIVE, 1 PENDING, 2 SEALED,
 * 3 COMPACTING, 4 WRITING): upload on PENDING, attach PQ and put again on SEALED.  The table owns the handles it
 * holds.  put frees a previous copy of the same segment and, with a budget set, least-recently-used segments;
 * get returns VS_EHANDLE when the segment is not resident and VS_ESTATE (handle_out still set, for an in-place
 * upgrade) when it is resident in another state; invalidate drops the copy (compaction rebuild,
 * J/tasks/MaintenanceService.java:388-390, vacuum, segment deletion). */
int32_t vs_residency_put(int64_t seg_id, int32_t state, uint64_t handle);
int32_t vs_residency_get(int64_t seg_id, int32_t state, uint64_t* handle_out);
int32_t vs_residency_invalidate(int64_t seg_id);
int32_t vs_residency_set_budget(int64_t bytes /* 0 = unlimited */);
int32_t vs_residency_stats(int64_t* segments_out, int64_t* bytes_out);

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
/* BEST_FIRST expansion scoring (J/fdb/FdbVectorIndex.java:741,746-759,950-963): the LUT is built once per (query,
 * sealed segment) and kept on the device; every expansion step scores an id list against the RESIDENT codes
 * (pqApproxDistance, bit-exact) instead of reading all codes of the segment into a HashMap per query.  ids without a
 * code (outside the segment) get valid 0 and distance NaN (the reference skips them, :957).  valid_out is nullable. */
int32_t vs_adc_query_begin(uint64_t h, const float* q, uint64_t* query_out);
int32_t vs_adc_query_gather(uint64_t query, const int64_t* ids, int64_t n_ids, double* out, uint8_t* valid_out);
int32_t vs_adc_query_end(uint64_t query);
/* the three in one call */
int32_t vs_adc_gather(uint64_t h, const float* q, const int64_t* ids, int64_t n_ids, double* out, uint8_t* valid_out);
/* Graph-construction distances: GraphBuilder.buildL2Neighbors(vectors, degree) (J/graph/GraphBuilder.java:41-56)
 * when l_build <= 0, buildPrunedNeighbors(vectors, degree, l_build, alpha) (:73-109) otherwise, over the rows of a
 * resident segment (no skip mask: the builder takes every vector).  The O(n^2) distance work is the batched brute
 * force of this library with the segment as its own query set; lists are in the reference's order: l2Squared
 * ascending, ties to the lower row, the node itself excluded.  neighbors_out is int32[n][degree] (row indices, -1
 * padded), counts_out (nullable) int32[n].  degree <= 512, candidate lists (degree resp. l_build) <= 1023. */
int32_t vs_knn_graph(uint64_t h, int32_t degree, int32_t l_build, double alpha, int32_t* neighbors_out, int32_t* counts_out);
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
/* The same with the all-reduce done by libvsgpu itself over the peer buffers of `comm` (vs_peer_*): every rank pushes
 * its [sums | counts] into all peers' buffers over NVLink and a reduce kernel waits for the arrival flags -- no
 * callback, no collective library, no host synchronisation per reduction.  exact_order != 0: the running sums pass
 * from rank to rank inside that one exchange (bit-identical centroids); 0: fp32 sums combined in ascending rank order
 * (deterministic, identical on every rank, re-associated across shards).  Collective over the communicator; every
 * rank must own at least one row.  The slots of `comm` must hold M * K * (d / M + 1) * 4 bytes. */
int32_t vs_pq_train_sharded_peer(uint64_t h, uint64_t comm, int64_t n_total, int64_t row_lo, int32_t exact_order, int32_t M,
                                 int32_t K, int32_t iterations, int64_t seed, float* centroids_out);
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
/* ADC lists of this shard for a cross-shard merge by ascending approximate distance: d_pack is [nq][2 n_cand] =
 * ids | approximate distance bits (merge with vs_merge_packed_dev / vs_exchange_merge_packed_dev, descending = 0). */
int32_t vs_adc_topk_packed_dev(uint64_t h, const float* d_q, int32_t nq, int32_t n_cand, int64_t* d_pack, int32_t* d_counts,
                               void* stream);
/* fetchExactAndScore over caller-supplied candidates when the rows are sharded: each shard scores the candidates it
 * owns; d_pack is [4][n_cand] in the ADC + re-rank layout with the candidate's POSITION as its approximate key, so the
 * cross-shard merge keeps ties in candidate order (J/fdb/FdbVectorIndex.java:1031). */
int32_t vs_rerank_packed_dev(uint64_t h, const float* d_q, const int64_t* d_cand, int32_t n_cand, int32_t metric,
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
 * corresponding streams (as with any collective).  Ring 0 (4 slots) belongs to the communicator's own stream, on which
 * the host-buffer entry points (vs_*_exchange, vs_pq_train_sharded_peer) run whichever thread calls them; every caller
 * stream the communicator sees through the *_dev entry points gets a ring of its own, in order of first use
 * (vs_peer_release_stream hands one back).  depth (a multiple of 4) = 4 x (1 + the number of caller streams) -- one
 * more stream is refused with VS_ESTATE.  A one-query exchange is ONE kernel that publishes and then waits for its
 * peers: every rank's kernel must be able to run at the same time (one GPU per rank); communicators whose ranks share
 * a device (vs_peer_connect_ptrs) publish with a kernel of their own and let the calling HOST thread wait for the
 * arrival flags before the merge is launched (no wait on the device can be deadlock-free there; the exchange is then
 * blocking for that thread).  A rank whose shard is empty takes
 * part with an all-empty list.  A peer that never arrives traps the waiting kernel after 20 s instead of hanging the GPU. */
#define VS_PEER_HANDLE_BYTES 64
int32_t vs_peer_create(int32_t rank, int32_t world, int64_t slot_bytes, int32_t depth, uint64_t* comm_out,
                       uint8_t* handle_out /* [VS_PEER_HANDLE_BYTES] */);
int32_t vs_peer_connect(uint64_t comm, const uint8_t* handles /* [world][VS_PEER_HANDLE_BYTES] */);
/* One process driving several GPUs (a JVM with one context per device; the tests with several communicators on one
 * GPU): no IPC -- pass the base addresses (vs_peer_base) of all ranks' buffers, peer access already enabled. */
int32_t vs_peer_base(uint64_t comm, uint64_t* base_out);
int32_t vs_peer_connect_ptrs(uint64_t comm, const uint64_t* bases /* [world] */);
int32_t vs_peer_release_stream(uint64_t comm, void* stream);
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
/* The sharded query as ONE stream call per rank (local scan into the stream's packed send buffer, peer exchange, merge;
 * device or pinned-host outputs, nothing synchronises): what a pipelined coordinator issues per query. */
int32_t vs_bruteforce_topk_exchange_dev(uint64_t h, uint64_t comm, const float* d_q, int32_t nq, int32_t k, int32_t metric,
                                        int64_t* d_ids, double* d_scores, int32_t* d_counts, void* stream);
int32_t vs_adc_rerank_topk_exchange_dev(uint64_t h, uint64_t comm, const float* d_q, int32_t nq, int32_t n_cand, int32_t k,
                                        int32_t metric, int32_t normalize_on_read, int64_t* d_ids, double* d_scores,
                                        int32_t* d_counts, void* stream);
/* Host-buffer forms of the other two query operations over row shards (collective, identical results on every rank) */
int32_t vs_adc_topk_exchange(uint64_t h, uint64_t comm, const float* q, int32_t nq, int32_t n_cand, int64_t* ids_out,
                             double* approx_out, int32_t* counts_out);
int32_t vs_rerank_topk_exchange(uint64_t h, uint64_t comm, const float* q, const int64_t* cand_ids, int32_t n_cand, int32_t k,
                                int32_t metric, int32_t normalize_on_read, int64_t* ids_out, double* scores_out,
                                int32_t* count_out);
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
 * "batch_warp_min_queries" (default 0 = never): batches of at least that many queries select with one warp
 * per query instead of a CTA per query; "batch_pairs": whether batches of more than 128 queries nominate on
 * CTA pairs (tcgen05 cta_group::2) -- 0 never, 1 wherever it fits, 2 (default) for long vectors only, where both
 * operands stream through shared memory; "batch_select_ctas" (0 = automatic) and "batch_prefetch_rounds"
 * (default 1, this device only): diagnostics of the selection kernel.  "scan_fp16" (default 1): one or two queries
 * against a segment that has the fp16 operand copy are nominated by a CUDA-core scan of that copy (half the bytes
 * of the fp32 rows; ids and scores still come from the fp32 rows, bit-identical) -- 0: the fp32 streaming scan;
 * "scan_half_ctas" (0 = automatic, 1, 2): CTAs per SM of that scan (diagnostics).
 * "adc_reserve_sms" (default 0): the same for the fast ADC scan (queries alternating between streams: the LUT build,
 * re-rank and merge of one query run beside the scan of the next).
 * "scan_reserve_sms": SMs the one-query scan leaves free so that
 * work of another stream (the next query's prologue, a collective's CTAs) runs beside it (default 0).
 * "peer_fused" (default 1): a one-query peer exchange publishes inside the merge kernel (one launch) instead
 * of a publishing kernel followed by the merge.  "peer_spin_shared" (default 0, for tests of the polling kernels on a
 * one-GPU box): communicators whose ranks share a device poll inside the kernels like any other instead of letting
 * the host wait -- only safe when no large kernel of a peer has to run beside the polling one.
 * "pdl" (default 1): the kernels of one call are chained with programmatic dependent launch (the next
 * kernel is set up while its predecessor runs and waits, in the kernel, for its results).
 * "pq_tensor_cores": how PQ assignment with 8-float sub-vectors nominates -- 2 (default) tcgen05 on fp16
 * hi/lo operand pairs (pq_tc.cu; needs n * M * 64 bytes of scratch, built slab by slab), 0 the FFMA kernel,
 * 1 mma.sync 3xTF32 (slower on B200, kept as a measured reference point).  "train_exact_order" (default 1): vs_pq_train
 * on a SHARDED handle continues the cluster sums rank after rank (bit-identical centroids); 0 = one rank-ordered
 * all-reduce per iteration.  "pq_tc_keep_bytes" (default 16 GiB):
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
