/*
 * vs_oracle.h -- CPU ORACLE for the scoring hot path of panghy/vectorsearch.
 *
 * THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load it.  The product
 * (libvsgpu.so, vectorsearch_b200/) never links, imports or calls anything in oracle/.
 *
 * It is a plain-C restatement of the reference's Java arithmetic.  The reference
 * cannot be compiled or run in the build container (no JVM), so there is no
 * oracle/_ref; parity is pinned by (a) the reference's own known-answer tests
 * (DistancesTest, PqEncoderTest, PqTrainerTest, VectorIndexTest.l2_query...), which
 * tests/test_oracle_golden.py replays against this file, and (b) well-known
 * java.util.Random outputs.  Everything the reference's tests leave unpinned
 * (k-means centroid values, ADC lists, re-rank order, cosine queries) is
 * "parity pinned by restatement only" -- see DESIGN.md.
 *
 * Paths: J/ = /root/reference/src/main/java/io/github/panghy/vectorsearch/
 *        B/ = /root/reference/src/jmh/java/io/github/panghy/vectorsearch/bench/
 */
#ifndef VS_ORACLE_H
#define VS_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VSO_METRIC_L2 0
#define VSO_METRIC_COSINE 1

/* ---- java.util.Random (JDK core; 48-bit LCG) ------------------------------- */
typedef struct {
  uint64_t state;
} vso_jrandom;

void vso_jr_init(vso_jrandom* r, int64_t seed);
int32_t vso_jr_next(vso_jrandom* r, int bits);
int32_t vso_jr_next_int(vso_jrandom* r);
/* returns -1 and leaves the state untouched if bound <= 0 (Java throws) */
int32_t vso_jr_next_int_bound(vso_jrandom* r, int32_t bound);
float vso_jr_next_float(vso_jrandom* r);
/* advance the generator by n draws of next() in O(log n) */
void vso_jr_skip(vso_jrandom* r, uint64_t n);

/* ---- SIMD lane model --------------------------------------------------------
 * FloatVector.SPECIES_PREFERRED lane count of the JVM being modelled
 * (16 = AVX-512, 8 = AVX2, 4 = NEON/SSE).  Default 16.  J/util/Distances.java:15 */
void vso_set_lanes(int lanes);
int vso_get_lanes(void);

/* ---- Distances (J/util/Distances.java) -------------------------------------- */
double vso_l2_squared(const float* a, const float* b, int len); /* :48-64 and :77-94 */
double vso_l2(const float* a, const float* b, int len);          /* :31-33 */
double vso_dot(const float* a, const float* b, int len);         /* :103-118 */
double vso_norm(const float* a, int len);                        /* :126-140 */
double vso_cosine(const float* a, const float* b, int len);      /* :149-153 */

/* ---- PQ (J/pq/PqEncoder.java, J/pq/PqTrainer.java) -------------------------- */
/* centroids: float[M][K][subDim] contiguous.  codes_out: uint8[M].  PqEncoder.java:18-37 */
void vso_pq_encode(const float* centroids, int M, int K, int subDim, const float* v,
                   uint8_t* codes_out);
void vso_pq_encode_batch(const float* centroids, int M, int K, int subDim, const float* rows,
                         int64_t n, uint8_t* codes_out, int threads);
/* PqTrainer.java:28-91.  rows: float[n][D].  centroids_out: float[M][K][D/M].
 * Returns 0, or -1 for the IllegalArgumentException cases (:29-34), or -2 for n == 0
 * (Java: IndexOutOfBoundsException from data.get(0) at :49).
 * draws_out (nullable) receives the number of Random draws consumed. */
int vso_pq_train(const float* rows, int64_t n, int D, int M, int K, int iterations, int64_t seed,
                 float* centroids_out, int64_t* draws_out);

/* The same two functions with the order-independent loops spread over `threads` host threads
 * (assignment per row; per-cluster sums owned by one thread each, rows still in ascending
 * order).  Bit-identical to the single-threaded restatements (checked in tests/); they exist so
 * that the BASELINE-size parity tests (10M x 128) finish in seconds. */
int vso_pq_train_mt(const float* rows, int64_t n, int D, int M, int K, int iterations, int64_t seed,
                    float* centroids_out, int64_t* draws_out, int threads);
void vso_pq_encode_batch_fast(const float* centroids, int M, int K, int subDim, const float* rows,
                              int64_t n, uint8_t* codes_out, int threads);

/* ---- ADC (J/fdb/FdbVectorIndex.java) ---------------------------------------- */
/* buildLut :1067-1079 -> double[M][K] */
void vso_build_lut(const float* centroids, int M, int K, int subDim, const float* q, double* lut);
/* pqApproxDistance :1057-1065 */
double vso_pq_approx_distance(const double* lut, const uint8_t* codes, int M, int K);
/* ADC scan + stable ascending sort + first n_cand :754-769,:820-822.
 * codes: uint8[n][M].  Returns the number of candidates written. */
int64_t vso_adc_topn(const double* lut, int M, int K, const uint8_t* codes, int64_t n,
                     int64_t n_cand, int64_t* ids_out, double* approx_out, int threads);

/* ---- exact scorers ----------------------------------------------------------- */
/* searchBruteForceSegment :676-721.  skip (nullable): non-zero = deleted or gid missing.
 * Stable sort by score descending (Double.compare), first k.  Returns count written. */
int64_t vso_bruteforce_topk(const float* rows, int64_t n, int d, const uint8_t* skip,
                            const float* q, int metric, int64_t k, int64_t* ids_out,
                            double* score_out, double* distance_out, int threads);
/* fetchExactAndScore :997-1043.  Candidates scored in the given order, ties keep that order.
 * cand ids < 0 or >= n are treated as a missing record (:1000 rec == null). */
int64_t vso_rerank_topk(const float* rows, int64_t n, int d, const uint8_t* skip, const float* q,
                        int metric, int normalize_on_read, const int64_t* cand, int64_t n_cand,
                        int64_t k, int64_t* ids_out, double* score_out, double* distance_out);
/* query merge :432-437.  Lists are concatenated in the order given, stably sorted by score
 * descending and truncated to k. */
int64_t vso_merge_topk(const int64_t* ids, const double* scores, int64_t total, int64_t k,
                       int64_t* ids_out, double* scores_out);

/* ---- GraphBuilder (J/graph/GraphBuilder.java:41-56, :73-109) ------------------ */
/* l_build <= 0: buildL2Neighbors(vectors, degree); else buildPrunedNeighbors(vectors, degree, l_build, alpha).
 * out: int32[n][degree] (-1 padded), counts: int32[n]. */
void vso_knn_graph(const float* rows, int64_t n, int d, int degree, int l_build, double alpha,
                   int32_t* out, int32_t* counts, int threads);

/* ---- FloatPacker (J/util/FloatPacker.java:21-39) ----------------------------- */
void vso_floats_to_bytes(const float* arr, int n, uint8_t* out);
void vso_bytes_to_floats(const uint8_t* bytes, int nbytes, float* out);

/* ---- synthetic inputs --------------------------------------------------------
 * Element e (0-based) of the stream is draw e of new Random(seed):
 *   kind 0: nextFloat()*2f-1f  (B/DistanceAndPqBenchmark.java:127-133)
 *   kind 1: nextFloat()        (:66-73)
 *   kind 2: nextFloat()*10f    (:79-85)
 * The slice [first, first+count) is produced by LCG skip-ahead. */
void vso_gen_floats(int64_t seed, int64_t first, int64_t count, int kind, float* out);
/* code byte e = (byte) nextInt(256) of draw e  (:86-89) */
void vso_gen_codes(int64_t seed, int64_t first, int64_t count, uint8_t* out);

/* JMH-like ns/op of the DistanceAndPqBenchmark bodies (config C1), kind 0 l2, 1 cosine, 2 pqEncode, 3 pqLutDistance */
double vso_bench_ns_per_op(int kind, const float* a, const float* b, int len, const float* centroids,
                           int M, int K, int subDim, const float* lut, const uint8_t* codes, int64_t iters);

#ifdef __cplusplus
}
#endif
#endif
