/*
 * vs_oracle.c -- CPU ORACLE (test infrastructure only; see vs_oracle.h).
 *
 * Plain-C restatement of the reference's Java arithmetic for the scoring hot path.
 * Compile with -ffp-contract=off: Java never contracts a*b+c into an FMA on its own,
 * the only fused operations are the explicit FloatVector.fma calls, restated with fmaf().
 *
 * Paths: J/ = /root/reference/src/main/java/io/github/panghy/vectorsearch/
 *        B/ = /root/reference/src/jmh/java/io/github/panghy/vectorsearch/bench/
 */
#include "vs_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* ========================================================================== */
/* java.util.Random -- JDK core class, fully specified 48-bit LCG.             */
/* ========================================================================== */
#define JR_MULT 0x5DEECE66DULL
#define JR_ADD 0xBULL
#define JR_MASK ((1ULL << 48) - 1)

void vso_jr_init(vso_jrandom* r, int64_t seed) {
  r->state = ((uint64_t)seed ^ JR_MULT) & JR_MASK; /* Random.initialScramble */
}

int32_t vso_jr_next(vso_jrandom* r, int bits) {
  r->state = (r->state * JR_MULT + JR_ADD) & JR_MASK;
  /* (int)(seed >>> (48 - bits)) : truncating cast of the 48-bit value */
  return (int32_t)(uint32_t)(r->state >> (48 - bits));
}

int32_t vso_jr_next_int(vso_jrandom* r) { return vso_jr_next(r, 32); }

int32_t vso_jr_next_int_bound(vso_jrandom* r, int32_t bound) {
  if (bound <= 0) return -1;
  int32_t x = vso_jr_next(r, 31);
  int32_t m = bound - 1;
  if ((bound & m) == 0) { /* power of two */
    x = (int32_t)(((int64_t)bound * (int64_t)x) >> 31);
  } else {
    /* for (int u = r; u - (r = u % bound) + m < 0; u = next(31)); -- int overflow wraps in Java */
    int32_t u = x;
    for (;;) {
      x = u % bound;
      int32_t t = (int32_t)((uint32_t)u - (uint32_t)x + (uint32_t)m);
      if (t >= 0) break;
      u = vso_jr_next(r, 31);
    }
  }
  return x;
}

float vso_jr_next_float(vso_jrandom* r) {
  /* next(24) / (float)(1 << 24); the product with 2^-24 is exact */
  return (float)vso_jr_next(r, 24) * (1.0f / 16777216.0f);
}

void vso_jr_skip(vso_jrandom* r, uint64_t n) {
  /* compose x -> a*x + c  n times by repeated squaring, mod 2^48 */
  uint64_t a = JR_MULT, c = JR_ADD;
  uint64_t acc_a = 1, acc_c = 0;
  while (n) {
    if (n & 1) {
      acc_a = (acc_a * a) & JR_MASK;
      acc_c = (acc_c * a + c) & JR_MASK;
    }
    c = ((a + 1) * c) & JR_MASK;
    a = (a * a) & JR_MASK;
    n >>= 1;
  }
  r->state = (acc_a * r->state + acc_c) & JR_MASK;
}

/* ========================================================================== */
/* Distances -- J/util/Distances.java                                          */
/* The vector loop keeps one fp32 accumulator per SIMD lane and updates it with */
/* a fused multiply-add (:52-57); reduceLanes(ADD) is modelled as an ordered    */
/* ascending-lane fp32 sum starting from 0.0f (HotSpot's strictly ordered       */
/* AddReductionVF on x86); the tail runs in double (:59-62).                    */
/* ========================================================================== */
static int g_lanes = 16;

void vso_set_lanes(int lanes) {
  if (lanes == 1 || lanes == 2 || lanes == 4 || lanes == 8 || lanes == 16) g_lanes = lanes;
}
int vso_get_lanes(void) { return g_lanes; }

static inline float reduce_lanes(const float* acc, int L) {
  float s = 0.0f;
  for (int l = 0; l < L; l++) s = s + acc[l];
  return s;
}

double vso_l2_squared(const float* a, const float* b, int len) {
  const int L = g_lanes;
  int i = 0;
  const int ub = len - (len % L); /* SPECIES.loopBound(len) */
  float acc[16] = {0};
  if (L == 16) {
    for (; i < ub; i += 16)
      for (int l = 0; l < 16; l++) {
        float diff = a[i + l] - b[i + l];
        acc[l] = fmaf(diff, diff, acc[l]);
      }
  } else {
    for (; i < ub; i += L)
      for (int l = 0; l < L; l++) {
        float diff = a[i + l] - b[i + l];
        acc[l] = fmaf(diff, diff, acc[l]);
      }
  }
  double sum = (double)reduce_lanes(acc, L);
  for (; i < len; i++) {
    double d = (double)a[i] - (double)b[i];
    sum += d * d;
  }
  return sum;
}

double vso_l2(const float* a, const float* b, int len) { return sqrt(vso_l2_squared(a, b, len)); }

double vso_dot(const float* a, const float* b, int len) {
  const int L = g_lanes;
  int i = 0;
  const int ub = len - (len % L);
  float acc[16] = {0};
  if (L == 16) {
    for (; i < ub; i += 16)
      for (int l = 0; l < 16; l++) acc[l] = fmaf(a[i + l], b[i + l], acc[l]);
  } else {
    for (; i < ub; i += L)
      for (int l = 0; l < L; l++) acc[l] = fmaf(a[i + l], b[i + l], acc[l]);
  }
  double s = (double)reduce_lanes(acc, L);
  for (; i < len; i++) s += (double)a[i] * (double)b[i];
  return s;
}

double vso_norm(const float* a, int len) {
  const int L = g_lanes;
  int i = 0;
  const int ub = len - (len % L);
  float acc[16] = {0};
  if (L == 16) {
    for (; i < ub; i += 16)
      for (int l = 0; l < 16; l++) acc[l] = fmaf(a[i + l], a[i + l], acc[l]);
  } else {
    for (; i < ub; i += L)
      for (int l = 0; l < L; l++) acc[l] = fmaf(a[i + l], a[i + l], acc[l]);
  }
  double s = (double)reduce_lanes(acc, L);
  for (; i < len; i++) s += (double)a[i] * (double)a[i];
  return sqrt(s);
}

double vso_cosine(const float* a, const float* b, int len) {
  double n = vso_norm(a, len) * vso_norm(b, len);
  if (n == 0.0) return 0.0;
  return vso_dot(a, b, len) / n;
}

/* ========================================================================== */
/* PqEncoder.encode -- J/pq/PqEncoder.java:18-37                               */
/* ========================================================================== */
void vso_pq_encode(const float* centroids, int M, int K, int subDim, const float* v,
                   uint8_t* codes_out) {
  for (int s = 0; s < M; s++) {
    const int off = s * subDim;
    int best = 0;
    double bestDist = INFINITY;
    for (int ci = 0; ci < K; ci++) {
      double d = vso_l2_squared(v + off, centroids + ((size_t)s * K + ci) * subDim, subDim);
      if (d < bestDist) { /* strict <: lowest ci wins ties, NaN never wins */
        bestDist = d;
        best = ci;
      }
    }
    codes_out[s] = (uint8_t)(best & 0xFF);
  }
}

void vso_pq_encode_batch(const float* centroids, int M, int K, int subDim, const float* rows,
                         int64_t n, uint8_t* codes_out, int threads) {
  const int D = M * subDim;
  (void)threads;
#pragma omp parallel for schedule(static) num_threads(threads > 0 ? threads : 1)
  for (int64_t i = 0; i < n; i++)
    vso_pq_encode(centroids, M, K, subDim, rows + (size_t)i * D, codes_out + (size_t)i * M);
}

/* ========================================================================== */
/* PqTrainer.train -- J/pq/PqTrainer.java:28-91                                */
/* ========================================================================== */
int vso_pq_train(const float* rows, int64_t n, int D, int M, int K, int iterations, int64_t seed,
                 float* centroids_out, int64_t* draws_out) {
  if (M <= 0 || K <= 0 || D <= 0) return -1; /* :29-31 */
  if (D % M != 0) return -1;                 /* :32-34 */
  if (n <= 0) return -2;                     /* data.get(0) on an empty list throws (:49) */
  if (n > 0x7fffffffLL) return -2;           /* a Java List cannot hold more */
  const int subDim = D / M;
  vso_jrandom rnd;
  vso_jr_init(&rnd, seed); /* :37 one generator shared by all subspaces */
  int64_t draws = 0;

  float* data = (float*)malloc((size_t)n * subDim * sizeof(float));
  int32_t* assign = (int32_t*)malloc((size_t)n * sizeof(int32_t));
  float* newC = (float*)malloc((size_t)K * subDim * sizeof(float));
  int32_t* counts = (int32_t*)malloc((size_t)K * sizeof(int32_t));

  for (int s = 0; s < M; s++) {
    float* C = centroids_out + (size_t)s * K * subDim;
    for (int64_t i = 0; i < n; i++) /* :41-45 sub-vector copies */
      memcpy(data + (size_t)i * subDim, rows + (size_t)i * D + (size_t)s * subDim,
             (size_t)subDim * sizeof(float));
    for (int ci = 0; ci < K; ci++) { /* :47-50 sampling WITH replacement */
      int32_t idx = vso_jr_next_int_bound(&rnd, (int32_t)n);
      draws++;
      memcpy(C + (size_t)ci * subDim, data + (size_t)idx * subDim, (size_t)subDim * sizeof(float));
    }
    for (int it = 0; it < iterations; it++) {
      for (int64_t i = 0; i < n; i++) { /* :56-68 assignment */
        const float* x = data + (size_t)i * subDim;
        int best = 0;
        double bestDist = INFINITY;
        for (int ci = 0; ci < K; ci++) {
          double d = vso_l2_squared(x, C + (size_t)ci * subDim, subDim);
          if (d < bestDist) {
            bestDist = d;
            best = ci;
          }
        }
        assign[i] = best;
      }
      memset(newC, 0, (size_t)K * subDim * sizeof(float)); /* :70-77 update, fp32 in row order */
      memset(counts, 0, (size_t)K * sizeof(int32_t));
      for (int64_t i = 0; i < n; i++) {
        const int a = assign[i];
        const float* x = data + (size_t)i * subDim;
        float* c = newC + (size_t)a * subDim;
        for (int d = 0; d < subDim; d++) c[d] += x[d];
        counts[a]++;
      }
      for (int ci = 0; ci < K; ci++) { /* :78-86 */
        float* c = newC + (size_t)ci * subDim;
        if (counts[ci] == 0) {
          int32_t idx = vso_jr_next_int_bound(&rnd, (int32_t)n);
          draws++;
          memcpy(c, data + (size_t)idx * subDim, (size_t)subDim * sizeof(float));
        } else {
          for (int d = 0; d < subDim; d++) c[d] /= (float)counts[ci];
        }
      }
      memcpy(C, newC, (size_t)K * subDim * sizeof(float)); /* :87 */
    }
  }
  free(data);
  free(assign);
  free(newC);
  free(counts);
  if (draws_out) *draws_out = draws;
  return 0;
}

/* ---- the same trainer with the order-independent parts spread over host threads ----------
 * Used by the BASELINE-size parity tests (C3: 10M x 128, 5 iterations = 2e11 sub-distances).
 * Bit-identical to vso_pq_train by construction, and checked against it in tests/:
 *   - assignment (:56-68): rows are independent; each (row, centroid) distance is the same
 *     sequence of operations.  For sub-vectors shorter than the SIMD register the vector loop
 *     of Distances.l2Squared runs zero times (:79 loopBound == 0) and the whole distance is the
 *     scalar double tail: d = (double)a - (double)b; sum += d * d (:88-91).  That case is laid
 *     out centroid-major here so that the compiler can evaluate 8 centroids per instruction --
 *     per (row, centroid) still mul-then-add in double, ascending component order.
 *   - update (:70-77): every (cluster, component) sum must add its rows in ascending row order.
 *     Thread t owns the clusters with ci % T == t and walks ALL rows in order, so each chain
 *     sees exactly the reference's sequence of fp32 additions. */
static void assign_rows_mt(const float* data, int64_t n, int subDim, const float* C, int K,
                           int32_t* assign, int threads) {
  const int L = g_lanes;
  if (subDim < L && subDim <= 64) {
    /* centroid-major copy in double: Ct[j][ci] = (double)C[ci][j] (the widening is exact) */
    double* Ct = (double*)malloc((size_t)subDim * K * sizeof(double));
    for (int ci = 0; ci < K; ci++)
      for (int j = 0; j < subDim; j++) Ct[(size_t)j * K + ci] = (double)C[(size_t)ci * subDim + j];
#pragma omp parallel num_threads(threads > 0 ? threads : 1)
    {
      double* dist = (double*)malloc((size_t)K * sizeof(double));
#pragma omp for schedule(static)
      for (int64_t i = 0; i < n; i++) {
        const float* x = data + (size_t)i * subDim;
        /* sum = (double)reduceLanes(zero lanes) = 0.0; then sum += d*d for j ascending */
        for (int ci = 0; ci < K; ci++) dist[ci] = 0.0;
        for (int j = 0; j < subDim; j++) {
          const double xj = (double)x[j];
          const double* cj = Ct + (size_t)j * K;
#pragma omp simd
          for (int ci = 0; ci < K; ci++) {
            const double d = xj - cj[ci];
            dist[ci] = dist[ci] + d * d;
          }
        }
        int best = 0;
        double bestDist = INFINITY;
        for (int ci = 0; ci < K; ci++)
          if (dist[ci] < bestDist) { /* strict <: lowest ci wins ties, NaN never wins */
            bestDist = dist[ci];
            best = ci;
          }
        assign[i] = best;
      }
      free(dist);
    }
    free(Ct);
    return;
  }
#pragma omp parallel for schedule(static) num_threads(threads > 0 ? threads : 1)
  for (int64_t i = 0; i < n; i++) {
    const float* x = data + (size_t)i * subDim;
    int best = 0;
    double bestDist = INFINITY;
    for (int ci = 0; ci < K; ci++) {
      double d = vso_l2_squared(x, C + (size_t)ci * subDim, subDim);
      if (d < bestDist) {
        bestDist = d;
        best = ci;
      }
    }
    assign[i] = best;
  }
}

int vso_pq_train_mt(const float* rows, int64_t n, int D, int M, int K, int iterations, int64_t seed,
                    float* centroids_out, int64_t* draws_out, int threads) {
  if (M <= 0 || K <= 0 || D <= 0) return -1;
  if (D % M != 0) return -1;
  if (n <= 0) return -2;
  if (n > 0x7fffffffLL) return -2;
  const int T = threads > 0 ? threads : 1;
  const int subDim = D / M;
  vso_jrandom rnd;
  vso_jr_init(&rnd, seed);
  int64_t draws = 0;
  float* data = (float*)malloc((size_t)n * subDim * sizeof(float));
  int32_t* assign = (int32_t*)malloc((size_t)n * sizeof(int32_t));
  float* newC = (float*)malloc((size_t)K * subDim * sizeof(float));
  int32_t* counts = (int32_t*)malloc((size_t)K * sizeof(int32_t));
  for (int s = 0; s < M; s++) {
    float* C = centroids_out + (size_t)s * K * subDim;
#pragma omp parallel for schedule(static) num_threads(T)
    for (int64_t i = 0; i < n; i++)
      memcpy(data + (size_t)i * subDim, rows + (size_t)i * D + (size_t)s * subDim,
             (size_t)subDim * sizeof(float));
    for (int ci = 0; ci < K; ci++) {
      int32_t idx = vso_jr_next_int_bound(&rnd, (int32_t)n);
      draws++;
      memcpy(C + (size_t)ci * subDim, data + (size_t)idx * subDim, (size_t)subDim * sizeof(float));
    }
    for (int it = 0; it < iterations; it++) {
      assign_rows_mt(data, n, subDim, C, K, assign, T);
      memset(newC, 0, (size_t)K * subDim * sizeof(float));
      memset(counts, 0, (size_t)K * sizeof(int32_t));
#pragma omp parallel num_threads(T)
      {
#ifdef _OPENMP
        const int t = omp_get_thread_num(), nt = omp_get_num_threads();
#else
        const int t = 0, nt = 1;
#endif
        for (int64_t i = 0; i < n; i++) {
          const int a = assign[i];
          if (a % nt != t) continue;
          const float* x = data + (size_t)i * subDim;
          float* c = newC + (size_t)a * subDim;
          for (int d = 0; d < subDim; d++) c[d] += x[d];
          counts[a]++;
        }
      }
      for (int ci = 0; ci < K; ci++) {
        float* c = newC + (size_t)ci * subDim;
        if (counts[ci] == 0) {
          int32_t idx = vso_jr_next_int_bound(&rnd, (int32_t)n);
          draws++;
          memcpy(c, data + (size_t)idx * subDim, (size_t)subDim * sizeof(float));
        } else {
          for (int d = 0; d < subDim; d++) c[d] /= (float)counts[ci];
        }
      }
      memcpy(C, newC, (size_t)K * subDim * sizeof(float));
    }
  }
  free(data);
  free(assign);
  free(newC);
  free(counts);
  if (draws_out) *draws_out = draws;
  return 0;
}

/* PqEncoder.encode over n rows with the same centroid-major evaluation (BASELINE-size checks) */
void vso_pq_encode_batch_fast(const float* centroids, int M, int K, int subDim, const float* rows,
                              int64_t n, uint8_t* codes_out, int threads) {
  const int D = M * subDim;
  const int T = threads > 0 ? threads : 1;
  float* data = (float*)malloc((size_t)n * subDim * sizeof(float));
  int32_t* assign = (int32_t*)malloc((size_t)n * sizeof(int32_t));
  for (int s = 0; s < M; s++) {
#pragma omp parallel for schedule(static) num_threads(T)
    for (int64_t i = 0; i < n; i++)
      memcpy(data + (size_t)i * subDim, rows + (size_t)i * D + (size_t)s * subDim,
             (size_t)subDim * sizeof(float));
    assign_rows_mt(data, n, subDim, centroids + (size_t)s * K * subDim, K, assign, T);
#pragma omp parallel for schedule(static) num_threads(T)
    for (int64_t i = 0; i < n; i++) codes_out[(size_t)i * M + s] = (uint8_t)(assign[i] & 0xFF);
  }
  free(data);
  free(assign);
}

/* ========================================================================== */
/* ADC -- J/fdb/FdbVectorIndex.java:1057-1079, :754-769                        */
/* ========================================================================== */
void vso_build_lut(const float* centroids, int M, int K, int subDim, const float* q, double* lut) {
  for (int s = 0; s < M; s++) {
    const int off = s * subDim;
    for (int ci = 0; ci < K; ci++)
      lut[(size_t)s * K + ci] =
          vso_l2_squared(q + off, centroids + ((size_t)s * K + ci) * subDim, subDim);
  }
}

double vso_pq_approx_distance(const double* lut, const uint8_t* codes, int M, int K) {
  double ad = 0.0;
  for (int s = 0; s < M; s++) {
    int ci = codes[s] & 0xFF;
    if (ci >= K) continue; /* :1061 */
    ad += lut[(size_t)s * K + ci];
  }
  return ad;
}

/* ---- Double.compare and a stable "first k of a sorted list" selector ------- */
static inline int64_t dbits(double d) {
  int64_t b;
  if (d != d) return 0x7ff8000000000000LL; /* doubleToLongBits canonicalises NaN */
  memcpy(&b, &d, 8);
  return b;
}
static inline int jcmp(double a, double b) { /* java.lang.Double.compare */
  if (a < b) return -1;
  if (a > b) return 1;
  int64_t x = dbits(a), y = dbits(b);
  return x == y ? 0 : (x < y ? -1 : 1);
}

typedef struct {
  double key;
  int64_t id;
} vso_ent;

typedef struct {
  vso_ent* e;
  int64_t k, cnt;
  int desc;
} vso_sel;

/* "a sorts strictly before b" under the comparator; equal keys keep arrival order */
static inline int sel_before(const vso_sel* s, double a, double b) {
  return s->desc ? (jcmp(b, a) < 0) : (jcmp(a, b) < 0);
}
static void sel_init(vso_sel* s, int64_t k, int desc) {
  s->e = (vso_ent*)malloc((size_t)(k > 0 ? k : 1) * sizeof(vso_ent));
  s->k = k;
  s->cnt = 0;
  s->desc = desc;
}
/* offer entries in list order: result == stable sort then subList(0,k) */
static inline void sel_offer(vso_sel* s, double key, int64_t id) {
  if (s->k <= 0) return;
  if (s->cnt == s->k && !sel_before(s, key, s->e[s->cnt - 1].key)) return;
  int64_t pos = s->cnt < s->k ? s->cnt : s->k - 1;
  while (pos > 0 && sel_before(s, key, s->e[pos - 1].key)) {
    s->e[pos] = s->e[pos - 1];
    pos--;
  }
  s->e[pos].key = key;
  s->e[pos].id = id;
  if (s->cnt < s->k) s->cnt++;
}

static int resolve_threads(int threads) {
#ifdef _OPENMP
  if (threads <= 0) return 1;
  return threads;
#else
  (void)threads;
  return 1;
#endif
}

int64_t vso_adc_topn(const double* lut, int M, int K, const uint8_t* codes, int64_t n,
                     int64_t n_cand, int64_t* ids_out, double* approx_out, int threads) {
  const int T = resolve_threads(threads);
  vso_sel* sels = (vso_sel*)malloc((size_t)T * sizeof(vso_sel));
  for (int t = 0; t < T; t++) sel_init(&sels[t], n_cand, 0);
#pragma omp parallel num_threads(T)
  {
#ifdef _OPENMP
    const int t = omp_get_thread_num();
#else
    const int t = 0;
#endif
    const int64_t lo = n * t / T, hi = n * (t + 1) / T;
    for (int64_t i = lo; i < hi; i++)
      sel_offer(&sels[t], vso_pq_approx_distance(lut, codes + (size_t)i * M, M, K), i);
  }
  vso_sel fin;
  sel_init(&fin, n_cand, 0);
  for (int t = 0; t < T; t++) { /* chunks are ascending row ranges: merge keeps stability */
    for (int64_t j = 0; j < sels[t].cnt; j++) sel_offer(&fin, sels[t].e[j].key, sels[t].e[j].id);
    free(sels[t].e);
  }
  free(sels);
  for (int64_t j = 0; j < fin.cnt; j++) {
    ids_out[j] = fin.e[j].id;
    approx_out[j] = fin.e[j].key;
  }
  int64_t c = fin.cnt;
  free(fin.e);
  return c;
}

/* ========================================================================== */
/* exact scorers                                                               */
/* ========================================================================== */
static inline void score_pair(const float* q, const float* emb, int d, int metric,
                              int normalize_on_read, double qNorm, double* score,
                              double* distance) {
  if (metric == VSO_METRIC_COSINE) {
    double sim;
    if (normalize_on_read) { /* J/fdb/FdbVectorIndex.java:1006-1010 */
      double denom = qNorm == 0.0 ? vso_norm(q, d) * vso_norm(emb, d) : qNorm * vso_norm(emb, d);
      sim = denom == 0.0 ? 0.0 : vso_dot(q, emb, d) / denom;
    } else {
      sim = vso_cosine(q, emb, d); /* :687 / :1012 */
    }
    *score = sim;
    *distance = 1.0 - sim;
  } else {
    double dist = vso_l2(q, emb, d); /* :691 / :1017 */
    *score = -dist;
    *distance = dist;
  }
}

int64_t vso_bruteforce_topk(const float* rows, int64_t n, int d, const uint8_t* skip,
                            const float* q, int metric, int64_t k, int64_t* ids_out,
                            double* score_out, double* distance_out, int threads) {
  const int T = resolve_threads(threads);
  vso_sel* sels = (vso_sel*)malloc((size_t)T * sizeof(vso_sel));
  for (int t = 0; t < T; t++) sel_init(&sels[t], k, 1);
#pragma omp parallel num_threads(T)
  {
#ifdef _OPENMP
    const int t = omp_get_thread_num();
#else
    const int t = 0;
#endif
    const int64_t lo = n * t / T, hi = n * (t + 1) / T;
    for (int64_t i = lo; i < hi; i++) {
      if (skip && skip[i]) continue; /* :681 deleted, :696 gid missing */
      double sc, di;
      score_pair(q, rows + (size_t)i * d, d, metric, 0, 0.0, &sc, &di);
      sel_offer(&sels[t], sc, i);
    }
  }
  vso_sel fin;
  sel_init(&fin, k, 1);
  for (int t = 0; t < T; t++) {
    for (int64_t j = 0; j < sels[t].cnt; j++) sel_offer(&fin, sels[t].e[j].key, sels[t].e[j].id);
    free(sels[t].e);
  }
  free(sels);
  for (int64_t j = 0; j < fin.cnt; j++) {
    ids_out[j] = fin.e[j].id;
    score_out[j] = fin.e[j].key;
    if (distance_out)
      distance_out[j] = (metric == VSO_METRIC_COSINE) ? 1.0 - fin.e[j].key : -fin.e[j].key;
  }
  int64_t c = fin.cnt;
  free(fin.e);
  return c;
}

int64_t vso_rerank_topk(const float* rows, int64_t n, int d, const uint8_t* skip, const float* q,
                        int metric, int normalize_on_read, const int64_t* cand, int64_t n_cand,
                        int64_t k, int64_t* ids_out, double* score_out, double* distance_out) {
  /* :823-826 qNorm is only precomputed for COSINE && normalizeOnRead */
  const double qNorm =
      (metric == VSO_METRIC_COSINE && normalize_on_read) ? vso_norm(q, d) : 0.0;
  vso_sel fin;
  sel_init(&fin, k, 1);
  for (int64_t j = 0; j < n_cand; j++) {
    const int64_t id = cand[j];
    if (id < 0 || id >= n) continue;  /* rec == null */
    if (skip && skip[id]) continue;   /* deleted / gid missing */
    double sc, di;
    score_pair(q, rows + (size_t)id * d, d, metric, normalize_on_read, qNorm, &sc, &di);
    sel_offer(&fin, sc, id);
  }
  for (int64_t j = 0; j < fin.cnt; j++) {
    ids_out[j] = fin.e[j].id;
    score_out[j] = fin.e[j].key;
    if (distance_out)
      distance_out[j] = (metric == VSO_METRIC_COSINE) ? 1.0 - fin.e[j].key : -fin.e[j].key;
  }
  int64_t c = fin.cnt;
  free(fin.e);
  return c;
}

int64_t vso_merge_topk(const int64_t* ids, const double* scores, int64_t total, int64_t k,
                       int64_t* ids_out, double* scores_out) {
  vso_sel fin;
  sel_init(&fin, k, 1);
  for (int64_t j = 0; j < total; j++) sel_offer(&fin, scores[j], ids[j]);
  for (int64_t j = 0; j < fin.cnt; j++) {
    ids_out[j] = fin.e[j].id;
    scores_out[j] = fin.e[j].key;
  }
  int64_t c = fin.cnt;
  free(fin.e);
  return c;
}

/* ========================================================================== */
/* GraphBuilder -- J/graph/GraphBuilder.java:41-56 (buildL2Neighbors) and      */
/* :73-109 (buildPrunedNeighbors).  Arrays.sort on Integer[] is a stable merge */
/* sort, so equal l2Squared keeps ascending j; comparingDouble = Double.compare */
/* ========================================================================== */
typedef struct {
  double d;
  int32_t j;
} vso_dj;
static int dj_cmp(const void* a, const void* b) {
  const vso_dj* x = (const vso_dj*)a;
  const vso_dj* y = (const vso_dj*)b;
  int c = jcmp(x->d, y->d);
  if (c) return c;
  return x->j < y->j ? -1 : (x->j > y->j ? 1 : 0);
}

/* l_build <= 0: buildL2Neighbors(vectors, degree); else buildPrunedNeighbors(vectors, degree, l_build, alpha).
 * out: int32[n][degree], -1 padded; counts: int32[n]. */
void vso_knn_graph(const float* rows, int64_t n, int d, int degree, int l_build, double alpha,
                   int32_t* out, int32_t* counts, int threads) {
  const int T = resolve_threads(threads);
  const int prune_mode = l_build > 0;
  const int prune = alpha > 1.0; /* :79 */
#pragma omp parallel num_threads(T)
  {
    vso_dj* dj = (vso_dj*)malloc((size_t)(n > 1 ? n - 1 : 1) * sizeof(vso_dj));
    int32_t* selected = (int32_t*)malloc((size_t)(degree > 0 ? degree : 1) * sizeof(int32_t));
#pragma omp for schedule(dynamic, 16)
    for (int64_t i = 0; i < n; i++) {
      int64_t p = 0;
      for (int64_t j = 0; j < n; j++)
        if (j != i) {
          dj[p].d = vso_l2_squared(rows + (size_t)i * d, rows + (size_t)j * d, d); /* :50 / :84 */
          dj[p].j = (int32_t)j;
          p++;
        }
      qsort(dj, (size_t)p, sizeof(vso_dj), dj_cmp);
      int32_t* o = out + (size_t)i * degree;
      for (int c = 0; c < degree; c++) o[c] = -1;
      if (!prune_mode) {
        int64_t take = degree < n - 1 ? degree : n - 1; /* :51 */
        if (take < 0) take = 0;
        for (int64_t k = 0; k < take; k++) o[k] = dj[k].j;
        counts[i] = (int32_t)take;
      } else {
        int64_t limit = l_build < n - 1 ? l_build : n - 1; /* :88 */
        if (limit < 0) limit = 0;
        const int64_t cap = degree < limit ? degree : limit; /* :89 */
        int64_t sN = 0;
        for (int64_t k = 0; k < limit && sN < cap; k++) {
          const int32_t u = dj[k].j;
          int keep = 1;
          if (prune) {
            const double diu = dj[k].d;
            for (int64_t t = 0; t < sN; t++) {
              const double dup = vso_l2_squared(rows + (size_t)u * d, rows + (size_t)selected[t] * d, d);
              if (dup <= alpha * diu) { /* :101 */
                keep = 0;
                break;
              }
            }
          }
          if (keep) selected[sN++] = u;
        }
        for (int64_t k = 0; k < sN; k++) o[k] = selected[k];
        counts[i] = (int32_t)sN;
      }
    }
    free(dj);
    free(selected);
  }
}

/* ========================================================================== */
/* FloatPacker -- J/util/FloatPacker.java:21-39 (little-endian fp32)           */
/* ========================================================================== */
void vso_floats_to_bytes(const float* arr, int n, uint8_t* out) {
  for (int i = 0; i < n; i++) {
    uint32_t u;
    memcpy(&u, &arr[i], 4);
    out[4 * i + 0] = (uint8_t)(u & 0xFF);
    out[4 * i + 1] = (uint8_t)((u >> 8) & 0xFF);
    out[4 * i + 2] = (uint8_t)((u >> 16) & 0xFF);
    out[4 * i + 3] = (uint8_t)((u >> 24) & 0xFF);
  }
}

void vso_bytes_to_floats(const uint8_t* bytes, int nbytes, float* out) {
  const int n = nbytes / 4;
  for (int i = 0; i < n; i++) {
    uint32_t u = (uint32_t)bytes[4 * i] | ((uint32_t)bytes[4 * i + 1] << 8) |
                 ((uint32_t)bytes[4 * i + 2] << 16) | ((uint32_t)bytes[4 * i + 3] << 24);
    memcpy(&out[i], &u, 4);
  }
}

/* ========================================================================== */
/* synthetic inputs -- B/DistanceAndPqBenchmark.java:39-90,127-133             */
/* ========================================================================== */
void vso_gen_floats(int64_t seed, int64_t first, int64_t count, int kind, float* out) {
  const int T = 8;
#pragma omp parallel for schedule(static) num_threads(T)
  for (int t = 0; t < T; t++) {
    const int64_t lo = count * t / T, hi = count * (t + 1) / T;
    vso_jrandom r;
    vso_jr_init(&r, seed);
    vso_jr_skip(&r, (uint64_t)(first + lo));
    for (int64_t i = lo; i < hi; i++) {
      float f = vso_jr_next_float(&r);
      if (kind == 0)
        f = f * 2.0f - 1.0f;
      else if (kind == 2)
        f = f * 10.0f;
      out[i] = f;
    }
  }
}

void vso_gen_codes(int64_t seed, int64_t first, int64_t count, uint8_t* out) {
  vso_jrandom r;
  vso_jr_init(&r, seed);
  vso_jr_skip(&r, (uint64_t)first);
  for (int64_t i = 0; i < count; i++) out[i] = (uint8_t)vso_jr_next_int_bound(&r, 256);
}


/* ========================================================================== */
/* JMH-like timing of the DistanceAndPqBenchmark bodies (config C1): the        */
/* benchmark method is called `iters` times in a C loop (B/DistanceAndPqBenchmark.java:95-123); */
/* returns nanoseconds per call.  kind 0: l2, 1: cosine, 2: pqEncode, 3: pqLutDistance (float LUT). */
/* ========================================================================== */
#include <time.h>
double vso_bench_ns_per_op(int kind, const float* a, const float* b, int len, const float* centroids,
                           int M, int K, int subDim, const float* lut, const uint8_t* codes, int64_t iters) {
  volatile double sink = 0.0;
  uint8_t out[256];
  struct timespec t0, t1;
  clock_gettime(CLOCK_MONOTONIC, &t0);
  for (int64_t i = 0; i < iters; i++) {
    switch (kind) {
      case 0: sink += vso_l2(a, b, len); break;
      case 1: sink += vso_cosine(a, b, len); break;
      case 2:
        vso_pq_encode(centroids, M, K, subDim, a, out);
        sink += out[0];
        break;
      default: {
        float dist = 0.0f; /* :117-122 */
        for (int m = 0; m < M; m++) dist += lut[(size_t)m * K + (codes[m] & 0xFF)];
        sink += dist;
      }
    }
    __asm__ volatile("" ::: "memory");
  }
  clock_gettime(CLOCK_MONOTONIC, &t1);
  (void)sink;
  return ((double)(t1.tv_sec - t0.tv_sec) * 1e9 + (double)(t1.tv_nsec - t0.tv_nsec)) / (double)(iters > 0 ? iters : 1);
}
