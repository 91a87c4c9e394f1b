package io.github.panghy.vectorsearch.gpu;

import static io.github.panghy.vectorsearch.gpu.VsGpu.A;
import static io.github.panghy.vectorsearch.gpu.VsGpu.I;
import static io.github.panghy.vectorsearch.gpu.VsGpu.J;
import static io.github.panghy.vectorsearch.gpu.VsGpu.call;
import static io.github.panghy.vectorsearch.gpu.VsGpu.check;
import static io.github.panghy.vectorsearch.gpu.VsGpu.sig;
import static java.lang.foreign.ValueLayout.JAVA_BYTE;
import static java.lang.foreign.ValueLayout.JAVA_DOUBLE;
import static java.lang.foreign.ValueLayout.JAVA_FLOAT;
import static java.lang.foreign.ValueLayout.JAVA_INT;
import static java.lang.foreign.ValueLayout.JAVA_LONG;

import java.lang.foreign.Arena;
import java.lang.foreign.MemorySegment;
import java.util.List;

/**
 * Facade over libvsgpu for the four call sites of the scoring path:
 *
 * <ul>
 *   <li>{@code FdbVectorIndex.searchBruteForceSegment} (J/fdb/FdbVectorIndex.java:676-721) -> {@link #bruteForceTopK}</li>
 *   <li>{@code searchSealedSegment} ADC scan + {@code fetchExactAndScore} (:754-769, :820-828, :997-1043) ->
 *       {@link #adcRerankTopK}, {@link #adcTopK}, {@link #rerankTopK}; BEST_FIRST expansion (:950-963) -> {@link AdcQuery}</li>
 *   <li>{@code SegmentBuildService} training and encoding (J/tasks/SegmentBuildService.java:180, :301) ->
 *       {@link #trainPq}, {@link #attachPq}</li>
 *   <li>the merge of per-segment lists (:432-437) -> {@link #mergeTopK}</li>
 * </ul>
 *
 * Segments are resident in HBM behind long handles; rows are row indices into the uploaded order (the caller keeps
 * the row -> vecId -> gid mapping it already has, {@link #uploadRecords} returns the vec_ids in row order). Scores are
 * the reference's doubles: {@code -l2} for L2, the similarity for COSINE (:687-693); distance = -score resp. 1 - score.
 * With {@link #init(int...)} given several devices the same calls shard every segment over all of them.
 */
public final class GpuScoring {
  private GpuScoring() {}

  /** One hit: row = id_base + row index of the segment. */
  public record Hit(long row, double score) {}

  /** SegmentMeta.State (vectorsearch.proto:84). */
  public static final int ACTIVE = 0, PENDING = 1, SEALED = 2, COMPACTING = 3, WRITING = 4;

  // ---- lifecycle ----------------------------------------------------------------------------------------------
  public static void init(int... devices) {
    if (devices.length == 0) devices = new int[] {0};
    try (Arena a = Arena.ofConfined()) {
      MemorySegment d = a.allocateFrom(JAVA_INT, devices);
      check(call("vs_init_multi", sig(I, A), devices.length, d));
    }
  }

  public static void shutdown() {
    check(call("vs_shutdown", sig()));
  }

  /** Lane count of this JVM's FloatVector.SPECIES_PREFERRED (16 AVX-512, 8 AVX2, 4 NEON): low-order bits of scores. */
  public static void setSimdLanes(int lanes) {
    check(call("vs_set_simd_lanes", sig(I), lanes));
  }

  // ---- residency ---------------------------------------------------------------------------------------------
  /** Rows as float[n][d] (copied once into an off-heap buffer, then one H2D copy). */
  public static long uploadSegment(float[][] rows, byte[] skipMask, long idBase) {
    int n = rows.length, d = n == 0 ? 1 : rows[0].length;
    try (Arena a = Arena.ofConfined()) {
      MemorySegment buf = a.allocate(JAVA_FLOAT, Math.max(1L, (long) n * d));
      for (int i = 0; i < n; i++) MemorySegment.copy(rows[i], 0, buf, JAVA_FLOAT, (long) i * d * 4, d);
      MemorySegment skip = skipMask == null ? MemorySegment.NULL : a.allocateFrom(JAVA_BYTE, skipMask);
      MemorySegment h = a.allocate(JAVA_LONG);
      check(call("vs_segment_upload", sig(A, J, I, A, J, A), buf, (long) n, d, skip, idBase, h));
      return h.get(JAVA_LONG, 0);
    }
  }

  /**
   * The stored VectorRecord values of a segment as the range read returns them (KeyValue.getValue()): embeddings are
   * consumed as bytes, {@code deleted} becomes the skip flag. Returns the handle; vecIdsOut (length n) receives vec_ids.
   */
  public static long uploadRecords(List<byte[]> records, int dimension, long idBase, int[] vecIdsOut) {
    int n = records.size();
    long total = 0;
    for (byte[] r : records) total += r.length;
    try (Arena a = Arena.ofConfined()) {
      MemorySegment buf = a.allocate(Math.max(1L, total));
      MemorySegment offs = a.allocate(JAVA_LONG, n + 1L);
      long o = 0;
      for (int i = 0; i < n; i++) {
        byte[] r = records.get(i);
        offs.setAtIndex(JAVA_LONG, i, o);
        MemorySegment.copy(r, 0, buf, JAVA_BYTE, o, r.length);
        o += r.length;
      }
      offs.setAtIndex(JAVA_LONG, n, o);
      MemorySegment ids = vecIdsOut == null ? MemorySegment.NULL : a.allocate(JAVA_INT, Math.max(1, n));
      MemorySegment h = a.allocate(JAVA_LONG);
      check(call("vs_segment_upload_records", sig(A, A, J, I, J, A, A), buf, offs, (long) n, dimension, idBase, ids, h));
      if (vecIdsOut != null) MemorySegment.copy(ids, JAVA_INT, 0, vecIdsOut, 0, n);
      return h.get(JAVA_LONG, 0);
    }
  }

  public static void setSkipMask(long handle, byte[] skipMask) {
    try (Arena a = Arena.ofConfined()) {
      MemorySegment skip = skipMask == null ? MemorySegment.NULL : a.allocateFrom(JAVA_BYTE, skipMask);
      check(call("vs_segment_set_skip", sig(J, A), handle, skip));
    }
  }

  /** Sealing: the stored PQCodebook message (SegmentBuildService.buildCodebookBytes) and, optionally, the stored codes. */
  public static void attachPq(long handle, byte[] codebookBytes, byte[] codesOrNull) {
    try (Arena a = Arena.ofConfined()) {
      MemorySegment cb = a.allocateFrom(JAVA_BYTE, codebookBytes);
      MemorySegment codes = codesOrNull == null ? MemorySegment.NULL : a.allocateFrom(JAVA_BYTE, codesOrNull);
      check(call("vs_segment_attach_pq_codebook", sig(J, A, J, A), handle, cb, (long) codebookBytes.length, codes));
    }
  }

  /** codes [n][m] of the resident rows (encoded on the device when attachPq was given none): what the build stores. */
  public static byte[] downloadCodes(long handle, long first, long count, int m) {
    try (Arena a = Arena.ofConfined()) {
      MemorySegment out = a.allocate(Math.max(1L, count * m));
      check(call("vs_segment_download_codes", sig(J, J, J, A), handle, first, count, out));
      return out.asSlice(0, count * m).toArray(JAVA_BYTE);
    }
  }

  public static void free(long handle) {
    check(call("vs_segment_free", sig(J), handle));
  }

  /** Residency table keyed by (segId, state): returns 0 when the segment is not resident in that state. */
  public static long resident(long segId, int state) {
    try (Arena a = Arena.ofConfined()) {
      MemorySegment h = a.allocate(JAVA_LONG);
      int rc = call("vs_residency_get", sig(J, I, A), segId, state, h);
      if (rc == VsGpu.VS_EHANDLE || rc == VsGpu.VS_ESTATE) return 0L;
      check(rc);
      return h.get(JAVA_LONG, 0);
    }
  }

  public static void markResident(long segId, int state, long handle) {
    check(call("vs_residency_put", sig(J, I, J), segId, state, handle));
  }

  /** Compaction rebuild / segment deletion (MaintenanceService.java:388-390). */
  public static void invalidate(long segId) {
    check(call("vs_residency_invalidate", sig(J), segId));
  }

  // ---- queries ----------------------------------------------------------------------------------------------
  private static Hit[] hits(MemorySegment ids, MemorySegment sc, int count) {
    Hit[] out = new Hit[count];
    for (int i = 0; i < count; i++) out[i] = new Hit(ids.getAtIndex(JAVA_LONG, i), sc.getAtIndex(JAVA_DOUBLE, i));
    return out;
  }

  /** searchBruteForceSegment: stable sort by score descending, first k (ties keep the lower row). */
  public static Hit[] bruteForceTopK(long handle, float[] q, int k, int metric) {
    try (Arena a = Arena.ofConfined()) {
      MemorySegment qs = a.allocateFrom(JAVA_FLOAT, q);
      MemorySegment ids = a.allocate(JAVA_LONG, k), sc = a.allocate(JAVA_DOUBLE, k), cn = a.allocate(JAVA_INT);
      check(call("vs_bruteforce_topk", sig(J, A, I, I, I, A, A, A), handle, qs, 1, k, metric, ids, sc, cn));
      return hits(ids, sc, cn.get(JAVA_INT, 0));
    }
  }

  /** A batch of queries against one segment in one call (tensor-core nomination, exact scores): result[i] for queries[i]. */
  public static Hit[][] bruteForceTopK(long handle, float[][] queries, int k, int metric) {
    int nq = queries.length, d = queries[0].length;
    try (Arena a = Arena.ofConfined()) {
      MemorySegment qs = a.allocate(JAVA_FLOAT, (long) nq * d);
      for (int i = 0; i < nq; i++) MemorySegment.copy(queries[i], 0, qs, JAVA_FLOAT, (long) i * d * 4, d);
      MemorySegment ids = a.allocate(JAVA_LONG, (long) nq * k), sc = a.allocate(JAVA_DOUBLE, (long) nq * k);
      MemorySegment cn = a.allocate(JAVA_INT, nq);
      check(call("vs_bruteforce_topk", sig(J, A, I, I, I, A, A, A), handle, qs, nq, k, metric, ids, sc, cn));
      Hit[][] out = new Hit[nq][];
      for (int i = 0; i < nq; i++)
        out[i] = hits(ids.asSlice((long) i * k * 8), sc.asSlice((long) i * k * 8), cn.getAtIndex(JAVA_INT, i));
      return out;
    }
  }

  /** buildLut + full ADC scan + first nCand of the ascending stable sort (:741, :754-769, :820-822); score = approx. */
  public static Hit[] adcTopK(long handle, float[] q, int nCand) {
    try (Arena a = Arena.ofConfined()) {
      MemorySegment qs = a.allocateFrom(JAVA_FLOAT, q);
      MemorySegment ids = a.allocate(JAVA_LONG, nCand), ap = a.allocate(JAVA_DOUBLE, nCand), cn = a.allocate(JAVA_INT);
      check(call("vs_adc_topk", sig(J, A, I, I, A, A, A), handle, qs, 1, nCand, ids, ap, cn));
      return hits(ids, ap, cn.get(JAVA_INT, 0));
    }
  }

  /** fetchExactAndScore over caller-supplied candidates: scored in the given order, ties keep it (:997-1043). */
  public static Hit[] rerankTopK(long handle, float[] q, long[] candidateRows, int k, int metric, boolean normalizeOnRead) {
    try (Arena a = Arena.ofConfined()) {
      MemorySegment qs = a.allocateFrom(JAVA_FLOAT, q);
      MemorySegment cand = a.allocateFrom(JAVA_LONG, candidateRows);
      MemorySegment ids = a.allocate(JAVA_LONG, k), sc = a.allocate(JAVA_DOUBLE, k), cn = a.allocate(JAVA_INT);
      check(call("vs_rerank_topk", sig(J, A, A, I, I, I, I, A, A, A), handle, qs, cand, candidateRows.length, k, metric,
          normalizeOnRead ? 1 : 0, ids, sc, cn));
      return hits(ids, sc, cn.get(JAVA_INT, 0));
    }
  }

  /** The sealed-segment query in one call: ADC top nCand, exact re-rank to k. */
  public static Hit[] adcRerankTopK(long handle, float[] q, int nCand, int k, int metric, boolean normalizeOnRead) {
    try (Arena a = Arena.ofConfined()) {
      MemorySegment qs = a.allocateFrom(JAVA_FLOAT, q);
      MemorySegment ids = a.allocate(JAVA_LONG, k), sc = a.allocate(JAVA_DOUBLE, k), cn = a.allocate(JAVA_INT);
      check(call("vs_adc_rerank_topk", sig(J, A, I, I, I, I, I, A, A, A), handle, qs, 1, nCand, k, metric,
          normalizeOnRead ? 1 : 0, ids, sc, cn));
      return hits(ids, sc, cn.get(JAVA_INT, 0));
    }
  }

  /** query() :432-437: per-segment lists concatenated in segment order, stable sort by score descending, first k. */
  public static Hit[] mergeTopK(long[] rows, double[] scores, int k) {
    try (Arena a = Arena.ofConfined()) {
      MemorySegment in = a.allocateFrom(JAVA_LONG, rows), is = a.allocateFrom(JAVA_DOUBLE, scores);
      MemorySegment ids = a.allocate(JAVA_LONG, k), sc = a.allocate(JAVA_DOUBLE, k), cn = a.allocate(JAVA_INT);
      check(call("vs_merge_topk", sig(A, A, J, I, A, A, A), in, is, (long) rows.length, k, ids, sc, cn));
      return hits(ids, sc, cn.get(JAVA_INT, 0));
    }
  }

  /**
   * BEST_FIRST expansion scoring (:741, :950-963): the LUT of one query against one sealed segment stays on the device;
   * each expansion step scores the frontier's neighbour rows against the RESIDENT codes (no codeMap, no range read).
   */
  public static final class AdcQuery implements AutoCloseable {
    private long handle;

    public AdcQuery(long segment, float[] q) {
      try (Arena a = Arena.ofConfined()) {
        MemorySegment qs = a.allocateFrom(JAVA_FLOAT, q), h = a.allocate(JAVA_LONG);
        check(call("vs_adc_query_begin", sig(J, A, A), segment, qs, h));
        handle = h.get(JAVA_LONG, 0);
      }
    }

    /** approx[i] = pqApproxDistance(lut, codes of rows[i]); NaN where the segment holds no code for that row (:957). */
    public double[] approx(long[] rows) {
      try (Arena a = Arena.ofConfined()) {
        MemorySegment in = a.allocateFrom(JAVA_LONG, rows);
        MemorySegment out = a.allocate(JAVA_DOUBLE, Math.max(1, rows.length));
        check(call("vs_adc_query_gather", sig(J, A, J, A, A), handle, in, (long) rows.length, out, MemorySegment.NULL));
        return out.asSlice(0, rows.length * 8L).toArray(JAVA_DOUBLE);
      }
    }

    @Override
    public void close() {
      if (handle != 0) check(call("vs_adc_query_end", sig(J), handle));
      handle = 0;
    }
  }

  // ---- build -------------------------------------------------------------------------------------------------------
  /** PqTrainer.train over the resident rows of a segment (SegmentBuildService.java:180 calls it with 5, 42L). */
  public static float[][][] trainPq(long handle, long n, int dimension, int m, int k, int iterations, long seed) {
    if (m <= 0 || k <= 0 || dimension <= 0) throw new IllegalArgumentException("Invalid PQ params (m,k,dimension)");
    if (dimension % m != 0) throw new IllegalArgumentException("dimension must be divisible by m");
    int sub = dimension / m;
    try (Arena a = Arena.ofConfined()) {
      MemorySegment out = a.allocate(JAVA_FLOAT, (long) m * k * sub);
      check(call("vs_pq_train", sig(A, J, J, I, I, I, I, J, A), MemorySegment.NULL, handle, n, dimension, m, k, iterations, seed, out));
      return unflatten(out, m, k, sub);
    }
  }

  /** float[M][K][subDim] -> the PQCodebook bytes the build stores (same bytes as buildCodebookBytes). */
  public static byte[] codebookBytes(float[][][] centroids) {
    int m = centroids.length, k = centroids[0].length, sub = centroids[0][0].length;
    try (Arena a = Arena.ofConfined()) {
      MemorySegment c = flatten(a, centroids), len = a.allocate(JAVA_LONG);
      check(call("vs_codebook_encode", sig(A, I, I, I, A, J, A), c, m, k, sub, MemorySegment.NULL, 0L, len));
      long n = len.get(JAVA_LONG, 0);
      MemorySegment out = a.allocate(n);
      check(call("vs_codebook_encode", sig(A, I, I, I, A, J, A), c, m, k, sub, out, n, len));
      return out.toArray(JAVA_BYTE);
    }
  }

  static MemorySegment flatten(Arena a, float[][][] c) {
    int m = c.length, k = c[0].length, sub = c[0][0].length;
    MemorySegment out = a.allocate(JAVA_FLOAT, (long) m * k * sub);
    for (int s = 0; s < m; s++)
      for (int ci = 0; ci < k; ci++) MemorySegment.copy(c[s][ci], 0, out, JAVA_FLOAT, ((long) s * k + ci) * sub * 4, sub);
    return out;
  }

  static float[][][] unflatten(MemorySegment flat, int m, int k, int sub) {
    float[][][] c = new float[m][k][sub];
    for (int s = 0; s < m; s++)
      for (int ci = 0; ci < k; ci++) MemorySegment.copy(flat, JAVA_FLOAT, ((long) s * k + ci) * sub * 4, c[s][ci], 0, sub);
    return c;
  }
}
