package io.github.panghy.vectorsearch.pq;

import static io.github.panghy.vectorsearch.gpu.VsGpu.call;
import static io.github.panghy.vectorsearch.gpu.VsGpu.check;
import static java.lang.foreign.ValueLayout.ADDRESS;
import static java.lang.foreign.ValueLayout.JAVA_BYTE;
import static java.lang.foreign.ValueLayout.JAVA_FLOAT;
import static java.lang.foreign.ValueLayout.JAVA_INT;
import static java.lang.foreign.ValueLayout.JAVA_LONG;

import java.lang.foreign.Arena;
import java.lang.foreign.MemoryLayout;
import java.lang.foreign.MemorySegment;
import java.util.List;

/**
 * Drop-in shim for J/pq/PqEncoder.java:18-37 (same static signature), plus the batch form the build loop wants
 * (SegmentBuildService.java:301 encodes one vector per FDB write; {@link #encodeAll} does the segment in one call).
 */
public final class PqEncoder {
  private PqEncoder() {}

  static MemorySegment flatten(Arena a, float[][][] c) {
    int m = c.length, k = c[0].length, sub = c[0][0].length;
    MemorySegment out = a.allocate(JAVA_FLOAT, (long) m * k * sub);
    for (int s = 0; s < m; s++)
      for (int ci = 0; ci < k; ci++) MemorySegment.copy(c[s][ci], 0, out, JAVA_FLOAT, ((long) s * k + ci) * sub * 4, sub);
    return out;
  }

  public static byte[] encode(float[][][] centroids, float[] vector) {
    int m = centroids.length, k = centroids[0].length, sub = centroids[0][0].length;
    try (Arena a = Arena.ofConfined()) {
      MemorySegment c = flatten(a, centroids), v = a.allocateFrom(JAVA_FLOAT, vector), out = a.allocate(m);
      check(call("vs_pq_encode", new MemoryLayout[] {ADDRESS, JAVA_INT, JAVA_INT, JAVA_INT, ADDRESS, ADDRESS}, c, m, k, sub, v, out));
      return out.toArray(JAVA_BYTE);
    }
  }

  /** codes[i] = encode(centroids, vectors.get(i)) for the whole list, one device pass. */
  public static byte[][] encodeAll(float[][][] centroids, List<float[]> vectors) {
    int m = centroids.length, k = centroids[0].length, sub = centroids[0][0].length, d = m * sub, n = vectors.size();
    try (Arena a = Arena.ofConfined()) {
      MemorySegment c = flatten(a, centroids), rows = a.allocate(JAVA_FLOAT, Math.max(1L, (long) n * d));
      for (int i = 0; i < n; i++) MemorySegment.copy(vectors.get(i), 0, rows, JAVA_FLOAT, (long) i * d * 4, d);
      MemorySegment out = a.allocate(Math.max(1L, (long) n * m));
      check(call("vs_pq_encode_batch", new MemoryLayout[] {ADDRESS, JAVA_INT, JAVA_INT, JAVA_INT, ADDRESS, JAVA_LONG, JAVA_LONG, ADDRESS},
          c, m, k, sub, rows, 0L, (long) n, out));
      byte[][] codes = new byte[n][];
      for (int i = 0; i < n; i++) codes[i] = out.asSlice((long) i * m, m).toArray(JAVA_BYTE);
      return codes;
    }
  }
}
