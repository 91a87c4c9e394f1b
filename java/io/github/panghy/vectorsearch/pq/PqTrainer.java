package io.github.panghy.vectorsearch.pq;

import static io.github.panghy.vectorsearch.gpu.VsGpu.call;
import static io.github.panghy.vectorsearch.gpu.VsGpu.check;
import static java.lang.foreign.ValueLayout.ADDRESS;
import static java.lang.foreign.ValueLayout.JAVA_FLOAT;
import static java.lang.foreign.ValueLayout.JAVA_INT;
import static java.lang.foreign.ValueLayout.JAVA_LONG;

import java.lang.foreign.Arena;
import java.lang.foreign.MemoryLayout;
import java.lang.foreign.MemorySegment;
import java.util.List;

/**
 * Drop-in shim for J/pq/PqTrainer.java:28-91 (same static signature and exceptions). The device trainer reproduces the
 * reference's Lloyd iterations bit for bit: the shared java.util.Random, sampling with replacement, strict-&lt; argmin,
 * fp32 sums in row order, empty-cluster re-initialisation.
 */
public final class PqTrainer {
  private PqTrainer() {}

  public static float[][][] train(List<float[]> vectors, int dimension, int m, int k, int iterations, long seed) {
    if (m <= 0 || k <= 0 || dimension <= 0) throw new IllegalArgumentException("Invalid PQ params (m,k,dimension)");
    if (dimension % m != 0) throw new IllegalArgumentException("dimension must be divisible by m");
    int n = vectors.size(), sub = dimension / m;
    try (Arena a = Arena.ofConfined()) {
      MemorySegment rows = a.allocate(JAVA_FLOAT, Math.max(1L, (long) n * dimension));
      for (int i = 0; i < n; i++) MemorySegment.copy(vectors.get(i), 0, rows, JAVA_FLOAT, (long) i * dimension * 4, dimension);
      MemorySegment out = a.allocate(JAVA_FLOAT, (long) m * k * sub);
      check(call("vs_pq_train",
          new MemoryLayout[] {ADDRESS, JAVA_LONG, JAVA_LONG, JAVA_INT, JAVA_INT, JAVA_INT, JAVA_INT, JAVA_LONG, ADDRESS},
          rows, 0L, (long) n, dimension, m, k, iterations, seed, out));  // VS_EEMPTY -> IndexOutOfBoundsException, as data.get(0) throws
      float[][][] c = new float[m][k][sub];
      for (int s = 0; s < m; s++)
        for (int ci = 0; ci < k; ci++) MemorySegment.copy(out, JAVA_FLOAT, ((long) s * k + ci) * sub * 4, c[s][ci], 0, sub);
      return c;
    }
  }
}
