package io.github.panghy.vectorsearch.util;

import static io.github.panghy.vectorsearch.gpu.VsGpu.call;
import static io.github.panghy.vectorsearch.gpu.VsGpu.check;
import static java.lang.foreign.ValueLayout.ADDRESS;
import static java.lang.foreign.ValueLayout.JAVA_DOUBLE;
import static java.lang.foreign.ValueLayout.JAVA_FLOAT;
import static java.lang.foreign.ValueLayout.JAVA_INT;

import java.lang.foreign.Arena;
import java.lang.foreign.MemoryLayout;
import java.lang.foreign.MemorySegment;

/**
 * Drop-in shim with the static signatures of the reference class (J/util/Distances.java:31-153), so that
 * DistanceAndPqBenchmark and every other caller compile unchanged; each call is one tiny kernel launch in libvsgpu and
 * returns the reference's own double (same per-lane fp32 FMA, ordered lane reduction, fp64 tail). A per-pair call is an
 * API-parity path, not a throughput path: queries go through {@code gpu.GpuScoring}.
 */
public final class Distances {
  private Distances() {}

  private static final MemoryLayout[] PAIR = {ADDRESS, ADDRESS, JAVA_INT, ADDRESS};

  private static double pair(String fn, float[] a, int aOff, float[] b, int bOff, int len) {
    try (Arena ar = Arena.ofConfined()) {
      MemorySegment sa = ar.allocate(JAVA_FLOAT, Math.max(1, len)), sb = ar.allocate(JAVA_FLOAT, Math.max(1, len));
      MemorySegment.copy(a, aOff, sa, JAVA_FLOAT, 0, len);
      MemorySegment.copy(b, bOff, sb, JAVA_FLOAT, 0, len);
      MemorySegment out = ar.allocate(JAVA_DOUBLE);
      check(call(fn, PAIR, sa, sb, len, out));
      return out.get(JAVA_DOUBLE, 0);
    }
  }

  public static double l2(float[] a, float[] b) {
    return pair("vs_l2", a, 0, b, 0, a.length);
  }

  public static double l2Squared(float[] a, float[] b) {
    return pair("vs_l2_squared", a, 0, b, 0, a.length);
  }

  public static double l2Squared(float[] a, int aOffset, float[] b, int bOffset, int length) {
    return pair("vs_l2_squared", a, aOffset, b, bOffset, length);
  }

  public static double dot(float[] a, float[] b) {
    return pair("vs_dot", a, 0, b, 0, a.length);
  }

  public static double norm(float[] a) {
    try (Arena ar = Arena.ofConfined()) {
      MemorySegment sa = ar.allocateFrom(JAVA_FLOAT, a), out = ar.allocate(JAVA_DOUBLE);
      check(call("vs_norm", new MemoryLayout[] {ADDRESS, JAVA_INT, ADDRESS}, sa, a.length, out));
      return out.get(JAVA_DOUBLE, 0);
    }
  }

  public static double cosine(float[] a, float[] b) {
    return pair("vs_cosine", a, 0, b, 0, a.length);
  }
}
