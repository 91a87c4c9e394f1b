// GoldenDump.java -- run by a maintainer WITH A JVM to turn "parity pinned by restatement" into "pinned by the reference".
//
// It calls the reference's own classes (io.github.panghy.vectorsearch.util.Distances, .pq.PqTrainer, .pq.PqEncoder --
// the UNMODIFIED ones, not the shims under java/) on seeded inputs and prints one JSON document with every result as
// the hex bit pattern of the double / float.  tests/test_oracle_golden.py::test_jvm_golden_vectors consumes
// tests/golden/jvm_golden.json when it exists and compares the C oracle bit for bit (with the lane count the JVM
// reports), which settles the one modelling assumption nobody can check in a JVM-less image: that
// FloatVector.reduceLanes(ADD) adds the lanes in ascending order (oracle/vs_oracle.c:93-97).
//
//   ./gradlew classes
//   java --add-modules jdk.incubator.vector -cp build/classes/java/main tools/GoldenDump.java > tests/golden/jvm_golden.json
//
// Inputs (reproduced by the test with the oracle's java.util.Random restatement):
//   distances: for dim in DIMS: Random(42 + dim); a[i] = nextFloat()*2-1 for all i, then b likewise (T/util/DistancesTest.java:160-166)
//   pq: Random(7): 2000 rows x 32 dims, nextFloat()*2-1 row by row; train(rows, 32, 4, 16, 5, 42); codes of the first 100 rows;
//       query = Random(8) 32 floats; lut[s][ci] as FdbVectorIndex.buildLut (:1067-1079); approx of the first 100 code rows
//       as pqApproxDistance (:1057-1065); brute-force scores -l2(q, row) of the first 200 rows and their stable descending order.
import io.github.panghy.vectorsearch.pq.PqEncoder;
import io.github.panghy.vectorsearch.pq.PqTrainer;
import io.github.panghy.vectorsearch.util.Distances;
import java.util.ArrayList;
import java.util.Comparator;
import java.util.List;
import java.util.Random;
import jdk.incubator.vector.FloatVector;

public class GoldenDump {
  static final int[] DIMS = {1, 3, 7, 16, 17, 100, 128, 768, 1000};

  static float[] vec(Random r, int n) {
    float[] v = new float[n];
    for (int i = 0; i < n; i++) v[i] = r.nextFloat() * 2f - 1f;
    return v;
  }

  static String d(double x) {
    return "\"" + Long.toHexString(Double.doubleToRawLongBits(x)) + "\"";
  }

  static String f(float x) {
    return "\"" + Integer.toHexString(Float.floatToRawIntBits(x)) + "\"";
  }

  public static void main(String[] args) {
    StringBuilder sb = new StringBuilder();
    sb.append("{\n \"lanes\": ").append(FloatVector.SPECIES_PREFERRED.length());
    sb.append(",\n \"java\": \"").append(System.getProperty("java.version")).append("\"");
    sb.append(",\n \"distances\": [");
    for (int t = 0; t < DIMS.length; t++) {
      int dim = DIMS[t];
      Random r = new Random(42 + dim);
      float[] a = vec(r, dim), b = vec(r, dim);
      sb.append(t == 0 ? "\n" : ",\n").append("  {\"dim\": ").append(dim)
          .append(", \"l2sq\": ").append(d(Distances.l2Squared(a, b)))
          .append(", \"l2\": ").append(d(Distances.l2(a, b)))
          .append(", \"dot\": ").append(d(Distances.dot(a, b)))
          .append(", \"norm_a\": ").append(d(Distances.norm(a)))
          .append(", \"cosine\": ").append(d(Distances.cosine(a, b)))
          .append(", \"l2sq_sub\": ").append(d(Distances.l2Squared(a, dim / 3, b, dim / 4, dim - dim / 3))).append("}");
    }
    sb.append("\n ],\n");
    int n = 2000, dim = 32, m = 4, k = 16, sub = dim / m;
    Random r = new Random(7);
    List<float[]> rows = new ArrayList<>();
    for (int i = 0; i < n; i++) rows.add(vec(r, dim));
    float[][][] c = PqTrainer.train(rows, dim, m, k, 5, 42L);
    sb.append(" \"pq\": {\"n\": ").append(n).append(", \"dim\": ").append(dim).append(", \"m\": ").append(m).append(", \"k\": ").append(k);
    sb.append(",\n  \"centroids\": [");
    for (int s = 0; s < m; s++)
      for (int ci = 0; ci < k; ci++)
        for (int j = 0; j < sub; j++) sb.append(s + ci + j == 0 ? "" : ",").append(f(c[s][ci][j]));
    sb.append("],\n  \"codes\": [");
    byte[][] codes = new byte[100][];
    for (int i = 0; i < 100; i++) {
      codes[i] = PqEncoder.encode(c, rows.get(i));
      for (int s = 0; s < m; s++) sb.append(i + s == 0 ? "" : ",").append(codes[i][s] & 0xFF);
    }
    float[] q = vec(new Random(8), dim);
    double[][] lut = new double[m][k];
    sb.append("],\n  \"lut\": [");
    for (int s = 0; s < m; s++)
      for (int ci = 0; ci < k; ci++) {
        lut[s][ci] = Distances.l2Squared(q, s * sub, c[s][ci], 0, sub);  // FdbVectorIndex.buildLut
        sb.append(s + ci == 0 ? "" : ",").append(d(lut[s][ci]));
      }
    sb.append("],\n  \"approx\": [");
    for (int i = 0; i < 100; i++) {
      double ad = 0.0;  // FdbVectorIndex.pqApproxDistance
      for (int s = 0; s < m; s++) {
        int ci = codes[i][s] & 0xFF;
        if (ci >= k) continue;
        ad += lut[s][ci];
      }
      sb.append(i == 0 ? "" : ",").append(d(ad));
    }
    sb.append("],\n  \"scores\": [");
    double[] score = new double[200];
    List<Integer> order = new ArrayList<>();
    for (int i = 0; i < 200; i++) {
      score[i] = -Distances.l2(q, rows.get(i));  // searchBruteForceSegment :691-693
      order.add(i);
      sb.append(i == 0 ? "" : ",").append(d(score[i]));
    }
    order.sort(Comparator.comparingDouble((Integer i) -> score[i]).reversed());  // :708
    sb.append("],\n  \"order\": ").append(order.subList(0, 20)).append("}\n}\n");
    System.out.print(sb);
  }
}
