/* c_abi_smoke.c -- plain C caller of libvsgpu (what a JNI / FFM shim does), no Python and no torch in the process.
 * Built and run by tests/test_abi_and_sharding.py: gcc c_abi_smoke.c -lvsgpu.
 *   exit 0  : a device was found and every check passed
 *   exit 77 : no usable CUDA device -- vs_init failed with VS_ECUDA and a message saying there is no CPU fallback
 *   exit 1  : a check failed
 * The known answers are the reference's own (T/util/DistancesTest.java:36-47: l2Squared 25.0, l2 5.0;
 * T/api/VectorIndexTest.java:599-609: q = (1,0,0) over 4 vectors of dimension 3, first hit row 0). */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../include/vsgpu.h"

#define CHECK(cond)                                                        \
  do {                                                                     \
    if (!(cond)) {                                                         \
      fprintf(stderr, "FAILED %s:%d: %s (%s)\n", __FILE__, __LINE__, #cond, vs_last_error()); \
      return 1;                                                            \
    }                                                                      \
  } while (0)

int main(int argc, char** argv) {
  int ngpu = argc > 1 ? atoi(argv[1]) : 1;
  printf("libvsgpu version %d\n", vs_version());
  int32_t devs[16] = {0};
  int rc = ngpu > 1 ? vs_init_multi(ngpu, devs) /* ranks share device 0 */ : vs_init(0);
  if (rc == VS_ECUDA) {
    printf("no device: %s\n", vs_last_error());
    return strstr(vs_last_error(), "no CPU fallback") ? 77 : 1;
  }
  CHECK(rc == VS_OK);
  /* DistancesTest: (0,0,0)-(3,4,0) */
  float a[3] = {0, 0, 0}, b[3] = {3, 4, 0};
  double out = -1;
  CHECK(vs_l2_squared(a, b, 3, &out) == VS_OK && out == 25.0);
  CHECK(vs_l2(a, b, 3, &out) == VS_OK && out == 5.0);
  CHECK(vs_cosine(a, b, 3, &out) == VS_OK && out == 0.0); /* zero norm -> 0.0 */
  /* VectorIndexTest.l2_query: 4 vectors, q = e0, k = 2 */
  float rows[12] = {1, 0, 0, 0, 1, 0, 0, 0, 1, 0.9f, 0.1f, 0};
  float q[3] = {1, 0, 0};
  uint64_t h = 0;
  CHECK(vs_segment_upload(rows, 4, 3, NULL, 0, &h) == VS_OK);
  int64_t ids[2];
  double sc[2];
  int32_t cnt = 0;
  CHECK(vs_bruteforce_topk(h, q, 1, 2, VS_METRIC_L2, ids, sc, &cnt) == VS_OK);
  CHECK(cnt == 2 && ids[0] == 0 && ids[1] == 3 && sc[0] == 0.0 && sc[1] < 0.0);
  /* PqEncoderTest.encodes_expected_codes (T/pq/PqEncoderTest.java:12-23): m = 2, k = 2, subDim = 2 -> codes [0,1] */
  float cent[2 * 2 * 2] = {0, 0, 1, 1, 0, 0, 2, 2};
  float v[4] = {0.1f, 0.1f, 1.9f, 2.1f};
  uint8_t codes[2] = {9, 9};
  CHECK(vs_pq_encode(cent, 2, 2, 2, v, codes) == VS_OK && codes[0] == 0 && codes[1] == 1);
  /* PqTrainerTest: IllegalArgumentException for dimension % m != 0 */
  float cout_[8];
  CHECK(vs_pq_train(rows, 0, 4, 3, 2, 2, 1, 42, cout_) == VS_EINVAL);
  /* the codebook wire format round-trips */
  uint8_t blob[256];
  int64_t len = 0;
  CHECK(vs_codebook_encode(cent, 2, 2, 2, blob, sizeof blob, &len) == VS_OK && len == 2 + 2 + 2 * (2 + 16));
  float back[8];
  int32_t M = 0, K = 0, sd = 0;
  CHECK(vs_codebook_decode(blob, len, back, 8, &M, &K, &sd) == VS_OK && M == 2 && K == 2 && sd == 2);
  CHECK(memcmp(back, cent, sizeof cent) == 0);
  CHECK(vs_segment_free(h) == VS_OK);
  CHECK(vs_segment_free(h) == VS_EHANDLE);
  CHECK(vs_shutdown() == VS_OK);
  printf("c_abi_smoke ok (%d device slot%s, %lld kernel launches)\n", ngpu, ngpu > 1 ? "s" : "", (long long)vs_kernel_launch_count());
  return 0;
}
