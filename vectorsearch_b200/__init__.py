"""vectorsearch_b200 -- B200-native scoring path of panghy/vectorsearch behind the C ABI of libvsgpu.

Host-side mirror of the reference's operator surface for this path (same names, argument meaning
and error behaviour): `Distances`, `PqEncoder`, `PqTrainer` (J/util/Distances.java, J/pq/*.java)
and `Segment`, the resident-segment scorer that replaces the loops of
J/fdb/FdbVectorIndex.java:676-721 (brute force), :754-769 (ADC) and :997-1043 (re-rank).
"""
from ._lib import METRIC_COSINE, METRIC_L2, VsError, load  # noqa: F401
from .ops import (AdcQuery, Distances, PqEncoder, PqTrainer, Residency, Segment, build_lut,  # noqa: F401
                  codebook_decode, codebook_encode, device_count, device_info, init, init_multi,
                  kernel_launch_count, merge_topk, pq_approx_distance, pq_lut_distance,
                  set_option, set_simd_lanes, shutdown)
