"""Development timing helper: host-call latency of one query through Segment.bruteforce_topk (H2D and D2H inside)."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vectorsearch_b200 as vs

vs.init(0)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
seg = vs.Segment.generate(42, 0, n, 128)
rng = np.random.default_rng(1)
qs = rng.random((600, 128), dtype=np.float32) * 2 - 1
for rep in range(3):
    for i in range(50):
        seg.bruteforce_topk(qs[i], 10)
    t0 = time.perf_counter()
    for i in range(500):
        r = seg.bruteforce_topk(qs[50 + i], 10)
    dt = (time.perf_counter() - t0) / 500
    print(f"{dt * 1e6:.1f} us per query, top-1 {r[0][0]}", flush=True)
seg.free()
