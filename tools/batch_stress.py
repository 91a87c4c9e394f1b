"""Development stress: batched results must equal the per-query kernel's, repeatedly, in every mode."""
import sys

import numpy as np

sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.abspath(__file__))))
import vectorsearch_b200 as vs

vs.init(0)
rng = np.random.default_rng(1)
shapes = [(300, 32, 1, 2), (7, 128, 10, 4), (9000, 72, 10, 5), (40000, 128, 10, 37), (70000, 64, 33, 260), (20000, 768, 50, 130)]
bad = 0
for (n, d, k, nq) in shapes:
    rows = (rng.random((n, d), dtype=np.float32) * 2 - 1)
    qs = (rng.random((nq, d), dtype=np.float32) * 2 - 1)
    vs.set_option("batch_min_queries", 1 << 30)
    seg = vs.Segment.upload(rows)
    ref = [seg.bruteforce_topk(qs[i], k) for i in range(nq)]
    seg.free()
    for mode in ("fp16", "tf32", "fp16-warp"):
        vs.set_option("batch_min_queries", 2)
        vs.set_option("batch_min_rows", 1)
        vs.set_option("batch_fp16", 0 if mode == "tf32" else 1)
        vs.set_option("batch_warp_min_queries", 2 if "warp" in mode else 0)
        fails = 0
        for trial in range(40):
            seg = vs.Segment.upload(rows)
            ids, sc, cn = seg.bruteforce_topk(qs, k)
            seg.free()
            for i in range(nq):
                c = len(ref[i][0])
                if cn[i] != c or not np.array_equal(ids[i, :c], ref[i][0]) or not np.array_equal(sc[i, :c], ref[i][1]):
                    fails += 1
                    if fails <= 2:
                        print("  mismatch", (n, d, k, nq), mode, "trial", trial, "query", i, ids[i, :4], ref[i][0][:4], sc[i, :2], ref[i][1][:2])
                    break
        print((n, d, k, nq), mode, "fails", fails, "/ 40", flush=True)
        bad += fails
print("TOTAL FAILS", bad)
