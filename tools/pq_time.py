"""Development timing helper: PQ assignment (encode) and training on one GPU."""
import sys
import time

import numpy as np
import torch

sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.abspath(__file__))))
import vectorsearch_b200 as vs

vs.init(0)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
seg = vs.Segment.generate(42, 0, n, 128)
tr = vs.Segment.generate(42, 0, min(n, 1_000_000), 128)
cent = vs.PqTrainer.train(None, 128, 16, 256, 5, 42, segment=tr)
tr.free()
for tc in (2, 0):
    vs.set_option("pq_tensor_cores", tc)
    seg.attach_pq(cent)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(3):
        seg.attach_pq(cent)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / 3
    codes = seg.codes(0, 100000)
    print(f"encode n={n} tensor_cores={tc}: {dt*1e3:.2f} ms  {n/dt:.3e} vectors/s  {n*16*256/dt:.3e} sub-distance evals/s  checksum {int(codes.astype(np.int64).sum())}", flush=True)
    t0 = time.perf_counter()
    c2 = vs.PqTrainer.train(None, 128, 16, 256, 5, 42, segment=seg)
    dt = time.perf_counter() - t0
    print(f"train 5 iters n={n} tensor_cores={tc}: {dt*1e3:.1f} ms  checksum {float(np.abs(c2).sum()):.6f}", flush=True)
seg.free()
