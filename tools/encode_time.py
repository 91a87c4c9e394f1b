"""Where does an encode call spend its time?  Host wall clock vs device time of three consecutive encodes of a resident
10M x 128 segment (M = 16, K = 256) after a training run, plus a 1M-row segment."""
import sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np
import torch
import vectorsearch_b200 as vs

vs.init(0)
for n in (10_000_000, 1_000_000):
    seg = vs.Segment.generate(42, 0, n, 128)
    t0 = time.perf_counter(); cent = vs.PqTrainer.train(None, 128, 16, 256, 5, 42, segment=seg); t_train = time.perf_counter() - t0
    t0 = time.perf_counter(); cent = vs.PqTrainer.train(None, 128, 16, 256, 5, 42, segment=seg); t_train2 = time.perf_counter() - t0
    print(f"n={n}: train {t_train:.4f}s, again {t_train2:.4f}s, launches {vs.kernel_launch_count()}")
    for rep in range(4):
        torch.cuda.synchronize()
        l0 = vs.kernel_launch_count()
        t0 = time.perf_counter()
        seg.attach_pq(cent)
        torch.cuda.synchronize()
        print(f"  attach_pq #{rep}: {time.perf_counter() - t0:.4f}s, {vs.kernel_launch_count() - l0} launches")
    for rep in range(2):
        t0 = time.perf_counter()
        codes = vs.PqEncoder.encode_batch(cent, segment=seg)
        print(f"  encode_batch(segment) #{rep}: {time.perf_counter() - t0:.4f}s (includes {codes.nbytes / 1e6:.0f} MB D2H)")
    seg.free()
