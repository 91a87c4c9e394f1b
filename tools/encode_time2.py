"""Is an encode call slower after the query paths have run (per-thread scratch, batched-path state)?"""
import sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np
import torch
import vectorsearch_b200 as vs

vs.init(0)

def encode_times(tag, n=10_000_000):
    seg = vs.Segment.generate(42, 0, n, 128)
    cent = vs.PqTrainer.train(None, 128, 16, 256, 5, 42, segment=seg)
    out = []
    for rep in range(3):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        seg.attach_pq(cent)
        torch.cuda.synchronize()
        out.append(time.perf_counter() - t0)
    seg.free()
    print(tag, " ".join(f"{t:.4f}" for t in out), flush=True)

encode_times("fresh process:            ")
seg = vs.Segment.generate(42, 0, 1_000_000, 128)
q = np.random.default_rng(0).random((1024, 128), dtype=np.float32)
seg.bruteforce_topk(q[0], 10)
encode_times("after one scan query:     ")
seg.bruteforce_topk(q, 10)
encode_times("after a 1024-query batch: ")
seg.free()
encode_times("after freeing that segment:")
if len(sys.argv) > 1:
    from oracle import pyoracle
    orc = pyoracle.get()
    rows = orc.gen_rows(42, 0, 1_000_000, 128)
    for i in range(5):
        orc.bruteforce_topk(rows, q[i], 10, threads=16)
    encode_times("after OpenMP oracle work: ")
