"""Development timing helper: batched-query brute force (batch.cu) on one GPU, device-timed."""
import sys

import numpy as np
import torch

sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.abspath(__file__))))
import vectorsearch_b200 as vs
from vectorsearch_b200 import _lib as L

vs.init(0)
lib = vs.load()
import os
if os.environ.get("VS_TF32"):
    vs.set_option("batch_fp16", 0)
if os.environ.get("VS_PAIRS"):
    vs.set_option("batch_pairs", int(os.environ["VS_PAIRS"]))
if os.environ.get("VS_WARPQ"):
    vs.set_option("batch_warp_min_queries", int(os.environ["VS_WARPQ"]))
if os.environ.get("VS_PFR"):
    vs.set_option("batch_prefetch_rounds", int(os.environ["VS_PFR"]))
if os.environ.get("VS_SELP"):
    vs.set_option("batch_select_ctas", int(os.environ["VS_SELP"]))
if os.environ.get("VS_GROUP"):
    vs.set_option("batch_group", int(os.environ["VS_GROUP"]))
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
d = int(sys.argv[2]) if len(sys.argv) > 2 else 128
k = int(sys.argv[3]) if len(sys.argv) > 3 else 10
metric = int(sys.argv[4]) if len(sys.argv) > 4 else 0
nqs = [int(x) for x in sys.argv[5].split(",")] if len(sys.argv) > 5 else [1, 4, 32, 128, 256, 1024]
seg = vs.Segment.generate(42, 0, n, d)
dev = torch.device("cuda:0")
st = torch.cuda.current_stream().cuda_stream
g = torch.Generator(device="cpu").manual_seed(43)
for nq in nqs:
    q = (torch.rand((nq, d), generator=g) * 2 - 1).to(dev)
    ids = torch.zeros(nq, k, dtype=torch.int64, device=dev)
    sc = torch.zeros(nq, k, dtype=torch.float64, device=dev)
    cn = torch.zeros(nq, dtype=torch.int32, device=dev)

    def run():
        L.check(lib.vs_bruteforce_topk_dev(seg.handle, q.data_ptr(), nq, k, metric, ids.data_ptr(), sc.data_ptr(), cn.data_ptr(), st))

    for _ in range(3):
        run()
    torch.cuda.synchronize()
    iters = 20 if nq <= 256 else 10
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        run()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    fl = 2.0 * n * d * nq
    print(f"n={n} d={d} k={k} metric={metric} nq={nq}: {ms*1e3:.1f} us/batch  {nq/ms*1e3:.3e} QPS  {n*nq/ms*1e3:.3e} evals/s  "
          f"{fl/ms/1e9:.1f} TFLOP/s  X-stream {n*d*4/ms/1e6:.1f} GB/s", flush=True)
seg.free()
