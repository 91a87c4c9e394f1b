"""One hot path per invocation, a handful of launches: what `ncu` captures for the profiles/ summaries.

    python tools/profile_paths.py scan|scan32|adc|batch|batch768|pq|rerank|knn|exchange

scan: C2 (1M x 128, one query); adc: C4 ADC top-100 over VS_ROWS rows (default 100M) + re-rank; batch: C2 query batch
1024; pq: encode of 10M x 128; knn: graph lists of a 20k x 128 segment; exchange: three ranks on this GPU
(vs_init_multi, the host-wait shape) -- publish, merge and all-reduce kernels.
"""
import os
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np
import torch

import vectorsearch_b200 as vs
from vectorsearch_b200 import _lib as L

which = sys.argv[1]
vs.init(0)
lib = vs.load()
dev = torch.device("cuda:0")
st = torch.cuda.current_stream().cuda_stream
rng = np.random.default_rng(0)

if which in ("scan", "scan32"):  # scan: the default single-query path (fp16-copy scan); scan32: the fp32 streaming scan
    vs.set_option("scan_fp16", 0 if which == "scan32" else 1)
    seg = vs.Segment.generate(42, 0, 1_000_000, 128)
    q = torch.from_numpy(rng.random((8, 128), dtype=np.float32) * 2 - 1).to(dev)
    ids = torch.empty((1, 10), dtype=torch.int64, device=dev); sc = torch.empty((1, 10), dtype=torch.float64, device=dev)
    cn = torch.empty((1,), dtype=torch.int32, device=dev)
    for i in range(8):
        L.check(lib.vs_bruteforce_topk_dev(seg.handle, q[i].data_ptr(), 1, 10, 0, ids.data_ptr(), sc.data_ptr(), cn.data_ptr(), st))
    torch.cuda.synchronize()
elif which in ("adc", "rerank"):
    n = int(os.environ.get("VS_ROWS", 100_000_000))
    seg = vs.Segment.generate(42, 0, n, 128)
    tr = vs.Segment.generate(42, 0, 1_000_000, 128)
    cent = vs.PqTrainer.train(None, 128, 16, 256, 5, 42, segment=tr); tr.free()
    seg.attach_pq(cent)
    q = torch.from_numpy(rng.random((6, 128), dtype=np.float32) * 2 - 1).to(dev)
    ids = torch.empty((1, 10), dtype=torch.int64, device=dev); sc = torch.empty((1, 10), dtype=torch.float64, device=dev)
    cn = torch.empty((1,), dtype=torch.int32, device=dev)
    for i in range(6):
        L.check(lib.vs_adc_rerank_topk_dev(seg.handle, q[i].data_ptr(), 1, 100, 10, 0, 0, ids.data_ptr(), sc.data_ptr(), cn.data_ptr(), st))
    torch.cuda.synchronize()
elif which == "batch":
    seg = vs.Segment.generate(42, 0, 1_000_000, 128)
    q = torch.from_numpy(rng.random((1024, 128), dtype=np.float32) * 2 - 1).to(dev)
    ids = torch.empty((1024, 10), dtype=torch.int64, device=dev); sc = torch.empty((1024, 10), dtype=torch.float64, device=dev)
    cn = torch.empty((1024,), dtype=torch.int32, device=dev)
    for i in range(4):
        L.check(lib.vs_bruteforce_topk_dev(seg.handle, q.data_ptr(), 1024, 10, 0, ids.data_ptr(), sc.data_ptr(), cn.data_ptr(), st))
    torch.cuda.synchronize()
elif which == "batch768":   # C5's per-GPU shape: 6.25M x 768, cosine top-50, 256 queries
    seg = vs.Segment.generate(42, 0, 6_250_000, 768)
    q = torch.from_numpy(rng.random((256, 768), dtype=np.float32) * 2 - 1).to(dev)
    ids = torch.empty((256, 50), dtype=torch.int64, device=dev); sc = torch.empty((256, 50), dtype=torch.float64, device=dev)
    cn = torch.empty((256,), dtype=torch.int32, device=dev)
    for i in range(4):
        L.check(lib.vs_bruteforce_topk_dev(seg.handle, q.data_ptr(), 256, 50, 1, ids.data_ptr(), sc.data_ptr(), cn.data_ptr(), st))
    torch.cuda.synchronize()
elif which == "pq":
    seg = vs.Segment.generate(42, 0, 10_000_000, 128)
    tr = vs.Segment.generate(42, 0, 1_000_000, 128)
    cent = vs.PqTrainer.train(None, 128, 16, 256, 5, 42, segment=tr); tr.free()
    vs.PqTrainer.train(None, 128, 16, 256, 2, 42, segment=seg)   # whole-image passes (10M rows per launch)
    seg.attach_pq(cent)
elif which == "knn":
    seg = vs.Segment.generate(42, 0, 20_000, 128)
    g = seg.knn_graph(32)
    print("knn", g[0][:5])
elif which == "exchange":
    vs.init_multi([0, 0, 0])
    seg = vs.Segment.generate(42, 0, 600_000, 128)
    q = rng.random((4, 128), dtype=np.float32) * 2 - 1
    for i in range(4):
        seg.bruteforce_topk(q[i], 10)
    cent = vs.PqTrainer.train(None, 128, 16, 256, 2, 42, segment=seg)
    seg.attach_pq(cent)
    seg.adc_rerank_topk(q[0], 100, 10)
    seg.free()
    vs.init(0)
print("ok", which)
