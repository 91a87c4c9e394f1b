"""Development timing helper (not the bench): the single-query scan on one GPU, launched alone and in a stream of
queries alternating between two CUDA streams, for the fp32 scan (scan_fp16 = 0) and the fp16-copy scan with one or two
CTAs per SM (scan_half_ctas).  Every variant's ids and scores are compared with the fp32 scan's."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vectorsearch_b200 as vs
from vectorsearch_b200 import _lib as L

vs.init(0)
lib = vs.load()
n = int(os.environ.get("VS_N", 1_000_000))
d = int(os.environ.get("VS_D", 128))
k = int(os.environ.get("VS_K", 10))
metric = int(os.environ.get("VS_METRIC", 0))
NQ = 256
seg = vs.Segment.generate(42, 0, n, d)
dev = torch.device("cuda:0")
torch.manual_seed(1)
q = torch.rand(NQ, d, device=dev) * 2 - 1
streams = [torch.cuda.Stream(), torch.cuda.Stream(), torch.cuda.Stream()]
outs = [(torch.zeros(NQ, k, dtype=torch.int64, device=dev), torch.zeros(NQ, k, dtype=torch.float64, device=dev),
         torch.zeros(NQ, dtype=torch.int32, device=dev)) for _ in range(3)]
main = torch.cuda.current_stream()


def call(i, slot, st):
    ids, sc, cn = outs[slot]
    L.check(lib.vs_bruteforce_topk_dev(seg.handle, q.data_ptr() + (i % NQ) * d * 4, 1, k, metric, ids.data_ptr() + (i % NQ) * k * 8,
                                       sc.data_ptr() + (i % NQ) * k * 8, cn.data_ptr() + (i % NQ) * 4, st))


def alone(iters=60):
    for i in range(5):
        call(i, 0, main.cuda_stream)
    torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(iters)]
    for i in range(iters):
        ev[i][0].record(main)
        call(i, 0, main.cuda_stream)
        ev[i][1].record(main)
    torch.cuda.synchronize()
    t = sorted(a.elapsed_time(b) for a, b in ev)
    return t[len(t) // 2] * 1e3, sum(t) / len(t) * 1e3


def flow(nstreams, iters=NQ):
    for i in range(8):
        call(i, i % nstreams, streams[i % nstreams].cuda_stream)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(main)
    for s in streams[:nstreams]:
        s.wait_stream(main)
    for i in range(iters):
        call(i, 0, streams[i % nstreams].cuda_stream)
    for s in streams[:nstreams]:
        main.wait_stream(s)
    e1.record(main)
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3


ref = None
variants = [("fp32 scan", dict(scan_fp16=0)), ("fp16 copy, 1 CTA/SM", dict(scan_fp16=1, scan_half_ctas=1)),
            ("fp16 copy, 2 CTAs/SM", dict(scan_fp16=1, scan_half_ctas=2))]
for name, opts in variants:
    for o, v in opts.items():
        vs.set_option(o, v)
    for reserve in (0, 4):
        vs.set_option("scan_reserve_sms", reserve)
        med, mean = alone()
        f1 = flow(1)
        f2 = " ".join(f"{flow(2):5.1f}" for _ in range(3))
        f3 = " ".join(f"{flow(3):5.1f}" for _ in range(3))
        f20 = " ".join(f"{flow(3, 20):5.1f}" for _ in range(3))
        ids, sc, _ = outs[0]
        got = (ids.cpu().numpy().copy(), sc.cpu().numpy().copy())
        if ref is None:
            ref = got
        same = np.array_equal(ref[0], got[0]) and np.array_equal(ref[1].view(np.uint64), got[1].view(np.uint64))
        print(f"{name:22s} reserve {reserve}: alone median {med:6.1f} mean {mean:6.1f} us | one stream {f1:6.1f} | two streams {f2} | three streams {f3} | three streams, 20 queries {f20} us per query"
              f" | identical to fp32 scan: {same}", flush=True)
seg.free()
