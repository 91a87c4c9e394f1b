"""Development timing helper (not the bench): device-timed kernels on one GPU."""
import ctypes as C
import sys
import time

import numpy as np
import torch

sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.abspath(__file__))))
import vectorsearch_b200 as vs
from vectorsearch_b200 import _lib as L

vs.init(0)
lib = vs.load()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
d = 128
seg = vs.Segment.generate(42, 0, n, d)
dev = torch.device("cuda:0")
q = torch.rand(1, d, device=dev) * 2 - 1
ids = torch.zeros(1, 10, dtype=torch.int64, device=dev)
sc = torch.zeros(1, 10, dtype=torch.float64, device=dev)
cn = torch.zeros(1, dtype=torch.int32, device=dev)
st = torch.cuda.current_stream().cuda_stream


def timeit(fn, iters=50, warm=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3  # us


def bf():
    L.check(lib.vs_bruteforce_topk_dev(seg.handle, q.data_ptr(), 1, 10, 0, ids.data_ptr(), sc.data_ptr(), cn.data_ptr(), st))


us = timeit(bf) if not __import__("os").environ.get('ADC_ONLY') else 1.0
print(f"bruteforce L2 top-10 n={n} d={d}: {us:.1f} us  {n*d*4/us/1e3:.1f} GB/s  {n/us*1e6:.3e} evals/s")
qh = q.cpu().numpy()[0]
t0 = time.perf_counter()
for _ in range(200):
    seg.bruteforce_topk(qh, 10)
print(f"host e2e: {(time.perf_counter()-t0)/200*1e6:.1f} us/query")

import os
if os.environ.get('BF_ONLY'):
    seg.free(); sys.exit(0)
# ADC
M, K = 16, 256
t0 = time.perf_counter()
if os.environ.get('ADC_ONLY'):
    tr = vs.Segment.generate(42, 0, min(n, 1000000), d)
    cent = vs.PqTrainer.train(None, d, M, K, 5, 42, segment=tr)
    tr.free()
else:
    cent = vs.PqTrainer.train(None, d, M, K, 5, 42, segment=seg)
print(f"pq_train 5 iters n={n}: {time.perf_counter()-t0:.3f} s")
t0 = time.perf_counter()
seg.attach_pq(cent)
print(f"attach+encode: {time.perf_counter()-t0:.3f} s")
ids100 = torch.zeros(1, 100, dtype=torch.int64, device=dev)
ap = torch.zeros(1, 100, dtype=torch.float64, device=dev)


def adc():
    L.check(lib.vs_adc_topk_dev(seg.handle, q.data_ptr(), 1, 100, ids100.data_ptr(), ap.data_ptr(), cn.data_ptr(), st))


us = timeit(adc)
st8 = (C.c_uint32 * 8)()
lib.vs_debug_adc_stats.argtypes = [C.POINTER(C.c_uint32)]
lib.vs_debug_adc_stats(st8)
print("adc stats (55 launches): candidates", st8[0], "T_final", st8[1], "survivors", st8[2], "fallbacks", st8[3])
print(f"ADC top-100 n={n} M={M}: {us:.1f} us  {n*M/us/1e3:.1f} GB/s  {n/us*1e6:.3e} evals/s")


def adcr():
    L.check(lib.vs_adc_rerank_topk_dev(seg.handle, q.data_ptr(), 1, 100, 10, 0, 0, ids.data_ptr(), sc.data_ptr(), cn.data_ptr(), st))


us = timeit(adcr)
print(f"ADC top-100 + rerank top-10: {us:.1f} us")
seg.free()
