import sys, os
sys.path.insert(0, "/root/repo")
import torch
import vectorsearch_b200 as vs
from vectorsearch_b200 import _lib as L
vs.init(0); lib = vs.load()
n = 100_000_000
seg = vs.Segment.generate(42, 0, n, 128)
tr = vs.Segment.generate(42, 0, 1000000, 128)
cent = vs.PqTrainer.train(None, 128, 16, 256, 5, 42, segment=tr); tr.free()
seg.attach_pq(cent)
dev = torch.device("cuda:0")
qs = vs.Segment.generate(43, 0, 64, 128); q = torch.from_numpy(qs.rows()).to(dev); qs.free()
st = torch.cuda.current_stream().cuda_stream
ids = torch.zeros(8, 100, dtype=torch.int64, device=dev); sc = torch.zeros(8, 100, dtype=torch.float64, device=dev); cn = torch.zeros(8, dtype=torch.int32, device=dev)
def go(nq, it):
    L.check(lib.vs_adc_topk_dev(seg.handle, q[it].data_ptr(), nq, 100, ids.data_ptr(), sc.data_ptr(), cn.data_ptr(), st))
    torch.cuda.synchronize()
for it in range(4): go(1, it)
go(2, 4)
for it in range(4): go(1, 5 + it)
