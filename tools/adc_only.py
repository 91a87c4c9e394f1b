"""Development helper: a few ADC top-100 queries over n rows (for ncu)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import vectorsearch_b200 as vs
from vectorsearch_b200 import _lib as L
vs.init(0); lib = vs.load()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 4
seg = vs.Segment.generate(42, 0, n, 128)
tr = vs.Segment.generate(42, 0, min(n, 1000000), 128)
cent = vs.PqTrainer.train(None, 128, 16, 256, 5, 42, segment=tr); tr.free()
seg.attach_pq(cent)
dev = torch.device("cuda:0")
q = torch.rand(iters, 128, device=dev) * 2 - 1
ids = torch.zeros(1, 100, dtype=torch.int64, device=dev); sc = torch.zeros(1, 100, dtype=torch.float64, device=dev); cn = torch.zeros(1, dtype=torch.int32, device=dev)
st = torch.cuda.current_stream().cuda_stream
for it in range(iters):
    L.check(lib.vs_adc_topk_dev(seg.handle, q[it].data_ptr(), 1, 100, ids.data_ptr(), sc.data_ptr(), cn.data_ptr(), st))
torch.cuda.synchronize()
print("ok", ids[0, :5].tolist())
