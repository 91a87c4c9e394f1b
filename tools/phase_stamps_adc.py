"""Development helper: per-CTA phase timestamps of the ADC fast scan (build with -DVS_PHASE_STAMPS)."""
import ctypes as C, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import vectorsearch_b200 as vs
from vectorsearch_b200 import _lib as L
vs.init(0); lib = vs.load()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
seg = vs.Segment.generate(42, 0, n, 128)
tr = vs.Segment.generate(42, 0, min(n, 1000000), 128)
cent = vs.PqTrainer.train(None, 128, 16, 256, 5, 42, segment=tr); tr.free()
seg.attach_pq(cent)
dev = torch.device("cuda:0")
q = torch.rand(1, 128, device=dev) * 2 - 1
ids = torch.zeros(1, 100, dtype=torch.int64, device=dev); sc = torch.zeros(1, 100, dtype=torch.float64, device=dev); cn = torch.zeros(1, dtype=torch.int32, device=dev)
st = torch.cuda.current_stream().cuda_stream
for it in range(5):
    L.check(lib.vs_adc_topk_dev(seg.handle, q.data_ptr(), 1, 100, ids.data_ptr(), sc.data_ptr(), cn.data_ptr(), st))
torch.cuda.synchronize()
buf = np.zeros(8 * 1024, dtype=np.uint64)
rc = lib.vs_debug_read_stamps_adc(buf.ctypes.data_as(C.c_void_p), buf.nbytes)
s = buf.reshape(1024, 8)[:148].astype(np.int64)
t0 = s[:, 0].min()
s = s - t0
names = ["start", "table_built", "warmup_done", "loop_done", "published", "Tf", "walked", "end"]
for i, nm in enumerate(names):
    c = s[:, i][s[:, i] > 0] if i >= 5 else s[:, i]
    if c.size == 0: continue
    print(f"{nm:12s} min {c.min()/1e3:7.2f} med {np.median(c)/1e3:7.2f} max {c.max()/1e3:7.2f} us")
ld = s[:, 3] - s[:, 2]
order = np.argsort(ld)
print("loop duration percentiles us:", [round(float(np.percentile(ld, p)) / 1e3, 1) for p in (0, 10, 50, 90, 95, 99, 100)])
print("slowest CTAs:", [(int(i), round(float(ld[i]) / 1e3, 1)) for i in order[-8:]])
print("fastest CTAs:", [(int(i), round(float(ld[i]) / 1e3, 1)) for i in order[:4]])
