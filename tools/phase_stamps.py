"""Development helper: per-CTA phase timestamps of the TMA scan kernel (build with -DVS_PHASE_STAMPS)."""
import ctypes as C, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import vectorsearch_b200 as vs
from vectorsearch_b200 import _lib as L
vs.init(0); lib = vs.load()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
seg = vs.Segment.generate(42, 0, n, 128)
dev = torch.device("cuda:0")
q = torch.rand(1, 128, device=dev) * 2 - 1
ids = torch.zeros(1, 10, dtype=torch.int64, device=dev); sc = torch.zeros(1, 10, dtype=torch.float64, device=dev); cn = torch.zeros(1, dtype=torch.int32, device=dev)
st = torch.cuda.current_stream().cuda_stream
cudart = C.CDLL("libcudart.so.12")
for it in range(5):
    L.check(lib.vs_bruteforce_topk_dev(seg.handle, q.data_ptr(), 1, 10, 0, ids.data_ptr(), sc.data_ptr(), cn.data_ptr(), st))
torch.cuda.synchronize()
buf = np.zeros(8 * 1024, dtype=np.uint64)
rc = lib.vs_debug_read_stamps(buf.ctypes.data_as(C.c_void_p), buf.nbytes)
print('survivors S =', int(buf.reshape(1024, 8)[1000, 0]))
s = buf.reshape(1024, 8)[:148].astype(np.int64)
t0 = s[:, 0].min()
s = s - t0
names = ["start", "setup_done", "5th_tile", "loop_done", "end", "combined", "bound", "final"]
for i, nm in enumerate(names):
    c = s[:, i][s[:, i] > 0] if i >= 6 else s[:, i]
    if c.size == 0: continue
    print(f"{nm:12s} min {c.min()/1e3:7.2f} med {np.median(c)/1e3:7.2f} max {c.max()/1e3:7.2f} us")
print("epilogue (end - loop_done) of the last-finishing CTA:", (s[:, 4] - s[:, 3]).max() / 1e3)

last = int(np.argmax(s[:, 7]))
r = s[last]
print("last CTA", last, "combined %.2f ticket %.2f heads_loaded %.2f bound %.2f filtered %.2f ranked %.2f end %.2f" % tuple(x / 1e3 for x in (r[5], r[6], r[1], r[2], r[3], r[7], r[4])))
