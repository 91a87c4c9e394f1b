"""Development helper: phase timestamps of batch_select_kernel per query (build libvsgpu with -DVS_BQ_STAMPS)."""
import ctypes as C, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import vectorsearch_b200 as vs
from vectorsearch_b200 import _lib as L
vs.init(0); lib = vs.load()
nq = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
n, d, k, metric = (int(os.environ.get(e, v)) for e, v in (("VS_N", 1_000_000), ("VS_D", 128), ("VS_K", 10), ("VS_METRIC", 0)))
if os.environ.get("VS_PAIRS"):
    vs.set_option("batch_pairs", int(os.environ["VS_PAIRS"]))
if os.environ.get("VS_SELP"):
    vs.set_option("batch_select_ctas", int(os.environ["VS_SELP"]))
if os.environ.get("VS_GROUP"):
    vs.set_option("batch_group", int(os.environ["VS_GROUP"]))
seg = vs.Segment.generate(42, 0, n, d)
dev = torch.device("cuda:0")
q = torch.rand(nq, d, device=dev) * 2 - 1
ids = torch.zeros(nq, k, dtype=torch.int64, device=dev); sc = torch.zeros(nq, k, dtype=torch.float64, device=dev); cn = torch.zeros(nq, dtype=torch.int32, device=dev)
st = torch.cuda.current_stream().cuda_stream
for it in range(4):
    L.check(lib.vs_bruteforce_topk_dev(seg.handle, q.data_ptr(), nq, k, metric, ids.data_ptr(), sc.data_ptr(), cn.data_ptr(), st))
torch.cuda.synchronize()
buf = np.zeros(8 * 1024, dtype=np.uint64)
lib.vs_debug_read_stamps_batch.argtypes = [C.c_void_p, C.c_int64]
lib.vs_debug_read_stamps_batch(buf.ctypes.data_as(C.c_void_p), buf.nbytes)
s = buf.reshape(1024, 8)[:min(nq, 1024), :7].astype(np.int64)
t0 = s[:, 0].min()
names = ["start", "setup(q, slack)", "phase1 (T)", "phase2 (list)", "prefetch issued", "phase3 (exact)", "epilogue"]
if nq <= 2:  # scan_half_kernel: CTA 0 records 0-3, the last CTA 4-7
    s8 = buf.reshape(1024, 8)[:nq].astype(np.int64)
    for qi in range(nq):
        r = s8[qi] - s8[qi, 0]
        print("query", qi, " ".join(f"{nm} {x/1e3:.2f}" for nm, x in zip(["start", "query staged", "loop done (CTA 0)", "published (CTA 0)", "last CTA begins", "candidates", "scored", "end"], r)))
    sys.exit(0)
print("kernel span: %.1f us; CTA start spread: %.1f us" % ((s[:, 6].max() - t0) / 1e3, (s[:, 0].max() - t0) / 1e3))
for i in range(1, 7):
    dt = (s[:, i] - s[:, i - 1]) / 1e3
    print(f"{names[i]:18s} median {np.median(dt):7.2f}  p90 {np.percentile(dt, 90):7.2f}  max {dt.max():7.2f} us")
dt = (s[:, 6] - s[:, 0]) / 1e3
print(f"{'whole CTA':18s} median {np.median(dt):7.2f}  p90 {np.percentile(dt, 90):7.2f}  max {dt.max():7.2f} us")
