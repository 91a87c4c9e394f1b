"""Development helper: phase timestamps of batch_select_kernel per query (build libvsgpu with -DVS_BQ_STAMPS)."""
import ctypes as C, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import vectorsearch_b200 as vs
from vectorsearch_b200 import _lib as L
if os.environ.get("VS_DEV_LIB"):  # the -DVS_BQ_STAMPS build (make BUILD=build_dev OUT=../libvsgpu_dev.so EXTRA=-DVS_BQ_STAMPS)
    from pathlib import Path
    L.LIB_PATH = Path(os.environ["VS_DEV_LIB"]).resolve()
vs.init(0); lib = vs.load()
nq = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
n, d, k, metric = (int(os.environ.get(e, v)) for e, v in (("VS_N", 1_000_000), ("VS_D", 128), ("VS_K", 10), ("VS_METRIC", 0)))
if os.environ.get("VS_PAIRS"):
    vs.set_option("batch_pairs", int(os.environ["VS_PAIRS"]))
if os.environ.get("VS_SELP"):
    vs.set_option("batch_select_ctas", int(os.environ["VS_SELP"]))
if os.environ.get("VS_SH_CTAS"):
    vs.set_option("scan_half_ctas", int(os.environ["VS_SH_CTAS"]))
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
if nq <= 2:  # scan_half_kernel: phases 0-9 from CTA 0, 10-15 from the last CTA
    names = ["start", "ring requested", "q2 known", "query staged", "loop done (warp 0)", "after pdl wait", "all warps done",
             "compacted + ranked", "published + fenced", "ticket drawn", "LAST CTA begins", "keys loaded", "bound found",
             "candidates listed", "scored", "end"]
    s16 = buf[4096:4096 + 32 * nq].reshape(nq, 32)[:, :16].astype(np.int64)
    for qi in range(nq):
        r = (s16[qi] - s16[qi, 0]) / 1e3
        print("query", qi)
        for i, nm in enumerate(names):
            print(f"  {nm:22s} {r[i]:7.2f}  (+{r[i] - r[i - 1] if i else 0.0:5.2f})")
    c = buf[1024:1024 + 2 * 320].reshape(320, 2).astype(np.int64)
    c = c[c[:, 0] > 0]
    t00 = c[:, 0].min()
    st_, en_ = (c[:, 0] - t00) / 1e3, (c[:, 1] - t00) / 1e3
    print(f"per CTA ({len(c)} CTAs of query 0): start min {st_.min():.2f} median {np.median(st_):.2f} max {st_.max():.2f} us | loop done min {en_.min():.2f}"
          f" p10 {np.percentile(en_, 10):.2f} median {np.median(en_):.2f} p90 {np.percentile(en_, 90):.2f} max {en_.max():.2f} us")
    order = np.argsort(en_)
    print("  slowest CTAs (index: start, done):", ", ".join(f"{i}: {st_[i]:.1f}, {en_[i]:.1f}" for i in order[-6:]))
    print("  fastest CTAs (index: start, done):", ", ".join(f"{i}: {st_[i]:.1f}, {en_[i]:.1f}" for i in order[:6]))
    sys.exit(0)
print("kernel span: %.1f us; CTA start spread: %.1f us" % ((s[:, 6].max() - t0) / 1e3, (s[:, 0].max() - t0) / 1e3))
for i in range(1, 7):
    dt = (s[:, i] - s[:, i - 1]) / 1e3
    print(f"{names[i]:18s} median {np.median(dt):7.2f}  p90 {np.percentile(dt, 90):7.2f}  max {dt.max():7.2f} us")
dt = (s[:, 6] - s[:, 0]) / 1e3
print(f"{'whole CTA':18s} median {np.median(dt):7.2f}  p90 {np.percentile(dt, 90):7.2f}  max {dt.max():7.2f} us")
