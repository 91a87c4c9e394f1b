"""Development helper: ADC top-100 alone over n rows, device-timed, before and after a 32-query launch."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import vectorsearch_b200 as vs
from vectorsearch_b200 import _lib as L
if os.environ.get("VS_DEV_LIB"):  # an experimental build (make BUILD=build_dev OUT=../libvsgpu_dev.so EXTRA=-D...)
    from pathlib import Path
    L.LIB_PATH = Path(os.environ["VS_DEV_LIB"]).resolve()
vs.init(0); lib = vs.load()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000_000
seg = vs.Segment.generate(42, 0, n, 128)
tr = vs.Segment.generate(42, 0, min(n, 1000000), 128)
cent = vs.PqTrainer.train(None, 128, 16, 256, 5, 42, segment=tr); tr.free()
seg.attach_pq(cent)
dev = torch.device("cuda:0")
qs = vs.Segment.generate(43, 0, 64, 128); q = torch.from_numpy(qs.rows()).to(dev); qs.free()
st = torch.cuda.current_stream().cuda_stream


def alone(nq, iters=20, label=""):
    ids = torch.zeros(nq, 100, dtype=torch.int64, device=dev); sc = torch.zeros(nq, 100, dtype=torch.float64, device=dev); cn = torch.zeros(nq, dtype=torch.int32, device=dev)
    for it in range(3):
        L.check(lib.vs_adc_topk_dev(seg.handle, q[it].data_ptr(), nq, 100, ids.data_ptr(), sc.data_ptr(), cn.data_ptr(), st))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for it in range(iters):
        L.check(lib.vs_adc_topk_dev(seg.handle, q[(5 + it) % 32].data_ptr(), nq, 100, ids.data_ptr(), sc.data_ptr(), cn.data_ptr(), st))
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    print(f"{label} nq={nq}: {ms*1e3:.1f} us per launch, {ms*1e3/nq:.1f} us per query, {n*16/ms/1e6/nq*nq/1:.0f} GB/s per launch-bytes x1", flush=True)


import pynvml
pynvml.nvmlInit(); hnd = pynvml.nvmlDeviceGetHandleByIndex(0)
def clk(): return pynvml.nvmlDeviceGetClockInfo(hnd, pynvml.NVML_CLOCK_SM), pynvml.nvmlDeviceGetClockInfo(hnd, pynvml.NVML_CLOCK_MEM), pynvml.nvmlDeviceGetPowerUsage(hnd) / 1000
alone(1, label="fresh"); print(clk())
alone(5, iters=40, label="b5 heavy"); print(clk())
alone(1, label="after b5 heavy"); print(clk())
alone(6, iters=1, label="b6 light"); print(clk())
alone(1, label="after b6"); print(clk())
alone(7, iters=1, label="b7 light"); print(clk())
alone(1, label="after b7"); print(clk())
alone(8, iters=1, label="b8 light"); print(clk())
alone(1, label="after b8 light"); print(clk())
import time; time.sleep(3)
alone(1, label="after 3 s idle"); print(clk())
