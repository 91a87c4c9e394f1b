"""Development helper: a few tensor-core PQ encodes over n rows (for ncu)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vectorsearch_b200 as vs
vs.init(0)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
mode = int(sys.argv[2]) if len(sys.argv) > 2 else 2
seg = vs.Segment.generate(42, 0, n, 128)
tr = vs.Segment.generate(42, 0, min(n, 1_000_000), 128)
cent = vs.PqTrainer.train(None, 128, 16, 256, 5, 42, segment=tr)
tr.free()
vs.set_option("pq_tensor_cores", mode)
for _ in range(3):
    seg.attach_pq(cent)
print("ok", int(seg.codes(0, 1000).sum()))
