"""Real multi-GPU check (torchrun, one rank per GPU): brute force, ADC + re-rank and sharded PQ training over
row-range shards against the CPU oracle on the whole corpus.  Run on the GPU box:
    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tools/sharded_check.py
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vectorsearch_b200 as vs
from oracle import pyoracle
from vectorsearch_b200.sharded import ShardedSegment, shard_range

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
vs.init(local)
dist.init_process_group("nccl", device_id=dev)
orc = pyoracle.get()
n, d, M, K, n_cand, k, nq = 80000, 64, 8, 64, 100, 10, 4
rows = orc.gen_rows(42, 0, n, d)
rows[500:540] = rows[17]
rows[n - 40:n - 10] = rows[17]
qs = orc.gen_rows(43, 0, nq, d)
qs[0] = rows[17]
lo, hi = shard_range(n, rank, world)
seg = vs.Segment.upload(rows[lo:hi], id_base=lo)
sh = ShardedSegment(seg, rank, world)
ok = True
# --- sharded PQ training (all-reduce of sums and counts per iteration)
cent = sh.pq_train(n, lo, M, K, 5, 42)
want = orc.pq_train(rows, d, M, K, 5, 42)
close = np.allclose(cent, want, rtol=2e-5, atol=1e-6)
c_all = [torch.zeros(cent.size, device=dev) for _ in range(world)]
dist.all_gather(c_all, torch.from_numpy(cent.reshape(-1)).to(dev))
same = all(torch.equal(c_all[0], c) for c in c_all)
print(f"[rank {rank}] pq_train sharded: close to reference {close}, identical on all ranks {same}, "
      f"max rel diff {np.max(np.abs(cent - want) / (np.abs(want) + 1e-6)):.2e}", flush=True)
ok &= close and same
# --- ADC + re-rank across shards (codes from the reference centroids so that the lists are comparable)
codes = orc.pq_encode_batch(want, rows, threads=4)
seg.attach_pq(want, codes[lo:hi])
q_dev = torch.from_numpy(qs).to(dev)
ids, sc, cn = sh.adc_rerank_topk_dev(q_dev, nq, n_cand, k)
torch.cuda.synchronize()
ids, sc = ids.cpu().numpy(), sc.cpu().numpy()
for i in range(nq):
    ci, _ = orc.adc_topn(orc.build_lut(want, qs[i]), codes, n_cand)
    ri, rs, _ = orc.rerank_topk(rows, qs[i], ci, k)
    good = np.array_equal(ids[i], ri) and np.array_equal(sc[i].view(np.uint64), rs.view(np.uint64))
    ok &= good
print(f"[rank {rank}] ADC + re-rank across {world} shards equals the single-segment reference: {ok}", flush=True)
# --- brute force across shards, single query and a batch (tensor-core nomination per shard)
vs.set_option("batch_min_rows", 1)
for batch in (qs[:1], qs):
    bi, bs, bc = sh.bruteforce_topk(batch, k)
    for i in range(batch.shape[0]):
        oi, os_, _ = orc.bruteforce_topk(rows, batch[i], k, threads=4)
        good = np.array_equal(bi[i], oi) and np.array_equal(bs[i].view(np.uint64), os_.view(np.uint64))
        ok &= good
print(f"[rank {rank}] brute force across shards equals the reference: {ok}", flush=True)
t = torch.tensor([1 if ok else 0], device=dev)
dist.all_reduce(t, op=dist.ReduceOp.MIN)
seg.free()
dist.barrier()
dist.destroy_process_group()
if rank == 0:
    print("SHARDED CHECK", "PASSED" if int(t.item()) == 1 else "FAILED", flush=True)
sys.exit(0 if int(t.item()) == 1 else 1)
