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
print(f"[rank {rank}] cross-shard exchange: {sh.exchange}", flush=True)
ok = True
# --- sharded PQ training (all-reduce of sums and counts per iteration)
want = orc.pq_train(rows, d, M, K, 5, 42)


def qerr(cent):
    codes = orc.pq_encode_batch(cent, rows, threads=4)
    rec = np.concatenate([cent[s][codes[:, s]] for s in range(M)], axis=1)
    return float(((rows - rec) ** 2).sum(1).mean())


for exact in (True, False):
    cent = sh.pq_train(n, lo, M, K, 5, 42, exact_order=exact)
    c_all = [torch.zeros(cent.size, device=dev) for _ in range(world)]
    dist.all_gather(c_all, torch.from_numpy(cent.reshape(-1)).to(dev))
    same = all(torch.equal(c_all[0], c) for c in c_all)
    if exact:
        good = np.array_equal(cent.view(np.uint32), want.view(np.uint32))
        print(f"[rank {rank}] pq_train sharded, exact order: bit-identical to the reference {good}, identical on all ranks {same}", flush=True)
    else:
        e_got, e_ref = qerr(cent), qerr(want)
        good = abs(e_got - e_ref) <= 0.02 * e_ref
        print(f"[rank {rank}] pq_train sharded, one all-reduce per iteration: quantisation error {e_got:.6f} vs reference "
              f"{e_ref:.6f} ({good}), identical on all ranks {same}", flush=True)
    ok &= good and same
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
    adc_ok = good if i == 0 else (adc_ok and good)
ok &= adc_ok
print(f"[rank {rank}] ADC + re-rank across {world} shards equals the single-segment reference: {adc_ok}", flush=True)
# --- brute force across shards, single query and a batch (tensor-core nomination per shard)
vs.set_option("batch_min_rows", 1)
bf_ok = True
for batch in (qs[:1], qs):
    bi, bs, bc = sh.bruteforce_topk(batch, k)
    for i in range(batch.shape[0]):
        oi, os_, _ = orc.bruteforce_topk(rows, batch[i], k, threads=4)
        good = np.array_equal(bi[i], oi) and np.array_equal(bs[i].view(np.uint64), os_.view(np.uint64))
        bf_ok &= good
ok &= bf_ok
print(f"[rank {rank}] brute force across shards equals the reference: {bf_ok}", flush=True)
t = torch.tensor([1 if ok else 0], device=dev)
dist.all_reduce(t, op=dist.ReduceOp.MIN)
seg.free()
dist.barrier()
dist.destroy_process_group()
if rank == 0:
    print("SHARDED CHECK", "PASSED" if int(t.item()) == 1 else "FAILED", flush=True)
sys.exit(0 if int(t.item()) == 1 else 1)
