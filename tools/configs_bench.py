"""BASELINE.json configs C2 (batch 1024), C3, C4 and C5 across the GPUs of one box: one rank per GPU, rows sharded
by ascending range, device-timed (CUDA events, max over ranks).  One JSON line per config on rank 0.

    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29544 tools/configs_bench.py [c2b c3 c4 c5]
    python tools/configs_bench.py c3 c4            # one GPU
"""
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vectorsearch_b200 as vs
from vectorsearch_b200.sharded import ShardedSegment, shard_range

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
vs.init(local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
which = [a for a in sys.argv[1:] if not a.startswith("-")] or ["c2b", "c3", "c4", "c5"]
SCALE = float(os.environ.get("VS_SCALE", "1"))  # shrink every corpus (development)


def barrier():
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()


def maxr(x):
    if world == 1:
        return x
    t = torch.tensor([x], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def same_on_all_ranks(t):
    if world == 1:
        return True
    g = [torch.zeros_like(t) for _ in range(world)]
    dist.all_gather(g, t.contiguous())
    return all(torch.equal(g[0], x) for x in g)


def timed(fn, iters, warm=3):
    for _ in range(warm):
        fn()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    barrier()
    return maxr(e0.elapsed_time(e1) / iters)


def emit(line):
    if rank == 0:
        line["n_gpus"] = world
        print(json.dumps(line), flush=True)


def queries(nq, d, seed):
    g = torch.Generator(device="cpu").manual_seed(seed)
    return (torch.rand((nq, d), generator=g, dtype=torch.float32) * 2 - 1).to(dev)


if "c2b" in which:  # weak scaling: 1M x 128 per GPU, 1024 queries per batch, top-10 L2
    n_per, d, nq, k = int(1_000_000 * SCALE), 128, 1024, 10
    seg = vs.Segment.generate(42, rank * n_per, n_per, d, id_base=rank * n_per)
    sh = ShardedSegment(seg, rank, world)
    q = queries(nq, d, 43)
    ms = timed(lambda: sh.bruteforce_topk_dev(q, nq, k, 0), 20)
    ids, sc, cn = sh.bruteforce_topk_dev(q, nq, k, 0)
    emit({"config": "C2 batch 1024 (weak: 1M x 128 per GPU), exact L2 top-10", "ms_per_batch": ms, "qps": nq / ms * 1e3,
          "distance_evals_per_s": world * n_per * nq / ms * 1e3, "identical_on_all_ranks": same_on_all_ranks(ids)})
    seg.free()

if "c3" in which:  # strong: 10M x 128 in total, PqTrainer.train(5 iterations, seed 42) + encode, M 16, K 256
    n, d, M, K = int(10_000_000 * SCALE), 128, 16, 256
    lo, hi = shard_range(n, rank, world)
    seg = vs.Segment.generate(42, lo, hi - lo, d, id_base=lo)
    sh = ShardedSegment(seg, rank, world)
    out = {}
    for name, exact in (("exact_order", True), ("one_allreduce", False)):
        sh.pq_train(n, lo, M, K, 1, 42, exact_order=exact)  # warm-up (allocations, NCCL channels)
        barrier()
        t0 = time.perf_counter()
        cent = sh.pq_train(n, lo, M, K, 5, 42, exact_order=exact)
        barrier()
        out[name + "_train_s"] = maxr(time.perf_counter() - t0)
        out[name + "_identical_on_all_ranks"] = same_on_all_ranks(torch.from_numpy(cent.reshape(-1)).to(dev))
    barrier()
    t0 = time.perf_counter()
    seg.attach_pq(cent)
    barrier()
    enc = maxr(time.perf_counter() - t0)
    out.update({"config": f"C3: PQ k-means train (5 Lloyd iterations) + encode, M=16 K=256 over {n}x128 in total", "encode_s": enc,
                "train_vectors_per_s_exact_order": 5 * n / out["exact_order_train_s"],
                "train_vectors_per_s_one_allreduce": 5 * n / out["one_allreduce_train_s"], "encode_vectors_per_s": n / enc})
    emit(out)
    seg.free()

if "c4" in which:  # strong: 100M x 128 in total, ADC top-100 + exact re-rank to top-10
    n, d, M, K, n_cand, k = int(100_000_000 * SCALE), 128, 16, 256, 100, 10
    lo, hi = shard_range(n, rank, world)
    seg = vs.Segment.generate(42, lo, hi - lo, d, id_base=lo)
    tr = vs.Segment.generate(42, 0, min(n, 1_000_000), d)
    cent = vs.PqTrainer.train(None, d, M, K, 5, 42, segment=tr)  # same codebook on every rank
    tr.free()
    seg.attach_pq(cent)
    sh = ShardedSegment(seg, rank, world)
    q = queries(64, d, 43)
    i = [0]

    def step():
        sh.adc_rerank_topk_dev(q[i[0] % 64:i[0] % 64 + 1], 1, n_cand, k)
        i[0] += 1

    ms = timed(step, 50)
    ids, sc, cn = sh.adc_rerank_topk_dev(q[:1], 1, n_cand, k)
    emit({"config": f"C4: PQ ADC top-100 + exact re-rank top-10 over {n}x128 in total (M=16 codes)", "ms_per_query": ms,
          "qps": 1e3 / ms, "adc_evals_per_s": n / ms * 1e3, "hbm_gbs_per_gpu": (hi - lo) * M / ms / 1e6,
          "identical_on_all_ranks": same_on_all_ranks(ids)})
    seg.free()

if "c5" in which:  # strong: 50M x 768 in total, cosine top-50, 256 queries per batch
    n, d, nq, k = int(50_000_000 * SCALE), 768, 256, 50
    lo, hi = shard_range(n, rank, world)
    need = (hi - lo) * d * 6 + (4 << 30)
    free_b = torch.cuda.mem_get_info()[0]
    if need > free_b:
        emit({"config": f"C5: cosine top-50 over {n}x768, batch 256", "skipped": f"needs {need >> 30} GiB per GPU, {free_b >> 30} free"})
    else:
        seg = vs.Segment.generate(42, lo, hi - lo, d, id_base=lo)
        sh = ShardedSegment(seg, rank, world)
        q = queries(nq, d, 43)
        probe = min(n - 1, 12_345_678)            # a corpus row as query 0: it must find itself with similarity 1
        owner = [r for r in range(world) if shard_range(n, r, world)[0] <= probe < shard_range(n, r, world)[1]][0]
        row = torch.from_numpy(seg.rows(probe - lo, 1)[0]).to(dev) if rank == owner else torch.zeros(d, device=dev)
        if world > 1:
            dist.broadcast(row, owner)
        q[0] = row
        ms = timed(lambda: sh.bruteforce_topk_dev(q, nq, k, 1), 10)
        ids, sc, cn = sh.bruteforce_topk_dev(q, nq, k, 1)
        torch.cuda.synchronize()
        emit({"config": f"C5: cosine brute-force top-50 over {n}x768 fp32 in total, query batch 256", "ms_per_batch": ms,
              "qps": nq / ms * 1e3, "distance_evals_per_s": n * nq / ms * 1e3,
              "tflops": 2.0 * n * d * nq / ms / 1e9, "probe_found_itself": bool(ids[0, 0].item() == probe and abs(sc[0, 0].item() - 1.0) < 1e-6),
              "identical_on_all_ranks": same_on_all_ranks(ids)})
        seg.free()

if world > 1:
    dist.barrier()
    dist.destroy_process_group()
