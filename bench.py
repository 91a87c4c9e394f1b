#!/usr/bin/env python
"""bench.py -- headline benchmark of the scoring hot path (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W [--impl reference] [--workload c2|c4]

Workload at N=1 (BASELINE.json configs[1], "C2"): exact L2 brute-force top-10 over 1M x 128 fp32
synthetic vectors, query batch 1.  A STEP is one query scanned over the whole resident segment
(one launch of the scan kernel = 1M distance evaluations = 512 MB of algorithmic HBM traffic; the
segment is 4x the 126 MB L2, so every step streams from HBM).  `value` is distance-evals/s with
the segment and the queries resident in HBM; `e2e` is the same metric through the C-ABI host call
(query copied host->device and ids/scores device->host inside every step).

N > 1 (weak scaling): every rank holds its own 1M-row range of an N x 1M corpus; a step scans all
ranks' rows for one query, all-gathers the per-rank top-10 lists (one NCCL collective) and merges
them.  value = N x 1M evals per step / max-over-ranks device time.

--impl reference times the reference's CPU algorithm (the C oracle port of
J/fdb/FdbVectorIndex.java:676-721; the reference is Java and no JVM exists in this image) on the
box's host cores with all threads, same workload, same metric.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

N_ROWS, DIM, TOPK = 1_000_000, 128, 10
C4_ROWS, C4_M, C4_K, C4_NCAND = 100_000_000, 16, 256, 100
CORPUS_SEED, QUERY_SEED = 42, 43


def host_threads() -> int:
    return len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)


def profiled_traffic(name: str) -> float | None:
    """dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed `ncu --set full`
    summary profiles/<name> (made by tools/ncu_summary.py), or None."""
    p = ROOT / "profiles" / name
    if not p.exists():
        return None
    tot, seen = 0.0, 0
    for line in p.read_text().splitlines():
        for key in ("dram__bytes_read.sum [", "dram__bytes_write.sum ["):
            if line.startswith(key):
                unit = line[len(key):line.index("]")]
                scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit)
                vals = [float(v) for v in line.split(":", 1)[1].split(",")]
                if scale and vals:
                    tot += scale * sum(vals) / len(vals)
                    seen += 1
    return tot if seen == 2 else None


def measured_peaks() -> tuple[float, str]:
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        try:
            return float(json.loads(p.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def measured_tensor_peak() -> tuple[float, str]:
    """Dense 16-bit tensor TFLOP/s (the fp16 nomination GEMM runs at the bf16 rate): burst figure, the kernel is timed alone."""
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        try:
            return float(json.loads(p.read_text())["bf16_tflops"]), "measured (MEASURED_PEAKS.json bf16_tflops, burst)"
        except Exception:
            pass
    return 1670.0, "fallback (B200_PROFILING.md dense bf16)"


# ---------------------------------------------------------------------------------------------------
# clocks: sampled DURING the timed regions with NVML
# ---------------------------------------------------------------------------------------------------
class ClockSampler:
    def __init__(self, index: int):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._thread = None
        try:
            import pynvml

            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = int(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
        except Exception:
            self.nv = None

    def _loop(self):
        nv = self.nv
        names = {"hw_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8),
                 "hw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40),
                 "sw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20),
                 "sw_power_cap": getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4)}
        while not self._stop.is_set():
            try:
                self.samples.append(int(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
                r = int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
                for n, bit in names.items():
                    if r & bit:
                        self.reasons.add(n)
            except Exception:
                pass
            time.sleep(0.002)

    def __enter__(self):
        if self.nv is not None:
            self._stop.clear()
            self._thread = threading.Thread(target=self._loop, daemon=True)
            self._thread.start()
        return self

    def __exit__(self, *a):
        if self._thread is not None:
            self._stop.set()
            self._thread.join()
            self._thread = None

    def summary(self) -> dict:
        return {"sm_mhz": statistics.median(self.samples) if self.samples else None,
                "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(self.samples)}


# ---------------------------------------------------------------------------------------------------
# CPU arm: the oracle port of the reference's brute-force scorer, all host threads
# ---------------------------------------------------------------------------------------------------
def cpu_bruteforce(steps: int, warmup: int, budget_s: float | None):
    """Returns (evals_per_s, ms_per_step, steps_done, threads, sample description)."""
    import numpy as np

    from oracle import pyoracle

    try:
        pyoracle.build(native=True, force=True)  # -march=native for THIS box's cores
        orc = pyoracle.get(native=True)
        kind = "-O3 -march=native"
    except Exception:
        orc = pyoracle.get()
        kind = "-O3 -march=x86-64-v3"
    threads = host_threads()
    rows = orc.gen_rows(CORPUS_SEED, 0, N_ROWS, DIM)
    qs = orc.gen_rows(QUERY_SEED, 0, max(steps + warmup, 1), DIM)
    for i in range(warmup):
        orc.bruteforce_topk(rows, qs[i], TOPK, threads=threads)
    t0 = time.perf_counter()
    done = 0
    for i in range(steps):
        orc.bruteforce_topk(rows, qs[warmup + i], TOPK, threads=threads)
        done += 1
        if budget_s is not None and time.perf_counter() - t0 > budget_s and done >= 3:
            break
    dt = time.perf_counter() - t0
    sample = (f"{done} queries x {N_ROWS} rows x {DIM} dims, C oracle port of the reference's Java loop "
              f"({kind}, OpenMP {threads} threads)")
    return N_ROWS * done / dt, dt / done * 1e3, done, threads, sample


def run_reference(args) -> None:
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    val, ms, done, threads, sample = cpu_bruteforce(args.steps, args.warmup, None)
    line = {
        "impl": "reference", "metric": "distance-evals/s (exact L2 brute-force top-10)", "value": val,
        "unit": "distance-evals/s", "n_gpus": args.gpus, "steps": done, "warmup": args.warmup,
        "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic (java.util.Random(42) nextFloat()*2-1 rows, Random(43) queries)",
        "config": {"workload": f"C2: exact L2 brute-force top-{TOPK} over {N_ROWS}x{DIM} fp32, query batch 1",
                   "rows": N_ROWS, "dim": DIM, "k": TOPK, "query_batch": 1},
        "qps": 1e3 / ms,
        "cpu_baseline": {"value": val, "unit": "distance-evals/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": "distance-evals/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------------------------------
def run_ours(args) -> None:
    import numpy as np
    import torch
    import torch.distributed as dist

    import vectorsearch_b200 as vs
    from vectorsearch_b200 import _lib as L
    from vectorsearch_b200.sharded import ShardedSegment

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    vs.init(local_rank)
    lib = vs.load()
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
        # the scan is a persistent one-CTA-per-SM kernel: leave a few SMs to the all-gather's CTAs, or the
        # collective of query i can only start once the scan of query i+1 has drained
        # (scan_reserve_sms below)

    K, W = args.steps, args.warmup
    hbm_peak, peak_src = measured_peaks()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # resident corpus shard: rows [rank*N, (rank+1)*N) of the Java LCG stream, generated on the device
    seg = vs.Segment.generate(CORPUS_SEED, rank * N_ROWS, N_ROWS, DIM, id_base=rank * N_ROWS)
    sh = ShardedSegment(seg, rank, world)
    # queries: identical on every rank (seeded), uniform in [-1, 1) like the reference's generators
    g = torch.Generator(device="cpu").manual_seed(QUERY_SEED)
    q_host = (torch.rand((W + K, DIM), generator=g, dtype=torch.float32) * 2 - 1).pin_memory()
    q_dev = q_host.to(dev)
    ids = torch.empty((1, TOPK), dtype=torch.int64, device=dev)
    sc = torch.empty((1, TOPK), dtype=torch.float64, device=dev)
    cn = torch.empty((1,), dtype=torch.int32, device=dev)
    stream = torch.cuda.current_stream()
    st = stream.cuda_stream
    q_base, ids_p, sc_p, cn_p = q_dev.data_ptr(), ids.data_ptr(), sc.data_ptr(), cn.data_ptr()

    # Independent queries alternate between two streams (libvsgpu keeps one scratch set per stream): the
    # prologue and the last-CTA merge of one scan overlap the streaming phase of the next.
    slots = [dict(stream=torch.cuda.Stream(device=dev), ids=torch.empty((1, TOPK), dtype=torch.int64, device=dev),
                  sc=torch.empty((1, TOPK), dtype=torch.float64, device=dev), cn=torch.empty((1,), dtype=torch.int32, device=dev))
             for _ in range(2)]
    # SMs the persistent scan leaves to the neighbouring stream: 4 are enough for the prologue / merge of the other
    # query and for the peer exchange's two small kernels; an NCCL all-gather needs room for its own CTAs
    reserve_default = "16" if (world > 1 and "peer" not in sh.exchange) else "4"
    reserve_sms = int(os.environ.get("VS_SCAN_RESERVE", reserve_default))
    vs.set_option("scan_reserve_sms", reserve_sms)

    def step_dev(i: int):
        if world == 1:
            s = slots[i & 1]
            L.check(lib.vs_bruteforce_topk_dev(seg.handle, q_base + i * DIM * 4, 1, TOPK, 0, s["ids"].data_ptr(),
                                               s["sc"].data_ptr(), s["cn"].data_ptr(), s["stream"].cuda_stream))
        else:
            # independent queries alternate between the coordinator's two streams: scan(i + 1) overlaps the
            # all-gather and merge of query i; every step's merged result is complete before the region ends
            sh.bruteforce_topk_pipelined(q_dev[i:i + 1], 1, TOPK, 0)

    clocks = ClockSampler(local_rank)
    # ---- device-resident throughput: W warm-up steps, then EXACTLY K timed steps ----------------------
    def drain():
        if world > 1:
            sh.drain()
        else:
            for s in slots:
                stream.wait_stream(s["stream"])

    for s in slots:
        s["stream"].wait_stream(stream)
    for i in range(W):
        step_dev(i)
    drain()
    barrier()
    launches0 = vs.kernel_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with clocks:
        e0.record(stream)
        for i in range(K):
            step_dev(W + i)
        drain()
        e1.record(stream)
        barrier()
    launches = vs.kernel_launch_count() - launches0
    ms_total = max_over_ranks(e0.elapsed_time(e1))
    ms_per_step = ms_total / K
    value = world * N_ROWS * K / (ms_total * 1e-3)

    # ---- per-launch duration of the scan kernel (roofline), events around every launch ------------------
    # (launched alone it has the whole GPU: all SMs; the reserve only pays off when queries overlap)
    vs.set_option("scan_reserve_sms", 0)
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
    with clocks:
        for i in range(K):
            evs[i][0].record(stream)
            L.check(lib.vs_bruteforce_topk_dev(seg.handle, q_base + (W + i) * DIM * 4, 1, TOPK, 0, ids_p, sc_p, cn_p, st))
            evs[i][1].record(stream)
        torch.cuda.synchronize()
    durs = sorted(a.elapsed_time(b) for a, b in evs)
    kern_ms = sum(durs) / len(durs)
    alg_bytes = N_ROWS * DIM * 4
    achieved = alg_bytes / (kern_ms * 1e-3) / 1e9

    # ---- end to end through the public host API: pinned host query in, ids + scores out, every step -----
    qn = q_host.numpy()
    for i in range(W):
        sh.bruteforce_topk(qn[i], TOPK) if world > 1 else seg.bruteforce_topk(qn[i], TOPK)
    barrier()
    with clocks:
        t0 = time.perf_counter()
        for i in range(K):
            r = sh.bruteforce_topk(qn[W + i], TOPK) if world > 1 else seg.bruteforce_topk(qn[W + i], TOPK)
        barrier()
        e2e_s = time.perf_counter() - t0
    e2e_s = max_over_ranks(e2e_s)
    e2e_value = world * N_ROWS * K / e2e_s
    last_ids = np.asarray(r[0]).tolist()

    # ---- CPU baseline: bounded sample on rank 0 at N=1 only ---------------------------------------------
    cpu = None
    if world == 1 and not args.no_cpu:
        cval, cms, cdone, cthreads, csample = cpu_bruteforce(200, 2, 12.0)
        cpu = {"value": cval, "unit": "distance-evals/s", "cores": cthreads, "kind": "port", "sample": csample,
               "ms_per_query": cms}

    extra = {}
    if world == 1:
        extra["c2_b1024"] = bench_c2_batch(vs, L, lib, torch, dev, seg, min(K, 50), W)
    if args.workload in ("c4", "all") and world == 1:
        extra["c4"] = bench_c4(vs, L, lib, torch, dev, min(K, 50), W, hbm_peak)

    if rank == 0:
        line = {
            "metric": "distance-evals/s (exact L2 brute-force top-10)", "value": value, "unit": "distance-evals/s",
            "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic (java.util.Random(42) nextFloat()*2-1 rows generated on device; seeded uniform queries)",
            "config": {"workload": f"C2: exact L2 brute-force top-{TOPK} over {N_ROWS}x{DIM} fp32 per GPU, query batch 1",
                       "rows_per_gpu": N_ROWS, "dim": DIM, "k": TOPK, "query_batch": 1,
                       "parallelism": (f"row-range shards x{world}, cross-shard top-k exchange per query = {sh.exchange}; "
                                       "independent queries alternate between two streams") if world > 1 else "1 GPU",
                       "l2_policy": "input 512 MB per step > 126 MB L2 (no flush needed)",
                       "pipelining": "independent queries alternate between two CUDA streams; the scan leaves "
                                     f"{reserve_sms} SMs free so that the neighbouring query's "
                                     "prologue / merge (and the cross-shard exchange at N > 1) run beside it"},
            "qps": 1e3 / ms_per_step,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s",
                         "frac": achieved / hbm_peak, "traffic": profiled_traffic("r1_c2_scan_full.txt"),
                         "traffic_source": "profiles/r1_c2_scan_full.txt (ncu --set full, dram read+write per launch)",
                         "peak_source": peak_src,
                         "kernel": "scan_tma_kernel<TPR=4,U=2,L2,WarpTopKReg> (K1)", "kernel_ms": kern_ms,
                         "kernel_ms_median": durs[len(durs) // 2], "algorithmic_bytes_per_launch": alg_bytes,
                         "achieved_in_step": alg_bytes / (ms_per_step * 1e-3) / 1e9,
                         "note": "achieved = bytes / duration of the kernel launched ALONE (events around every launch); "
                                 "achieved_in_step = bytes / ms_per_step of the timed region, where consecutive queries overlap "
                                 "on two streams (a read-only stream can exceed the read+write copy rate used as peak)"},
            "e2e": {"value": e2e_value, "unit": "distance-evals/s", "h2d_bytes_per_step": DIM * 4,
                    "d2h_bytes_per_step": TOPK * 16 + 4, "ms_per_step": e2e_s / K * 1e3, "qps": K / e2e_s},
            "gpu_launches": int(launches),
            "clocks": clocks.summary(),
            "top10_last_query": last_ids,
        }
        if cpu is not None:
            line["cpu_baseline"] = cpu
        if extra:
            line["extra"] = extra
        print(json.dumps(line), flush=True)
    seg.free()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def bench_c2_batch(vs, L, lib, torch, dev, seg, K, W, nq=1024):
    """C2, query batch 1024: one step = 1024 queries against the resident 1M x 128 segment (batch.cu:
    tcgen05 fp16 nomination GEMM + exact re-score).  Tensor-bound: 2 * N * D * B flop per batch."""
    import numpy as np

    tpeak, tsrc = measured_tensor_peak()
    g = torch.Generator(device="cpu").manual_seed(QUERY_SEED + 1)
    q_host = (torch.rand((nq, DIM), generator=g, dtype=torch.float32) * 2 - 1).pin_memory()
    q_dev = q_host.to(dev)
    ids = torch.empty((nq, TOPK), dtype=torch.int64, device=dev)
    sc = torch.empty((nq, TOPK), dtype=torch.float64, device=dev)
    cn = torch.empty((nq,), dtype=torch.int32, device=dev)
    stream = torch.cuda.current_stream()
    st = stream.cuda_stream

    def step():
        L.check(lib.vs_bruteforce_topk_dev(seg.handle, q_dev.data_ptr(), nq, TOPK, 0, ids.data_ptr(), sc.data_ptr(), cn.data_ptr(), st))

    for _ in range(max(W, 3)):
        step()
    torch.cuda.synchronize()
    l0 = vs.kernel_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(K):
        step()
    e1.record(stream)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / K
    launches = vs.kernel_launch_count() - l0
    qn = q_host.numpy()
    for _ in range(3):
        seg.bruteforce_topk(qn, TOPK)
    t0 = time.perf_counter()
    for _ in range(K):
        r = seg.bruteforce_topk(qn, TOPK)
    e2e_ms = (time.perf_counter() - t0) / K * 1e3
    flops = 2.0 * N_ROWS * DIM * nq
    tf = flops / (ms * 1e-3) / 1e12
    # the same 1024 queries one at a time through the streaming scan must give the same lists (spot check)
    same = all(np.array_equal(seg.bruteforce_topk(qn[i], TOPK)[0], r[0][i]) for i in (0, 511, 1023))
    return {"workload": f"C2: exact L2 brute-force top-{TOPK} over {N_ROWS}x{DIM} fp32, query batch {nq}, 1 GPU",
            "ms_per_batch": ms, "qps": nq / (ms * 1e-3), "distance_evals_per_s": N_ROWS * nq / (ms * 1e-3),
            "roofline": {"bound": "tensor", "achieved": tf, "peak": tpeak, "unit": "TFLOP/s", "frac": tf / tpeak,
                         "peak_source": tsrc, "algorithmic_flops_per_batch": flops,
                         "kernel": "batch_gemm_kernel<STAT,HALF,L2,G64> (K2, tcgen05 kind::f16) + batch_select + fallback; "
                                   "achieved is over the WHOLE batch, the GEMM alone is the share profiles/r1_c2_b1024_launches.txt shows",
                         "traffic": None},
            "e2e": {"ms_per_batch": e2e_ms, "qps": nq / (e2e_ms * 1e-3), "distance_evals_per_s": N_ROWS * nq / (e2e_ms * 1e-3),
                    "h2d_bytes_per_step": nq * DIM * 4, "d2h_bytes_per_step": nq * (TOPK * 16 + 4)},
            "gpu_launches": int(launches), "matches_per_query_scan": bool(same)}


def bench_c4(vs, L, lib, torch, dev, K, W, hbm_peak):
    """C4 (1 GPU): PQ ADC top-100 scan + exact re-rank to top-10 over 100M x 128 (M=16 codes)."""
    n = int(os.environ.get("VS_C4_ROWS", C4_ROWS))
    t0 = time.perf_counter()
    seg = vs.Segment.generate(CORPUS_SEED, 0, n, DIM)
    torch.cuda.synchronize()
    t_gen = time.perf_counter() - t0
    t0 = time.perf_counter()
    # codebook trained on the first 1M rows (PqTrainer.train(..., 5, 42)), codes for all rows on device
    train = vs.Segment.generate(CORPUS_SEED, 0, min(n, 1_000_000), DIM)
    cent = vs.PqTrainer.train(None, DIM, C4_M, C4_K, 5, 42, segment=train)
    train.free()
    t_train = time.perf_counter() - t0
    t0 = time.perf_counter()
    seg.attach_pq(cent)
    t_enc = time.perf_counter() - t0
    g = torch.Generator(device="cpu").manual_seed(QUERY_SEED)
    q_dev = (torch.rand((W + K, DIM), generator=g, dtype=torch.float32) * 2 - 1).to(dev)
    ids = torch.empty((1, TOPK), dtype=torch.int64, device=dev)
    sc = torch.empty((1, TOPK), dtype=torch.float64, device=dev)
    cn = torch.empty((1,), dtype=torch.int32, device=dev)
    st = torch.cuda.current_stream().cuda_stream
    for i in range(W):
        L.check(lib.vs_adc_rerank_topk_dev(seg.handle, q_dev[i].data_ptr(), 1, C4_NCAND, TOPK, 0, 0, ids.data_ptr(),
                                           sc.data_ptr(), cn.data_ptr(), st))
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(K):
        L.check(lib.vs_adc_rerank_topk_dev(seg.handle, q_dev[W + i].data_ptr(), 1, C4_NCAND, TOPK, 0, 0,
                                           ids.data_ptr(), sc.data_ptr(), cn.data_ptr(), st))
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / K
    # the ADC scan alone (LUT build + fast scan + conditional fallback launch), for its roofline
    ids100 = torch.empty((1, C4_NCAND), dtype=torch.int64, device=dev)
    ap100 = torch.empty((1, C4_NCAND), dtype=torch.float64, device=dev)
    e0.record()
    for i in range(K):
        L.check(lib.vs_adc_topk_dev(seg.handle, q_dev[W + i].data_ptr(), 1, C4_NCAND, ids100.data_ptr(),
                                    ap100.data_ptr(), cn.data_ptr(), st))
    e1.record()
    torch.cuda.synchronize()
    ms_adc = e0.elapsed_time(e1) / K
    gbs = n * C4_M / (ms_adc * 1e-3) / 1e9
    # end to end through the host API: query from host memory in, ids + scores out, every query
    qh = q_dev.cpu().numpy()
    for i in range(3):
        seg.adc_rerank_topk(qh[i], C4_NCAND, TOPK)
    t0 = time.perf_counter()
    for i in range(K):
        seg.adc_rerank_topk(qh[W + i], C4_NCAND, TOPK)
    e2e_ms = (time.perf_counter() - t0) / K * 1e3
    out = {"workload": f"C4: ADC top-{C4_NCAND} + exact re-rank top-{TOPK} over {n}x{DIM} (M={C4_M} uint8 codes), 1 GPU",
           "ms_per_query": ms, "adc_evals_per_s": n / (ms * 1e-3), "qps": 1e3 / ms,
           "e2e": {"ms_per_query": e2e_ms, "adc_evals_per_s": n / (e2e_ms * 1e-3), "h2d_bytes_per_step": DIM * 4,
                   "d2h_bytes_per_step": TOPK * 16 + 4},
           "roofline": {"bound": "hbm", "achieved": gbs, "peak": hbm_peak, "unit": "GB/s", "frac": gbs / hbm_peak,
                        "kernel": "build_lut_mm + adc_fastscan_kernel<4> (K5+K6)", "ms_per_launch_group": ms_adc,
                        "algorithmic_bytes_per_launch": n * C4_M, "traffic": None},
           "generate_s": t_gen, "train_1M_5iters_s": t_train, "encode_s": t_enc,
           "encode_vectors_per_s": n / t_enc}
    seg.free()
    return out


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="all", choices=["c2", "c4", "all"],
                    help="c2 = the headline line only; c4/all also time C4 (ADC + re-rank, 100M rows) into `extra`")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
