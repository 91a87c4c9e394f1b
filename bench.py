#!/usr/bin/env python
"""bench.py -- headline benchmark of the scoring hot path (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W [--impl reference] [--workload c2|all]

Workload (BASELINE.json configs[1], "C2"): exact L2 brute-force top-10 over 1M x 128 fp32 synthetic vectors per
GPU, query batch 1.  A STEP is one query scanned over the whole resident corpus: at N GPUs every rank holds its
own 1M-row range of an N x 1M corpus (weak scaling), scans it (one launch of the scan kernel = 1M distance
evaluations = 512 MB of algorithmic HBM traffic; the shard is 4x the 126 MB L2, so every step streams from
HBM), pushes its top-10 into every peer's buffer over NVLink and merges the N lists.  `value` is
distance-evals/s with the corpus and the queries resident in HBM; `e2e` is the same metric through the C-ABI
host call (query copied host->device and ids/scores device->host inside every step).

After the timed regions rank 0 re-scans all N x 1M rows of the last query with the CPU oracle and asserts the
merged ids and scores ("parity_checked"); a mismatch fails the run.

`extra` carries the other BASELINE configs at this N: C2 with query batch 1024 (N = 1), C3 (PQ train + encode
over 10M x 128 in total) and C4 (ADC top-100 + re-rank to top-10 over 100M x 128 in total, strong scaling).

--impl reference times the reference's CPU algorithm (the C oracle port of J/fdb/FdbVectorIndex.java:676-721;
the reference is Java and no JVM exists in this image) on the box's host cores with all threads, same
workload (N x 1M rows per step), same query stream, same metric and config.
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
C3_ROWS, C4_ROWS, PQ_M, PQ_K, C4_NCAND = 10_000_000, 100_000_000, 16, 256, 100
CORPUS_SEED, QUERY_SEED = 42, 43
METRIC = "distance-evals/s (exact L2 brute-force top-10)"
DATA = "synthetic (java.util.Random(42) nextFloat()*2-1 rows, Random(43) queries; the reference's own generator)"


def config_for(world: int) -> dict:
    """The same dict in both arms (the driver compares them)."""
    return {"workload": f"C2: exact L2 brute-force top-{TOPK} over {N_ROWS}x{DIM} fp32 per GPU, query batch 1",
            "rows_per_gpu": N_ROWS, "rows_total": N_ROWS * world, "dim": DIM, "k": TOPK, "query_batch": 1,
            "l2_policy": "input 512 MB per GPU per step > 126 MB L2 (no flush needed)"}


def host_threads() -> int:
    return len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)


def profiled_traffic(*names: str):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch from the first committed `ncu --set full` summary
    among profiles/<names> (made by tools/ncu_summary.py) -> (bytes, file) or (None, None)."""
    for name in names:
        p = ROOT / "profiles" / name
        if not p.exists():
            continue
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
        if seen == 2:
            return tot, f"profiles/{name}"
    return None, None


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


def sustained_tensor_peak():
    """The same GEMM back to back for seconds (MEASURED_PEAKS.json bf16_tflops_sustained) or None."""
    try:
        return float(json.loads((ROOT / "MEASURED_PEAKS.json").read_text())["bf16_tflops_sustained"])
    except Exception:
        return None


# ---------------------------------------------------------------------------------------------------
# clocks: sampled DURING the timed regions with NVML (the thread runs for the whole process; only
# samples taken while a region is open are kept, so opening a region costs nothing on the host)
# ---------------------------------------------------------------------------------------------------
class ClockSampler:
    """SM clock and throttle reasons of one GPU, sampled through NVML by a thread while a timed region is open.
    Only rank 0 samples (enabled): eight processes polling NVML at once serialise on the driver and slow every rank's
    kernel launches -- measured at N = 8: 64 us of host time per step with eight samplers, 19 us without."""

    def __init__(self, index: int, enabled: bool = True):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._active = False
        self.period_s = float(os.environ.get("VS_BENCH_SAMPLER_MS", "0.2")) * 1e-3
        self._stop = threading.Event()
        self._thread = None
        try:
            if not enabled:
                raise RuntimeError("sampling is rank 0's job")
            import pynvml

            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = int(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self._thread = threading.Thread(target=self._loop, daemon=True)
            self._thread.start()
        except Exception:
            self.nv = None

    def _loop(self):
        nv = self.nv
        names = {"hw_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8),
                 "hw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40),
                 "sw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20),
                 "sw_power_cap": getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4)}
        while not self._stop.is_set():
            if self._active:
                try:
                    self.samples.append(int(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
                    r = int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
                    for n, bit in names.items():
                        if r & bit:
                            self.reasons.add(n)
                except Exception:
                    pass
            time.sleep(self.period_s)

    def __enter__(self):
        self._active = True
        return self

    def __exit__(self, *a):
        self._active = False

    def close(self):
        self._stop.set()

    def summary(self) -> dict:
        return {"sm_mhz": statistics.median(self.samples) if self.samples else None,
                "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(self.samples)}


# ---------------------------------------------------------------------------------------------------
# CPU arm: the oracle port of the reference's brute-force scorer
# ---------------------------------------------------------------------------------------------------
def _native_oracle():
    from oracle import pyoracle

    try:
        pyoracle.build(native=True, force=True)  # -march=native for THIS box's cores
        return pyoracle.get(native=True), "-O3 -march=native"
    except Exception:
        return pyoracle.get(), "-O3 -march=x86-64-v3"


def cpu_bruteforce(world: int, steps: int, warmup: int, budget_s: float | None, threads: int | None = None):
    """The reference's segment scorer over world x 1M rows per step (searchBruteForceSegment per 1M-row segment,
    merged as J/fdb/FdbVectorIndex.java:432-437).  Returns (evals_per_s, ms_per_step, steps_done, threads, sample)."""
    import numpy as np

    orc, kind = _native_oracle()
    threads = threads or host_threads()
    segs = [orc.gen_rows(CORPUS_SEED, r * N_ROWS, N_ROWS, DIM) for r in range(world)]
    qs = orc.gen_rows(QUERY_SEED, 0, max(steps + warmup, 1), DIM)

    def step(q):
        ids, scs = [], []
        for r, rows in enumerate(segs):
            i, s, _ = orc.bruteforce_topk(rows, q, TOPK, threads=threads)
            ids.append(i + r * N_ROWS)
            scs.append(s)
        return orc.merge_topk(np.concatenate(ids), np.concatenate(scs), TOPK)

    for i in range(warmup):
        step(qs[i])
    t0 = time.perf_counter()
    done = 0
    for i in range(steps):
        step(qs[warmup + i])
        done += 1
        if budget_s is not None and time.perf_counter() - t0 > budget_s and done >= 3:
            break
    dt = time.perf_counter() - t0
    sample = (f"{done} queries x {world} x {N_ROWS} rows x {DIM} dims, C oracle port of the reference's Java loop "
              f"({kind}, {'OpenMP ' + str(threads) + ' threads' if threads > 1 else 'single thread, as the reference scores a segment'})")
    return world * N_ROWS * done / dt, dt / done * 1e3, done, threads, sample


def run_reference(args) -> None:
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    world = max(1, args.gpus)
    val, ms, done, threads, sample = cpu_bruteforce(world, args.steps, args.warmup, None)
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": "distance-evals/s", "n_gpus": args.gpus,
        "steps": done, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": DATA, "config": config_for(world), "qps": 1e3 / ms,
        "cpu_baseline": {"value": val, "unit": "distance-evals/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": "distance-evals/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def oracle_check_c2(world: int, q, ids, scores) -> None:
    """Rank 0: the last query over all world x 1M rows with the oracle; raises on any difference."""
    import numpy as np

    from oracle import pyoracle

    orc = pyoracle.get()
    oi, os_ = [], []
    for r in range(world):
        rows = orc.gen_rows(CORPUS_SEED, r * N_ROWS, N_ROWS, DIM)
        i, s, _ = orc.bruteforce_topk(rows, q, TOPK, threads=host_threads())
        oi.append(i + r * N_ROWS)
        os_.append(s)
    wi, ws = orc.merge_topk(np.concatenate(oi), np.concatenate(os_), TOPK)
    if not (np.array_equal(np.asarray(ids), wi) and np.array_equal(np.asarray(scores).view(np.uint64), ws.view(np.uint64))):
        raise SystemExit(f"PARITY FAILURE (C2, {world} GPUs): got {list(ids)} {list(scores)}, oracle {wi.tolist()} {ws.tolist()}")


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

    K, W = args.steps, args.warmup
    hbm_peak, peak_src = measured_peaks()
    clocks = ClockSampler(local_rank, enabled=(rank == 0))

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
    n_slots = int(os.environ.get("VS_BENCH_SLOTS", "3"))
    sh = ShardedSegment(seg, rank, world, slots=n_slots)
    # queries: draws of java.util.Random(43), the stream the reference arm uses, generated by libvsgpu's own generator
    qseg = vs.Segment.generate(QUERY_SEED, 0, W + K + 1, DIM)
    qn = qseg.rows()
    qseg.free()
    q_host = torch.from_numpy(qn).pin_memory()
    q_dev = q_host.to(dev)
    ids = torch.empty((1, TOPK), dtype=torch.int64, device=dev)
    sc = torch.empty((1, TOPK), dtype=torch.float64, device=dev)
    cn = torch.empty((1,), dtype=torch.int32, device=dev)
    stream = torch.cuda.current_stream()
    st = stream.cuda_stream
    q_base, ids_p, sc_p, cn_p = q_dev.data_ptr(), ids.data_ptr(), sc.data_ptr(), cn.data_ptr()

    # Independent queries alternate between the coordinator's streams (libvsgpu keeps one scratch set and one
    # exchange ring per stream): the prologue, the last-CTA merge and the cross-shard exchange of one query overlap
    # the streaming phase of the next.  ONE C call per step at every N.
    slots = [dict(stream=s_, ids=torch.empty((1, TOPK), dtype=torch.int64, device=dev),
                  sc=torch.empty((1, TOPK), dtype=torch.float64, device=dev), cn=torch.empty((1,), dtype=torch.int32, device=dev))
             for s_ in sh._streams]
    # SMs the persistent scan leaves to the neighbouring stream: enough for the prologue / merge of the other query
    # and for the peer exchange's small kernels; an NCCL all-gather (fallback path) needs room for its own CTAs
    reserve_default = "16" if (world > 1 and sh._comm is None) else "4"
    reserve_sms = int(os.environ.get("VS_SCAN_RESERVE", reserve_default))
    vs.set_option("scan_reserve_sms", reserve_sms)
    comm = sh._comm

    def step_dev(i: int):
        s = slots[i % len(slots)]
        if world == 1:
            L.check(lib.vs_bruteforce_topk_dev(seg.handle, q_base + i * DIM * 4, 1, TOPK, 0, s["ids"].data_ptr(),
                                               s["sc"].data_ptr(), s["cn"].data_ptr(), s["stream"].cuda_stream))
        elif comm is not None:
            L.check(lib.vs_bruteforce_topk_exchange_dev(seg.handle, comm, q_base + i * DIM * 4, 1, TOPK, 0, s["ids"].data_ptr(),
                                                        s["sc"].data_ptr(), s["cn"].data_ptr(), s["stream"].cuda_stream))
        else:
            sh.bruteforce_topk_pipelined(q_dev[i:i + 1], 1, TOPK, 0)

    def drain():
        if world > 1 and comm is None:
            sh.drain()
        for s in slots:
            stream.wait_stream(s["stream"])

    # ---- device-resident throughput: W warm-up steps, then EXACTLY K timed steps ----------------------
    for s in slots:
        s["stream"].wait_stream(stream)
    for i in range(W):
        step_dev(i)
    drain()
    barrier()
    # One untimed lock-step exchange right before e0: its merge kernel waits ON THE DEVICE for every rank's list, so
    # all GPUs leave it together and e0 is recorded on every rank when the slowest host has arrived -- the host skew
    # left by the barrier (tens to hundreds of microseconds) stays out of a 20-step window.
    if world > 1 and comm is not None:
        for s in slots:
            L.check(lib.vs_bruteforce_topk_exchange_dev(seg.handle, comm, q_base + (W + K) * DIM * 4, 1, TOPK, 0, s["ids"].data_ptr(),
                                                        s["sc"].data_ptr(), s["cn"].data_ptr(), s["stream"].cuda_stream))
        drain()
    launches0 = vs.kernel_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with clocks:
        e0.record(stream)
        for s in slots:
            s["stream"].wait_stream(stream)
        t_issue = time.perf_counter()
        for i in range(K):
            step_dev(W + i)
        issue_us = (time.perf_counter() - t_issue) / K * 1e6  # host time to enqueue one step (must stay below ms_per_step)
        drain()
        e1.record(stream)
        barrier()
    launches = vs.kernel_launch_count() - launches0
    ms_total = max_over_ranks(e0.elapsed_time(e1))
    ms_per_step = ms_total / K
    value = world * N_ROWS * K / (ms_total * 1e-3)
    last = slots[(W + K - 1) % len(slots)]
    dev_last = (last["ids"].cpu().numpy()[0].copy(), last["sc"].cpu().numpy()[0].copy())

    # ---- per-launch duration of the scan kernel (roofline), events around every launch ------------------
    # (launched alone it has the whole GPU: all SMs; the reserve only pays off when queries overlap)
    def alone_ms():
        vs.set_option("scan_reserve_sms", 0)
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
        for i in range(3):  # (a changed option makes the first call plan its launch anew)
            L.check(lib.vs_bruteforce_topk_dev(seg.handle, q_base + i * DIM * 4, 1, TOPK, 0, ids_p, sc_p, cn_p, st))
        torch.cuda.synchronize()
        l0 = vs.kernel_launch_count()
        with clocks:
            for i in range(K):
                evs[i][0].record(stream)
                L.check(lib.vs_bruteforce_topk_dev(seg.handle, q_base + (W + i) * DIM * 4, 1, TOPK, 0, ids_p, sc_p, cn_p, st))
                evs[i][1].record(stream)
            torch.cuda.synchronize()
        vs.set_option("scan_reserve_sms", reserve_sms)
        d_ = sorted(a.elapsed_time(b) for a, b in evs)
        return sum(d_) / len(d_), d_[len(d_) // 2], (vs.kernel_launch_count() - l0) / K

    kern_ms, kern_med, launches_per_call = alone_ms()
    # Which kernel answers a single query: by default the scan of the fp16 operand copy (scan_half_kernel, 2 launches per
    # call: the scan and the conditional exact fallback); the fp32 streaming scan (scan_tma_kernel, 1 launch) when the copy
    # is switched off or does not exist.  The roofline uses the bytes THAT kernel has to read.
    half_path = launches_per_call > 1.5
    dp = (DIM + 7) // 8 * 8
    alg_bytes = N_ROWS * (dp * 2 + 4) if half_path else N_ROWS * DIM * 4
    achieved = alg_bytes / (kern_ms * 1e-3) / 1e9
    fp32_leg = None
    if half_path and world == 1:
        # the same query stream through the fp32 scan (option scan_fp16 = 0): SURVEY 8(d)'s 512 B per evaluation
        vs.set_option("scan_fp16", 0)
        f_ms, f_med, _ = alone_ms()
        for i in range(W):
            step_dev(i)
        drain()
        torch.cuda.synchronize()
        f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        f0.record(stream)
        for s in slots:
            s["stream"].wait_stream(stream)
        for i in range(K):
            step_dev(W + i)
        drain()
        f1.record(stream)
        torch.cuda.synchronize()
        f_step = f0.elapsed_time(f1) / K
        vs.set_option("scan_fp16", 1)
        f_bytes = N_ROWS * DIM * 4
        f_traffic, f_src = profiled_traffic("r2_c2_scan_full.txt", "r1/r1_c2_scan_full.txt")
        fp32_leg = {"kernel": "scan_tma_kernel<TPR=4,U=2,L2,WarpTopKReg> (K1, option scan_fp16 = 0)", "algorithmic_bytes_per_launch": f_bytes,
                    "kernel_ms_in_timed_region": f_step, "distance_evals_per_s": N_ROWS / (f_step * 1e-3),
                    "achieved": f_bytes / (f_step * 1e-3) / 1e9,
                    "frac": f_bytes / (f_step * 1e-3) / 1e9 / hbm_peak,
                    "launched_alone": {"kernel_ms": f_ms, "kernel_ms_median": f_med, "achieved": f_bytes / (f_ms * 1e-3) / 1e9,
                                       "frac": f_bytes / (f_ms * 1e-3) / 1e9 / hbm_peak},
                    "traffic": f_traffic, "traffic_source": f_src}

    # ---- end to end through the public host API: pinned host query in, ids + scores out, every step -----
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
    last_ids, last_sc = np.asarray(r[0]), np.asarray(r[1])
    # The same K host calls issued by TWO request threads: the reference's caller scores segments on many executor
    # threads at once (J/fdb/FdbVectorIndex.java:418-432) and libvsgpu gives every calling thread its own stream and
    # staging, so consecutive queries overlap on the device as they do in the device-resident region.  At N > 1 every
    # request thread has its own peer communicator (thread t of every rank exchanges with thread t of the others).
    e2e_multi = {}
    extra_sh = []

    def run_request_threads(nt):
        """The same K host calls issued by nt request threads (thread t of every rank exchanges with thread t of the others)."""
        callers = [(lambda q: seg.bruteforce_topk(q, TOPK))] * nt
        if world > 1:
            while len(extra_sh) < nt - 1:
                extra_sh.append(ShardedSegment(seg, rank, world, slots=1))
            callers = [(lambda q: sh.bruteforce_topk(q, TOPK))] + [(lambda q, s_=s_: s_.bruteforce_topk(q, TOPK)) for s_ in extra_sh[:nt - 1]]
        start, errs = threading.Barrier(nt + 1), []

        def worker(t):
            try:
                for i in range(W):
                    callers[t](qn[i])
                start.wait()
                for i in range(t, K, nt):
                    callers[t](qn[W + i])
            except Exception as e:  # noqa: BLE001
                errs.append(e)
                try:
                    start.abort()
                except Exception:  # noqa: BLE001
                    pass

        th = [threading.Thread(target=worker, args=(t,), daemon=True) for t in range(nt)]
        for t_ in th:
            t_.start()
        try:
            barrier()
            start.wait(timeout=60)
            with clocks:
                t0 = time.perf_counter()
                for t_ in th:
                    t_.join(timeout=60)
                dt = time.perf_counter() - t0
        except threading.BrokenBarrierError:
            dt = None
        ok = dt is not None and not errs and not any(t_.is_alive() for t_ in th)
        if world > 1:
            flag = torch.tensor([1 if ok else 0], dtype=torch.int32, device=dev)
            dist.all_reduce(flag, op=dist.ReduceOp.MIN)
            ok = bool(flag.item())
        if not ok:
            return None
        dt = max_over_ranks(dt)
        return {"value": world * N_ROWS * K / dt, "unit": "distance-evals/s", "ms_per_step": dt / K * 1e3,
                "qps": K / dt, "request_threads": nt}

    if not args.one_request_thread:
        vs.set_option("scan_reserve_sms", reserve_sms)
        for nt in (2,):  # (three Python request threads measured slower than two: 71 vs 55 us per query at N = 1)
            r_ = run_request_threads(nt)
            if r_ is None:
                break
            e2e_multi[nt] = r_
        for s_ in extra_sh:
            s_.close()

    # ---- parity: the last query against the oracle over ALL ranks' rows (rank 0), both paths -------------
    parity = False
    if rank == 0 and not args.no_check:
        oracle_check_c2(world, qn[W + K - 1], last_ids, last_sc)
        oracle_check_c2(world, qn[W + K - 1], dev_last[0], dev_last[1])
        parity = True

    # ---- CPU baseline: bounded sample on rank 0 at N=1 only ---------------------------------------------
    cpu = None
    if world == 1 and not args.no_cpu:
        cval, cms, cdone, cthreads, csample = cpu_bruteforce(1, 200, 2, 10.0)
        sval, sms_, sdone, _, ssample = cpu_bruteforce(1, 20, 1, 6.0, threads=1)
        cpu = {"value": cval, "unit": "distance-evals/s", "cores": cthreads, "kind": "port", "sample": csample,
               "ms_per_query": cms,
               "single_thread": {"value": sval, "unit": "distance-evals/s", "cores": 1, "ms_per_query": sms_, "sample": ssample}}

    extra = {}
    if args.workload == "all":
        if world == 1:
            extra["c2_b1024"] = bench_c2_batch(vs, L, lib, torch, dev, seg, min(K, 50), W)
        seg.free()
        seg = None
        ctx = dict(vs=vs, L=L, lib=lib, torch=torch, dist=dist, dev=dev, rank=rank, world=world, barrier=barrier,
                   max_over_ranks=max_over_ranks, hbm_peak=hbm_peak, check=not args.no_check,
                   n_slots=int(os.environ.get("VS_BENCH_C4_SLOTS", "2")))
        extra["c3"] = bench_c3(ctx)
        extra["c4"] = bench_c4(ctx, min(K, 50), W)
        extra["c5"] = bench_c5(ctx, min(K, 10), W)
        if world == 1 and not args.no_cpu:
            extra["c1"] = bench_c1(ctx)

    if rank == 0:
        # the headline e2e is whole-job throughput through the host API: the better of one and two request threads
        # (both reported; host buffers in and out on every step in either case)
        one = {"value": e2e_value, "unit": "distance-evals/s", "ms_per_step": e2e_s / K * 1e3, "qps": K / e2e_s, "request_threads": 1}
        best = max([one] + list(e2e_multi.values()), key=lambda x: x["value"])
        e2e_obj = {"value": best["value"], "unit": "distance-evals/s", "h2d_bytes_per_step": DIM * 4,
                   "d2h_bytes_per_step": TOPK * 16 + 4, "ms_per_step": best["ms_per_step"], "qps": best["qps"],
                   "request_threads": best["request_threads"], "one_request_thread": one,
                   "two_request_threads": e2e_multi.get(2)}
        if half_path:
            traffic, traffic_src = profiled_traffic("r2_c2_scan_half_full.txt")
            kname = "scan_half_kernel<L2,KK=16,CPL=2> (K1h: streams the fp16 operand copy, exact scores of the candidates from the fp32 rows)"
            unit_note = (f"{dp * 2 + 4} B per distance evaluation = the row of the fp16 operand copy ({dp * 2} B) + its fp32 coefficient (4 B); "
                         f"SURVEY 8(d)'s figure for an fp32 scan is {DIM * 4} B")
        else:
            traffic, traffic_src = profiled_traffic("r2_c2_scan_full.txt", "r1/r1_c2_scan_full.txt")
            kname = "scan_tma_kernel<TPR=4,U=2,L2,WarpTopKReg> (K1)"
            unit_note = f"{DIM * 4} B per distance evaluation (SURVEY 8(d))"
        in_step = alg_bytes / (ms_per_step * 1e-3) / 1e9
        roofline = {"bound": "hbm", "achieved": in_step, "peak": hbm_peak, "unit": "GB/s",
                    "frac": in_step / hbm_peak, "traffic": traffic,
                    "traffic_source": f"{traffic_src} (ncu --set full, dram read+write per launch)" if traffic_src else None,
                    "peak_source": peak_src, "kernel": kname,
                    "algorithmic_bytes_per_launch": alg_bytes, "algorithmic_bytes_per_unit": unit_note,
                    "kernel_ms_in_timed_region": ms_per_step, "launches_per_call": launches_per_call,
                    "launched_alone": {"kernel_ms": kern_ms, "kernel_ms_median": kern_med, "achieved": achieved,
                                       "frac": achieved / hbm_peak},
                    "note": "achieved = the kernel's algorithmic bytes / its average duration over the timed region (CUDA events on the "
                            f"launching stream around the K steps / K launches: {len(slots)} independent queries are in flight on "
                            f"{len(slots)} streams, so a launch's set-up, tail and last-CTA merge overlap its neighbours' streaming); "
                            "launched_alone = the same bytes / the duration of one call with nothing else on the GPU (events around "
                            "every call: the scan and, on the fp16 path, the launch of the conditional fallback check). A read-only "
                            "stream can exceed the read+write copy rate used as peak"}
        if half_path:
            roofline["fp32_scan_equivalent"] = {
                "bytes_per_launch": N_ROWS * DIM * 4, "GBps_alone": N_ROWS * DIM * 4 / (kern_ms * 1e-3) / 1e9,
                "GBps_in_step": N_ROWS * DIM * 4 / (ms_per_step * 1e-3) / 1e9,
                "note": "rate an fp32 scan would need for the same time: above the HBM peak, i.e. below the fp32 scan's floor of "
                        f"{N_ROWS * DIM * 4 / hbm_peak / 1e3:.1f} us per query -- possible only because half the bytes are read"}
        if fp32_leg is not None:
            roofline["fp32_scan"] = fp32_leg
        line = {
            "metric": METRIC, "value": value, "unit": "distance-evals/s",
            "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": DATA, "config": config_for(world),
            "arithmetic": ("returned ids and scores come from the reference's fp32 / fp64 arithmetic on the fp32 rows (bit-identical to the "
                           "oracle, parity_checked); the fp16 operand copy only nominates candidates within a proven error bound"
                           if half_path else "the reference's fp32 / fp64 arithmetic on the fp32 rows"),
            "parallelism": (f"row-range shards x{world}, cross-shard top-k exchange per query = {sh.exchange}; one C call per step "
                            f"(scan + push + merge); {len(slots)} queries in flight on {len(slots)} streams; one untimed lock-step "
                            "exchange before the timed region") if world > 1 else
                           f"1 GPU; independent queries alternate between {len(slots)} CUDA streams",
            "pipelining": f"the scan leaves {reserve_sms} SMs free so that the neighbouring query's prologue / merge "
                          "(and the cross-shard exchange at N > 1) run beside it",
            "qps": 1e3 / ms_per_step,
            "host_issue_us_per_step": issue_us,
            "parity_checked": parity,
            "roofline": roofline,
            "e2e": e2e_obj,
            "gpu_launches": int(launches),
            "clocks": clocks.summary(),
            "top10_last_query": last_ids.tolist(),
        }
        if cpu is not None:
            line["cpu_baseline"] = cpu
        if extra:
            line["extra"] = extra
        print(json.dumps(line), flush=True)
    clocks.close()
    if seg is not None:
        seg.free()
    sh.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def bench_c2_batch(vs, L, lib, torch, dev, seg, K, W, nq=1024):
    """C2, query batch 1024: one step = 1024 queries against the resident 1M x 128 segment (batch.cu:
    tcgen05 fp16 nomination GEMM + exact re-score).  Tensor-bound: 2 * N * D * B flop per batch."""
    import numpy as np

    tpeak, tsrc = measured_tensor_peak()
    qseg = vs.Segment.generate(QUERY_SEED + 1, 0, nq, DIM)
    qn = qseg.rows()
    qseg.free()
    q_host = torch.from_numpy(qn).pin_memory()
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
    traffic, traffic_src = profiled_traffic("r2_c2_b1024_gemm_full.txt")
    return {"workload": f"C2: exact L2 brute-force top-{TOPK} over {N_ROWS}x{DIM} fp32, query batch {nq}, 1 GPU",
            "ms_per_batch": ms, "qps": nq / (ms * 1e-3), "distance_evals_per_s": N_ROWS * nq / (ms * 1e-3),
            "roofline": {"bound": "tensor", "achieved": tf, "peak": tpeak, "unit": "TFLOP/s", "frac": tf / tpeak,
                         "peak_source": tsrc, "algorithmic_flops_per_batch": flops,
                         "kernel": "batch_gemm_kernel<STAT,HALF,L2,G64> (K2, tcgen05 kind::f16) + batch_select + fallback; "
                                   "achieved is over the WHOLE batch",
                         "traffic": traffic, "traffic_source": traffic_src},
            "e2e": {"ms_per_batch": e2e_ms, "qps": nq / (e2e_ms * 1e-3), "distance_evals_per_s": N_ROWS * nq / (e2e_ms * 1e-3),
                    "h2d_bytes_per_step": nq * DIM * 4, "d2h_bytes_per_step": nq * (TOPK * 16 + 4)},
            "gpu_launches": int(launches), "matches_per_query_scan": bool(same)}


def _shard(n_total: int, rank: int, world: int):
    from vectorsearch_b200.sharded import shard_range

    return shard_range(n_total, rank, world)


def bench_c3(ctx):
    """C3: PqTrainer.train (5 iterations, seed 42) + PqEncoder.encode over 10M x 128 IN TOTAL, M=16, K=256, rows sharded
    by range over the GPUs.  At N > 1 the per-iteration all-reduce of cluster sums and counts is libvsgpu's own exchange
    over the peer buffers (K10); both modes are timed: exact_order (centroids bit-identical to the reference) and one
    rank-ordered all-reduce per iteration (north_star's scheme)."""
    import numpy as np

    from vectorsearch_b200.sharded import ShardedSegment

    vs, torch, rank, world = ctx["vs"], ctx["torch"], ctx["rank"], ctx["world"]
    n_total = int(os.environ.get("VS_C3_ROWS", C3_ROWS))
    lo, hi = _shard(n_total, rank, world)
    seg = vs.Segment.generate(CORPUS_SEED, lo, hi - lo, DIM, id_base=lo)
    sh = ShardedSegment(seg, rank, world, slots=1) if world > 1 else None
    out = {"workload": f"C3: PQ k-means train (5 iterations, seed 42) + encode, M={PQ_M} K={PQ_K}, {n_total}x{DIM} in total over {world} GPU(s)"}

    def train(exact: bool):
        if world == 1:
            return vs.PqTrainer.train(None, DIM, PQ_M, PQ_K, 5, 42, segment=seg)
        return sh.pq_train(n_total, lo, PQ_M, PQ_K, 5, 42, exact_order=exact)

    cent = None
    for name, exact in (("exact_order", True), ("one_allreduce", False)):
        if world == 1 and not exact:
            continue
        train(exact)  # warm (scratch pools, plans)
        ctx["barrier"]()
        best = None
        for _ in range(2):
            t0 = time.perf_counter()
            c = train(exact)
            dt = ctx["max_over_ranks"](time.perf_counter() - t0)
            ctx["barrier"]()
            best = dt if best is None else min(best, dt)
        if exact:
            cent = c
        out[f"train_s_{name}" if world > 1 else "train_s"] = best
    # encode: every rank its own rows, no collective
    seg.attach_pq(cent)  # warm
    ctx["barrier"]()
    enc = None
    for _ in range(3):
        t0 = time.perf_counter()
        seg.attach_pq(cent)
        torch.cuda.synchronize()
        dt = ctx["max_over_ranks"](time.perf_counter() - t0)
        if os.environ.get("VS_BENCH_DEBUG"):
            print(f"[c3] attach_pq {dt:.4f}s", file=sys.stderr, flush=True)
        enc = dt if enc is None else min(enc, dt)
    out["encode_s"] = enc
    out["encode_vectors_per_s"] = n_total / enc
    tr = out.get("train_s_one_allreduce", out.get("train_s"))
    out["train_subdistance_evals_per_s"] = 5.0 * n_total * PQ_M * PQ_K / tr
    passes_s = (tr + enc) / 6.0
    out["roofline"] = {"bound": "hbm", "achieved": n_total / world * DIM * 4 / passes_s / 1e9, "peak": ctx["hbm_peak"], "unit": "GB/s",
                       "frac": n_total / world * DIM * 4 / passes_s / 1e9 / ctx["hbm_peak"],
                       "kernel": "pq_tc_assign_kernel (K3) + partition + chain sums (K4), per pass over the rows",
                       "algorithmic_bytes_per_pass_per_gpu": n_total // world * DIM * 4, "traffic": None,
                       "note": "6 passes (5 Lloyd iterations + 1 encode); ALU-bound argmin epilogue, not HBM"}
    if ctx["check"]:
        # parity: a code sample of this rank against the oracle (every rank checks its own rows), centroid digest on rank 0
        from oracle import pyoracle

        orc = pyoracle.get()
        m = min(20000, hi - lo)
        got = seg.codes(0, m)
        want = orc.pq_encode_batch_fast(cent, orc.gen_rows(CORPUS_SEED, lo, m, DIM), threads=max(1, host_threads() // world))
        if not np.array_equal(got, want):
            raise SystemExit(f"PARITY FAILURE (C3 codes, rank {rank})")
        out["parity_checked"] = "codes of the first 20000 rows of every shard equal the oracle's for the trained codebook"
    seg.free()
    if sh is not None:
        sh.close()
    return out


def bench_c4(ctx, K, W):
    """C4: PQ ADC top-100 scan + exact re-rank to top-10 over 100M x 128 IN TOTAL (M=16 codes), rows sharded by range over
    the GPUs (strong scaling).  The codebook is trained on the first 1M rows (identical on every rank)."""
    import numpy as np

    from vectorsearch_b200.sharded import ShardedSegment

    vs, L, lib, torch, dev = ctx["vs"], ctx["L"], ctx["lib"], ctx["torch"], ctx["dev"]
    rank, world, hbm_peak = ctx["rank"], ctx["world"], ctx["hbm_peak"]
    n_total = int(os.environ.get("VS_C4_ROWS", C4_ROWS))
    lo, hi = _shard(n_total, rank, world)
    n = hi - lo
    t0 = time.perf_counter()
    seg = vs.Segment.generate(CORPUS_SEED, lo, n, DIM, id_base=lo)
    torch.cuda.synchronize()
    t_gen = time.perf_counter() - t0
    t0 = time.perf_counter()
    train = vs.Segment.generate(CORPUS_SEED, 0, min(n_total, 1_000_000), DIM)
    cent = vs.PqTrainer.train(None, DIM, PQ_M, PQ_K, 5, 42, segment=train)
    t_train_cold = time.perf_counter() - t0
    t0 = time.perf_counter()
    vs.PqTrainer.train(None, DIM, PQ_M, PQ_K, 5, 42, segment=train)
    t_train = time.perf_counter() - t0
    train.free()
    t0 = time.perf_counter()
    seg.attach_pq(cent)
    t_enc_cold = ctx["max_over_ranks"](time.perf_counter() - t0)
    t0 = time.perf_counter()
    seg.attach_pq(cent)
    t_enc = ctx["max_over_ranks"](time.perf_counter() - t0)
    sh = ShardedSegment(seg, rank, world, slots=ctx["n_slots"])
    qseg = vs.Segment.generate(QUERY_SEED, 0, max(W + K, 64), DIM)
    qh = qseg.rows()
    qseg.free()
    q_dev = torch.from_numpy(qh).to(dev)
    stream = torch.cuda.current_stream()

    def run(nq_per_step: int, steps: int):
        """steps x (nq_per_step queries) alternating between the coordinator's streams; ms per QUERY, max over ranks"""
        for s_ in sh._streams:
            s_.wait_stream(stream)
        for i in range(min(W, 3)):
            sh.adc_rerank_topk_pipelined(q_dev[:nq_per_step], nq_per_step, C4_NCAND, TOPK)
        sh.drain()
        ctx["barrier"]()
        sh.adc_rerank_topk_pipelined(q_dev[:nq_per_step], nq_per_step, C4_NCAND, TOPK)  # untimed lock-step exchange
        sh.drain()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for s_ in sh._streams:
            s_.wait_stream(stream)
        for i in range(steps):
            j = (W + i * nq_per_step) % max(1, qh.shape[0] - nq_per_step)
            sh.adc_rerank_topk_pipelined(q_dev[j:j + nq_per_step], nq_per_step, C4_NCAND, TOPK)
        sh.drain()
        e1.record(stream)
        ctx["barrier"]()
        return ctx["max_over_ranks"](e0.elapsed_time(e1)) / (steps * nq_per_step)

    # the ADC scan alone on this rank's codes (LUT build + fast scan + conditional fallback launch), for its roofline:
    # once before the sustained runs (after a short idle: the board's state for a kernel timed alone) and once after
    ids100 = torch.empty((1, C4_NCAND), dtype=torch.int64, device=dev)
    ap100 = torch.empty((1, C4_NCAND), dtype=torch.float64, device=dev)
    cn = torch.empty((1,), dtype=torch.int32, device=dev)
    st = stream.cuda_stream

    def adc_alone():
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        for i in range(3):
            L.check(lib.vs_adc_topk_dev(seg.handle, q_dev[i].data_ptr(), 1, C4_NCAND, ids100.data_ptr(), ap100.data_ptr(), cn.data_ptr(), st))
        e0.record()
        for i in range(K):
            L.check(lib.vs_adc_topk_dev(seg.handle, q_dev[W + i].data_ptr(), 1, C4_NCAND, ids100.data_ptr(),
                                        ap100.data_ptr(), cn.data_ptr(), st))
        e1.record()
        torch.cuda.synchronize()
        return ctx["max_over_ranks"](e0.elapsed_time(e1) / K)

    torch.cuda.synchronize()
    time.sleep(2.0)
    ms_adc = adc_alone()
    # (option adc_reserve_sms: the scan leaves a few SMs to the other stream's LUT build / re-rank / exchange.  Measured at
    # N = 1 with 4 SMs: 341 / 284 / 303 us per query at 1 / 8 / 32 per launch against 336 / 285 / 298 without -- off by default)
    adc_reserve = int(os.environ.get("VS_ADC_RESERVE", "0"))
    vs.set_option("adc_reserve_sms", adc_reserve)
    ms = run(1, K)
    ms_b8 = run(8, max(4, K // 4))
    ms_b32 = run(32, max(4, K // 8))
    vs.set_option("adc_reserve_sms", 0)
    ms_adc_loaded = adc_alone()
    gbs = n * PQ_M / (ms_adc * 1e-3) / 1e9
    # end to end through the host API: query from host memory in, ids + scores out, every query
    call = (lambda q: sh.adc_rerank_topk(q, C4_NCAND, TOPK)) if world > 1 else (lambda q: seg.adc_rerank_topk(q, C4_NCAND, TOPK))
    for i in range(3):
        call(qh[i])
    ctx["barrier"]()
    t0 = time.perf_counter()
    for i in range(K):
        res = call(qh[W + i])
    ctx["barrier"]()
    e2e_ms = ctx["max_over_ranks"](time.perf_counter() - t0) / K * 1e3
    traffic, traffic_src = profiled_traffic("r2_c4_adc_full.txt")
    out = {"workload": f"C4: ADC top-{C4_NCAND} + exact re-rank top-{TOPK} over {n_total}x{DIM} in total (M={PQ_M} uint8 codes), {world} GPU(s)",
           "rows_per_gpu": n, "ms_per_query": ms, "adc_evals_per_s": n_total / (ms * 1e-3), "qps": 1e3 / ms,
           "ms_per_query_batch8": ms_b8, "adc_evals_per_s_batch8": n_total / (ms_b8 * 1e-3),
           "ms_per_query_batch32": ms_b32, "adc_evals_per_s_batch32": n_total / (ms_b32 * 1e-3),
           "adc_reserve_sms": adc_reserve, "floor_ms_per_query": n * PQ_M / (hbm_peak * 1e9) * 1e3,
           "streams": len(sh._streams),
           "e2e": {"ms_per_query": e2e_ms, "adc_evals_per_s": n_total / (e2e_ms * 1e-3), "h2d_bytes_per_step": DIM * 4,
                   "d2h_bytes_per_step": TOPK * 16 + 4},
           "roofline": {"bound": "hbm", "achieved": gbs, "peak": hbm_peak, "unit": "GB/s", "frac": gbs / hbm_peak,
                        "kernel": "build_lut_mm + adc_fastscan_kernel<4> (K5+K6), one rank's codes, launched alone",
                        "ms_per_launch_group": ms_adc, "ms_per_launch_group_after_sustained_load": ms_adc_loaded,
                        "algorithmic_bytes_per_launch": n * PQ_M,
                        "note": "timed alone after 2 s of idle; the same loop right after the sustained runs is slower "
                                "(the board stays in a lower clock state for seconds after ~50 ms of full load)",
                        "achieved_in_step": n * PQ_M / (ms * 1e-3) / 1e9, "traffic": traffic, "traffic_source": traffic_src},
           "generate_s": t_gen, "train_1M_5iters_s": t_train, "train_1M_5iters_cold_s": t_train_cold,
           "encode_s": t_enc, "encode_cold_s": t_enc_cold, "encode_vectors_per_s": n_total / t_enc,
           "encode_vectors_per_s_cold": n_total / t_enc_cold}
    if ctx["check"]:
        out["parity_checked"] = check_c4(ctx, seg, cent, lo, n, qh[W + K - 1], res)
    seg.free()
    sh.close()
    return out


def bench_c5(ctx, K, W):
    """C5: cosine brute-force top-50, query batch 256, 768-d rows sharded by range.  BASELINE's shape is 50M x 768 over 8
    GPUs = 6.25M rows (19.2 GB) per GPU; every N runs THAT per-GPU shape (N x 6.25M rows in total: the 153.6 GB corpus does
    not fit fewer than 8 GPUs next to its fp16 operand copy), so the 8-GPU line is C5 itself."""
    import numpy as np

    from vectorsearch_b200.sharded import ShardedSegment

    vs, torch, dev, rank, world = ctx["vs"], ctx["torch"], ctx["dev"], ctx["rank"], ctx["world"]
    n, d, nq, k = int(os.environ.get("VS_C5_ROWS", 6_250_000)), 768, 256, 50
    seg = vs.Segment.generate(CORPUS_SEED, rank * n, n, d, id_base=rank * n)
    sh = ShardedSegment(seg, rank, world, slots=1)
    qseg = vs.Segment.generate(QUERY_SEED, 0, nq, d)
    qn = qseg.rows()
    qseg.free()
    q_dev = torch.from_numpy(qn).to(dev)
    stream = torch.cuda.current_stream()
    for _ in range(max(3, min(W, 3))):  # the first call also builds the nomination state (row norms, fp16 copy)
        ids, sc, cn = sh.bruteforce_topk_dev(q_dev, nq, k, 1)
    torch.cuda.synchronize()
    ctx["barrier"]()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(K):
        ids, sc, cn = sh.bruteforce_topk_dev(q_dev, nq, k, 1)
    e1.record(stream)
    ctx["barrier"]()
    ms = ctx["max_over_ranks"](e0.elapsed_time(e1)) / K
    for _ in range(2):
        res = sh.bruteforce_topk(qn, k, 1) if world > 1 else seg.bruteforce_topk(qn, k, 1)
    ctx["barrier"]()
    t0 = time.perf_counter()
    for _ in range(max(2, K // 2)):
        res = sh.bruteforce_topk(qn, k, 1) if world > 1 else seg.bruteforce_topk(qn, k, 1)
    ctx["barrier"]()
    e2e_ms = ctx["max_over_ranks"](time.perf_counter() - t0) / max(2, K // 2) * 1e3
    tpeak, tsrc = measured_tensor_peak()
    traffic, traffic_src = profiled_traffic("r2_c5_gemm_pair_full.txt")
    flops = 2.0 * n * d * nq  # per GPU
    tf = flops / (ms * 1e-3) / 1e12
    out = {"workload": f"C5: cosine brute-force top-{k}, query batch {nq}, {world} x {n}x{d} fp32 rows ({world} GPU(s), rows sharded by range)",
           "rows_per_gpu": n, "rows_total": n * world, "ms_per_batch": ms, "qps": nq / (ms * 1e-3),
           "distance_evals_per_s": float(n) * world * nq / (ms * 1e-3),
           "e2e": {"ms_per_batch": e2e_ms, "qps": nq / (e2e_ms * 1e-3), "h2d_bytes_per_step": nq * d * 4, "d2h_bytes_per_step": nq * (k * 16 + 4)},
           "roofline": {"bound": "tensor", "achieved": tf, "peak": tpeak, "unit": "TFLOP/s", "frac": tf / tpeak, "peak_source": tsrc,
                        "algorithmic_flops_per_batch_per_gpu": flops,
                        "peak_sustained": sustained_tensor_peak(),
                        "frac_of_sustained": (tf / sustained_tensor_peak()) if sustained_tensor_peak() else None,
                        "hbm_view": {"algorithmic_bytes_per_batch_per_gpu": n * d * 4, "achieved_GBs": n * d * 4 / (ms * 1e-3) / 1e9,
                                     "frac_of_measured_peak": n * d * 4 / (ms * 1e-3) / 1e9 / ctx["hbm_peak"]},
                        "kernel": "batch_gemm_pair_kernel<streaming operands, HALF, COSINE> (K2 on CTA pairs, tcgen05 cta_group::2) + batch_select "
                                  "+ exchange; per GPU; achieved is over the WHOLE batch",
                        "traffic": traffic, "traffic_source": traffic_src,
                        "note": "profiles/r2_c5_gemm_pair_full.txt: the nomination kernel alone keeps the tensor pipe 97 % active "
                                "(sm__pipe_tensor_cycles_active) at the clock the board sustains under this load"}}
    if ctx["check"]:
        # parity: query 0 of the batch against the oracle over ALL shards' rows -- every rank scans its own rows (regenerated
        # slab by slab), rank 0 merges the per-rank lists like the reference merges segments
        from oracle import pyoracle

        orc = pyoracle.get()
        thr = max(1, host_threads() // world)
        oi, os_ = [], []
        for r0 in range(0, n, 500_000):
            cnt = min(500_000, n - r0)
            rows = orc.gen_rows(CORPUS_SEED, rank * n + r0, cnt, d)
            i_, s_, _ = orc.bruteforce_topk(rows, qn[0], k, 1, threads=thr)
            oi.append(i_ + rank * n + r0)
            os_.append(s_)
        li, ls = orc.merge_topk(np.concatenate(oi), np.concatenate(os_), k)
        pack = np.zeros((2, k), np.int64)
        pack[0], pack[1] = li, ls.view(np.int64)
        if world > 1:
            t = torch.from_numpy(pack).to(dev)
            g = torch.empty((world, 2, k), dtype=torch.int64, device=dev)
            ctx["dist"].all_gather_into_tensor(g.view(-1), t.view(-1))
            g = g.cpu().numpy()
        else:
            g = pack[None]
        if rank == 0:
            wi, ws = orc.merge_topk(g[:, 0].ravel(), g[:, 1].ravel().view(np.float64), k)
            gi, gs = np.asarray(res[0])[0], np.asarray(res[1])[0]
            if not (np.array_equal(gi, wi) and np.array_equal(gs.view(np.uint64), ws.view(np.uint64))):
                raise SystemExit(f"PARITY FAILURE (C5, {world} GPUs): got {gi.tolist()}, oracle {wi.tolist()}")
            out["parity_checked"] = "cosine top-50 of query 0 equals the oracle's over all shards' rows"
    seg.free()
    sh.close()
    return out


def bench_c1(ctx):
    """C1: the DistanceAndPqBenchmark shapes (JMH, single thread, Random(42) inputs).  The CPU column is the oracle port in a
    C loop (ns per call, the unit JMH reports); the GPU column is one C-ABI call per operation -- a kernel launch each,
    which is why the throughput path is the segment API and not these pair operations."""
    import numpy as np

    from oracle import pyoracle

    vs = ctx["vs"]
    orc, kind = _native_oracle()
    out = {"workload": "C1: DistanceAndPqBenchmark shapes: l2 / cosine dim 128 and 768, pqEncode M=16 K=256 dim=128, pqLutDistance M=16",
           "cpu": {"kind": f"oracle port, C loop, one thread ({kind})", "unit": "ns/op"}, "gpu": {"unit": "us/call through the C ABI (one launch each)"}}
    for dim in (128, 768):
        a, b = orc.gen_floats(42, 0, dim), orc.gen_floats(42, dim, dim)  # DistanceState: a then b from Random(42)
        out["cpu"][f"l2_{dim}"] = orc.bench_ns_per_op(0, a, b, iters=300_000)
        out["cpu"][f"cosine_{dim}"] = orc.bench_ns_per_op(1, a, b, iters=300_000)
        for name, fn in (("l2", vs.Distances.l2), ("cosine", vs.Distances.cosine)):
            fn(a, b)
            t0 = time.perf_counter()
            for _ in range(200):
                fn(a, b)
            out["gpu"][f"{name}_{dim}"] = (time.perf_counter() - t0) / 200 * 1e6
    cent = orc.gen_floats(42, 0, PQ_M * PQ_K * 8, 1).reshape(PQ_M, PQ_K, 8)   # PqState: centroids nextFloat()
    v = orc.gen_floats(42, PQ_M * PQ_K * 8, DIM)
    lut = orc.gen_floats(42, PQ_M * PQ_K * 8 + DIM, PQ_M * PQ_K, 2).reshape(PQ_M, PQ_K)
    codes = orc.gen_codes(42, PQ_M * PQ_K * 8 + DIM + PQ_M * PQ_K, PQ_M)
    out["cpu"]["pqEncode"] = orc.bench_ns_per_op(2, v, centroids=cent, iters=20_000)
    out["cpu"]["pqLutDistance"] = orc.bench_ns_per_op(3, centroids=cent, lut=lut, codes=codes, iters=2_000_000)
    vs.PqEncoder.encode(cent, v)
    t0 = time.perf_counter()
    for _ in range(100):
        vs.PqEncoder.encode(cent, v)
    out["gpu"]["pqEncode"] = (time.perf_counter() - t0) / 100 * 1e6
    assert np.array_equal(vs.PqEncoder.encode(cent, v), orc.pq_encode(cent, v))
    return out


def check_c4(ctx, seg, cent, lo, n, q, res):
    """The last query of the e2e loop against the oracle: every rank scans ITS codes with the oracle (first 100 by
    approximate distance, the reference's stable order), rank 0 merges the lists by (approx, global row) -- shards are
    ascending row ranges -- and re-ranks the global first 100 exactly from regenerated rows."""
    import numpy as np

    from oracle import pyoracle

    torch, dist, rank, world = ctx["torch"], ctx["dist"], ctx["rank"], ctx["world"]
    orc = pyoracle.get()
    codes = seg.codes()
    lut = orc.build_lut(cent, q)
    ci, ca = orc.adc_topn(lut, codes, C4_NCAND, threads=max(1, host_threads() // world))
    pack = np.full((2, C4_NCAND), -1, np.int64)
    pack[0, :len(ci)] = ci + lo
    pack[1, :len(ci)] = ca.view(np.int64)
    if world > 1:
        t = torch.from_numpy(pack).to(ctx["dev"])
        g = torch.empty((world,) + pack.shape, dtype=torch.int64, device=ctx["dev"])
        dist.all_gather_into_tensor(g.view(-1), t.view(-1))
        g = g.cpu().numpy()
    else:
        g = pack[None]
    if rank != 0:
        return None
    ids = np.concatenate([g[r, 0][g[r, 0] >= 0] for r in range(world)])
    ap = np.concatenate([g[r, 1][g[r, 0] >= 0] for r in range(world)]).view(np.float64)
    order = np.lexsort((ids, ap))[:C4_NCAND]   # ascending approximate distance, ties to the lower global row
    first = ids[order]
    cand_rows = np.concatenate([orc.gen_rows(CORPUS_SEED, int(i), 1, DIM) for i in first])
    pos, sc, _ = orc.rerank_topk(cand_rows, q, np.arange(len(first)), TOPK)
    want_i, want_s = first[pos], sc
    got_i, got_s = np.asarray(res[0]), np.asarray(res[1])
    if not (np.array_equal(got_i, want_i) and np.array_equal(got_s.view(np.uint64), want_s.view(np.uint64))):
        raise SystemExit(f"PARITY FAILURE (C4, {world} GPUs): got {got_i.tolist()}, oracle {want_i.tolist()}")
    return "ADC top-100 + re-rank top-10 of the last query equal the oracle's over all shards' codes"


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="all", choices=["c2", "all"],
                    help="c2 = the headline line only; all = also C2 batch 1024 (N = 1), C3 and C4 into `extra`")
    ap.add_argument("--one-request-thread", action="store_true", help="skip the two-request-thread end-to-end leg")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-check", action="store_true", help="skip the oracle parity checks after the timed regions")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
