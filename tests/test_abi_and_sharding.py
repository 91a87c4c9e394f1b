"""CPU-side checks: the C ABI (include/vsgpu.h) is fully exported and bound, the library refuses to
compute without a GPU, and the multi-GPU host logic (row-range shards + rank-ordered merge) is
right under a real 2-process gloo group."""
import os
import re
import socket
import subprocess
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent


def _declared_symbols():
    text = (ROOT / "include" / "vsgpu.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(vs_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from vectorsearch_b200 import _lib

    lib = _lib.load()  # raises if the in-tree .so is missing: there is no fallback
    names = _declared_symbols()
    assert len(names) >= 35
    for n in names:
        assert hasattr(lib, n), f"libvsgpu.so does not export {n}"
    assert sorted(_lib.SIGNATURES) == names, "ctypes binding and include/vsgpu.h disagree"


def test_no_cpu_fallback():
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    import vectorsearch_b200 as vs
    from vectorsearch_b200 import _lib

    with pytest.raises(_lib.VsError) as e:
        vs.init(0)
    assert e.value.code == _lib.VS_ECUDA and "no CPU fallback" in str(e.value)
    lib = _lib.load()
    import ctypes as C

    out = C.c_double()
    a = np.ones(4, np.float32)
    rc = lib.vs_l2(a.ctypes.data_as(_lib.f32p), a.ctypes.data_as(_lib.f32p), 4, C.byref(out))
    assert rc == _lib.VS_ECUDA


def test_product_does_not_import_the_oracle():
    for p in (ROOT / "vectorsearch_b200").rglob("*"):
        if p.suffix in (".py", ".cu", ".cuh", ".h"):
            t = p.read_text()
            assert not re.search(r"^\s*(from|import)\s+oracle\b", t, flags=re.M), p
            assert "vs_oracle" not in t and "libvsoracle" not in t, p


def test_shard_range_partitions_rows():
    from vectorsearch_b200.sharded import shard_range

    for n in (0, 1, 7, 8, 1000003):
        for w in (1, 2, 3, 8):
            spans = [shard_range(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(w - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


def test_merge_gathered_host_matches_oracle_merge(oracle):
    from vectorsearch_b200.sharded import merge_gathered_host

    rng = np.random.default_rng(5)
    world, k = 4, 6
    ids = rng.integers(0, 1000, size=(world, k)).astype(np.int64)
    sc = np.round(rng.normal(size=(world, k)), 1)  # many ties
    sc[1, 2] = np.nan
    sc[2, 0] = -0.0
    sc[3, 0] = 0.0
    for r in range(world):
        order = np.argsort(-np.nan_to_num(sc[r], nan=np.inf), kind="stable")
        ids[r], sc[r] = ids[r][order], sc[r][order]
    counts = np.array([k, k, 3, 0])
    cat_i = np.concatenate([ids[r, :counts[r]] for r in range(world)])
    cat_s = np.concatenate([sc[r, :counts[r]] for r in range(world)])
    oi, os_ = oracle.merge_topk(cat_i, cat_s, k)
    gi, gs = merge_gathered_host(ids, sc, counts, k, desc=True)
    assert np.array_equal(gi, oi) and np.array_equal(gs.view(np.uint64), os_.view(np.uint64))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n, d, k, metric, q, qv):
    import torch.distributed as dist

    from oracle import pyoracle
    from vectorsearch_b200.sharded import gather_and_merge_host, shard_range

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        orc = pyoracle.get()
        lo, hi = shard_range(n, rank, world)
        rows = orc.gen_rows(9, lo, hi - lo, d)  # this rank's row range of the corpus
        rows[(hi - lo) // 2] = orc.gen_rows(9, 0, 1, d)[0]  # a row shared by every shard: cross-rank ties
        ids, sc, _ = orc.bruteforce_topk(rows, qv, k, metric)  # stands in for the device scan of the shard
        pad_i = np.full(k, -1, np.int64)
        pad_s = np.full(k, np.nan)
        pad_i[:len(ids)] = ids + lo
        pad_s[:len(ids)] = sc
        gi, gs = gather_and_merge_host(pad_i, pad_s, len(ids), k, desc=True)
        q.put((rank, gi.tolist(), gs.view(np.uint64).tolist()))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("metric", [0, 1])
def test_two_rank_gloo_merge_equals_single_segment(oracle, metric):
    import torch.multiprocessing as mp

    world, n, d, k = 2, 4001, 16, 9
    ctx = mp.get_context("spawn")
    queue = ctx.Queue()
    port = _free_port()
    qv = oracle.gen_floats(10, 0, d)
    procs = [ctx.Process(target=_worker, args=(r, world, port, n, d, k, metric, queue, qv)) for r in range(world)]
    for p in procs:
        p.start()
    got = sorted(queue.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    # the whole corpus as ONE segment: what the reference would return (ties to the lowest row)
    from vectorsearch_b200.sharded import shard_range

    full = oracle.gen_rows(9, 0, n, d)
    for r in range(world):
        lo, hi = shard_range(n, r, world)
        full[lo + (hi - lo) // 2] = full[0]
    oi, os_, _ = oracle.bruteforce_topk(full, _q(oracle, d), k, metric)
    for rank, gi, gs in got:
        assert gi == oi.tolist() and gs == os_.view(np.uint64).tolist(), f"rank {rank}"


def _q(oracle, d):
    return oracle.gen_floats(10, 0, d)


def _adc_worker(rank, world, port, n, d, M, K, n_cand, k, q, qv):
    """One rank of the cross-shard ADC + re-rank: the oracle stands in for the device kernels of the shard,
    the exchange is the same single all-gather of the packed candidates, the merge is the host restatement
    of vs_merge_adc_rerank_packed_dev."""
    import torch
    import torch.distributed as dist

    from oracle import pyoracle
    from vectorsearch_b200.sharded import merge_adc_rerank_host, shard_range

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        orc = pyoracle.get()
        full = orc.gen_rows(9, 0, n, d)
        full[n // 2:n // 2 + 20] = full[3]            # duplicates that straddle nothing but tie in both distances
        full[5:25] = full[3]
        cent = orc.pq_train(full[:2000], d, M, K, 2, 42)
        lo, hi = shard_range(n, rank, world)
        rows = full[lo:hi]
        codes = orc.pq_encode_batch(cent, rows)
        ci, ca = orc.adc_topn(orc.build_lut(cent, qv), codes, n_cand)
        pack = np.zeros((4, n_cand), np.int64)
        pack[3] = -1
        c = len(ci)
        pack[0, :c] = ci + lo
        pack[1, :c] = np.asarray(ca, np.float64).view(np.int64)
        for j in range(c):                             # exact score of every candidate, in candidate order
            _, sc, _ = orc.rerank_topk(rows, qv, ci[j:j + 1], 1, 0)
            pack[2, j] = np.asarray(sc[:1], np.float64).view(np.int64)[0]
            pack[3, j] = 1
        tp = torch.from_numpy(pack.reshape(-1))
        gath = [torch.zeros_like(tp) for _ in range(world)]
        dist.all_gather(gath, tp)
        g = torch.stack(gath).numpy().reshape(world, 4, n_cand)
        gi, gs = merge_adc_rerank_host(g, k)
        q.put((rank, gi.tolist(), gs.view(np.uint64).tolist()))
    finally:
        dist.destroy_process_group()


def test_two_rank_gloo_adc_rerank_equals_single_segment(oracle):
    import torch.multiprocessing as mp

    world, n, d, M, K, n_cand, k = 2, 3001, 16, 4, 16, 40, 7
    ctx = mp.get_context("spawn")
    queue = ctx.Queue()
    port = _free_port()
    qv = oracle.gen_floats(10, 0, d)
    procs = [ctx.Process(target=_adc_worker, args=(r, world, port, n, d, M, K, n_cand, k, queue, qv)) for r in range(world)]
    for p in procs:
        p.start()
    got = sorted(queue.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    full = oracle.gen_rows(9, 0, n, d)
    full[n // 2:n // 2 + 20] = full[3]
    full[5:25] = full[3]
    cent = oracle.pq_train(full[:2000], d, M, K, 2, 42)
    codes = oracle.pq_encode_batch(cent, full)
    ci, _ = oracle.adc_topn(oracle.build_lut(cent, qv), codes, n_cand)
    ri, rs, _ = oracle.rerank_topk(full, qv, ci, k, 0)
    for rank, gi, gs in got:
        assert gi == ri.tolist() and gs == rs.view(np.uint64).tolist(), f"rank {rank}"


# ---- round 2: a plain-C caller, host-only wire-format functions, the order-preserving threaded oracle --------------
def _build_c_smoke(tmp_path):
    exe = tmp_path / "c_abi_smoke"
    libdir = ROOT / "vectorsearch_b200"
    cc = "/usr/bin/gcc" if Path("/usr/bin/gcc").exists() else "gcc"
    subprocess.run([cc, "-O1", "-Wall", "-o", str(exe), str(ROOT / "tests" / "c_abi_smoke.c"), f"-L{libdir}", "-lvsgpu",
                    f"-Wl,-rpath,{libdir}", "-lm"], check=True, capture_output=True)
    return exe


def test_plain_c_caller_links_and_fails_loudly_without_a_device(tmp_path):
    """tests/c_abi_smoke.c includes only include/vsgpu.h and links -lvsgpu: the boundary needs neither Python nor
    torch.  Without a GPU it must stop at vs_init with VS_ECUDA and the no-CPU-fallback message (exit 77)."""
    import torch

    exe = _build_c_smoke(tmp_path)
    r = subprocess.run([str(exe)], capture_output=True, text=True, timeout=120)
    if torch.cuda.is_available():
        assert r.returncode == 0, r.stdout + r.stderr
    else:
        assert r.returncode == 77, r.stdout + r.stderr
        assert "no CPU fallback" in r.stdout


@pytest.mark.gpu
def test_plain_c_caller_on_the_device(tmp_path):
    exe = _build_c_smoke(tmp_path)
    for ngpu in ("1", "2"):  # "2": vs_init_multi with two ranks on device 0 -- the in-library coordinator from plain C
        r = subprocess.run([str(exe), ngpu], capture_output=True, text=True, timeout=300)
        assert r.returncode == 0, r.stdout + r.stderr
        assert "c_abi_smoke ok" in r.stdout


def test_codebook_wire_format_on_the_host(oracle):
    """vs_codebook_encode / _decode are host-side byte handling (no device needed): the bytes equal what the real
    protobuf runtime serializes for the reference's PQCodebook message (vectorsearch.proto:135-142)."""
    from google.protobuf import descriptor_pb2, descriptor_pool, message_factory

    import vectorsearch_b200 as vs

    fdp = descriptor_pb2.FileDescriptorProto()
    fdp.name, fdp.package, fdp.syntax = "vs_cpu_wire_test.proto", "vscpu", "proto3"
    T = descriptor_pb2.FieldDescriptorProto
    cb = fdp.message_type.add()
    cb.name = "PQCodebook"
    for name, num, typ, lab in (("m", 1, T.TYPE_INT32, T.LABEL_OPTIONAL), ("k", 2, T.TYPE_INT32, T.LABEL_OPTIONAL),
                                ("centroids", 3, T.TYPE_BYTES, T.LABEL_REPEATED)):
        f = cb.field.add()
        f.name, f.number, f.type, f.label = name, num, typ, lab
    pool = descriptor_pool.DescriptorPool()
    pool.Add(fdp)
    PQCodebook = message_factory.GetMessageClass(pool.FindMessageTypeByName("vscpu.PQCodebook"))
    for M, K, sub in ((16, 256, 8), (2, 2, 2), (1, 300, 3), (4, 16, 48)):
        cent = oracle.gen_floats(5, 0, M * K * sub).reshape(M, K, sub)
        msg = PQCodebook(m=M, k=K)
        for s in range(M):  # SegmentBuildService.java:330-336: putFloat in (ci, di) order, little-endian
            msg.centroids.append(b"".join(oracle.floats_to_bytes(cent[s, ci]) for ci in range(K)))
        want = msg.SerializeToString()
        assert vs.codebook_encode(cent) == want
        back = vs.codebook_decode(want)
        assert back.shape == (M, K, sub) and np.array_equal(back.view(np.uint32), cent.view(np.uint32))
        again = PQCodebook()
        again.ParseFromString(vs.codebook_encode(cent))  # and the real runtime parses ours
        assert again.m == M and again.k == K and len(again.centroids) == M
    for bad in (b"\x08", b"\x08\x02\x10\x02\x1a\x05abcd", b"\x08\x02\x10\x02\x1a\x10" + b"x" * 16):
        with pytest.raises(ValueError):
            vs.codebook_decode(bad)


@pytest.mark.parametrize("lanes", [16, 8])
def test_threaded_oracle_trainer_is_bit_identical(oracle, lanes):
    """vso_pq_train_mt / vso_pq_encode_batch_fast (used by the BASELINE-size GPU tests) against the literal
    single-threaded restatements: same centroids, same draws, same codes, for every thread count."""
    oracle.set_lanes(lanes)
    try:
        for n, d, M, K, it in ((6000, 64, 8, 32, 3), (500, 16, 2, 64, 4), (3000, 96, 4, 16, 2)):
            rows = oracle.gen_rows(21, 0, n, d)
            rows[100:160] = rows[7]  # duplicates: empty clusters and re-initialisation draws
            want, draws = oracle.pq_train(rows, d, M, K, it, 42, return_draws=True)
            for t in (1, 3, 8):
                got, gd = oracle.pq_train(rows, d, M, K, it, 42, return_draws=True, threads=t)
                assert gd == draws and np.array_equal(got.view(np.uint32), want.view(np.uint32)), (n, d, t)
            assert np.array_equal(oracle.pq_encode_batch_fast(want, rows, threads=5), oracle.pq_encode_batch(want, rows, threads=2))
    finally:
        oracle.set_lanes(16)
