"""GPU parity of the second-round boundary: one process driving several GPUs (vs_init_multi; on a one-GPU box
the ranks share the device), the reference's wire formats, the residency table, the expansion-scoring gather,
empty shards in an exchange and concurrent vs_segment_free.  Every numeric result is compared bit-exactly with
the oracle or with the single-GPU path (itself oracle-checked in test_gpu_parity.py)."""
import ctypes as C
import threading

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def vs():
    import vectorsearch_b200 as v

    v.init(0)
    yield v
    v.init(0)
    v.set_simd_lanes(16)


def _same(a, b):
    a, b = np.asarray(a), np.asarray(b)
    if a.shape != b.shape:
        return False
    if a.dtype != np.float64:
        return np.array_equal(a, b)
    na, nb = np.isnan(a), np.isnan(b)
    return np.array_equal(na, nb) and np.array_equal(a[~na].view(np.uint64), b[~nb].view(np.uint64))


def _devices(world):
    import torch

    n = torch.cuda.device_count()
    return [i % n for i in range(world)]  # distinct GPUs when the box has them, else ranks share device 0


# ---- one process, several GPUs ------------------------------------------------------------------------------------
@pytest.mark.parametrize("world", [2, 3])
def test_single_process_group_queries(vs, oracle, world):
    """vs_init_multi: vs_segment_upload shards the rows, every query entry point fans out inside libvsgpu and
    returns what one GPU returns (J/fdb/FdbVectorIndex.java:418-437, 676-721, 754-769, 820-828, 997-1043)."""
    n, d, M, K, k, n_cand = 50_021, 64, 8, 64, 10, 100
    rows = oracle.gen_rows(42, 0, n, d)
    rows[40_000:40_006] = rows[7]        # ties across shards: the lower global row wins
    rows[123] = np.nan                    # a NaN row sorts first under Double.compare
    skip = np.zeros(n, np.uint8)
    skip[[3, 20_000, 45_000]] = 1
    qs = oracle.gen_rows(43, 0, 6, d)
    qs[0] = rows[7]
    cent = oracle.pq_train(rows[200:4200], d, M, K, 2, 42)  # (not the NaN row)
    vs.init_multi(_devices(world))
    try:
        assert vs.device_count() == world
        seg = vs.Segment.upload(rows, skip=skip, id_base=1000)
        assert (seg.n, seg.d, seg.id_base) == (n, d, 1000)
        assert np.array_equal(seg.rows(16_000, 20_000).view(np.uint32), rows[16_000:36_000].view(np.uint32))
        for metric in (0, 1):
            for i in range(3):
                gi, gs = seg.bruteforce_topk(qs[i], k, metric)
                oi, os_, _ = oracle.bruteforce_topk(rows, qs[i], k, metric, skip=skip)
                assert np.array_equal(gi, oi + 1000) and _same(gs, os_), (metric, i, gi, oi)
        bi, bs, bc = seg.bruteforce_topk(qs, k)  # a batch: tensor-core nomination per shard
        for i in range(len(qs)):
            oi, os_, _ = oracle.bruteforce_topk(rows, qs[i], k, 0, skip=skip)
            assert bc[i] == len(oi) and np.array_equal(bi[i], oi + 1000) and _same(bs[i], os_)
        # sealing: codes encoded on every shard's device
        seg.attach_pq(cent)
        codes = oracle.pq_encode_batch(cent, rows, threads=4)
        assert np.array_equal(seg.codes(), codes)
        assert np.array_equal(seg.codes(16_600, 20_000), codes[16_600:36_600])
        assert np.array_equal(vs.PqEncoder.encode_batch(cent, segment=seg), codes)
        for i in range(3):
            lut = oracle.build_lut(cent, qs[i])
            ci, ca = oracle.adc_topn(lut, codes, n_cand)
            gi, ga = seg.adc_topk(qs[i], n_cand)
            assert np.array_equal(gi, ci + 1000) and _same(ga, ca)
            for metric in (0, 1):
                ri, rs, _ = oracle.rerank_topk(rows, qs[i], ci, k, metric, skip=skip)
                fi, fs = seg.adc_rerank_topk(qs[i], n_cand, k, metric)
                assert np.array_equal(fi, ri + 1000) and _same(fs, rs)
            # caller-supplied candidates: duplicates, ids of every shard interleaved, missing and skipped records
            cand = np.concatenate([ci[::-1], [-5, n + 7, 3, 20_000, ci[0], 40_003, 40_001]]).astype(np.int64)
            ri, rs, _ = oracle.rerank_topk(rows, qs[i], cand, k, 0, skip=skip)
            gi, gs = seg.rerank_topk(qs[i], np.where(cand >= 0, cand + 1000, cand), k)
            assert np.array_equal(gi, ri + 1000) and _same(gs, rs)
        # expansion scoring against the resident codes of all shards
        ids = np.array([0, 1, n - 1, n // world, n // world - 1, 33_333, -1, n, 5], np.int64)
        lut = oracle.build_lut(cent, qs[1])
        with seg.adc_query(qs[1]) as aq:
            for rep in range(2):
                dist, ok = aq.gather(np.where(ids >= 0, ids + 1000, ids))  # global ids = id_base + row; -1 and n + 1000 are nobody's
                for j, g in enumerate(ids):
                    if 0 <= g < n:
                        assert ok[j] and dist[j] == oracle.pq_approx_distance(lut, codes[g])
                    else:
                        assert not ok[j] and np.isnan(dist[j])
        seg.set_skip(None)
        gi, gs = seg.bruteforce_topk(qs[2], k)
        oi, os_, _ = oracle.bruteforce_topk(rows, qs[2], k)
        assert np.array_equal(gi, oi + 1000) and _same(gs, os_)
        seg.free()
    finally:
        vs.init(0)


@pytest.mark.parametrize("world,exact", [(2, True), (3, True), (2, False)])
def test_single_process_group_training(vs, oracle, world, exact):
    """vs_pq_train on a sharded handle = vs_pq_train_sharded_peer on every worker: sums and counts are combined by
    libvsgpu's own all-reduce over the peer buffers.  exact_order: centroids bit-identical to PqTrainer.train
    (J/pq/PqTrainer.java:28-91); otherwise the rank-ordered sum of the shards' partial sums, emulated here."""
    n, d, M, K, iters = 9001, 32, 4, 32, 4
    rows = oracle.gen_rows(11, 0, n, d)
    rows[3000:3300] = rows[5]            # duplicate rows: empty clusters, re-initialisation draws from other shards
    vs.init_multi(_devices(world))
    try:
        vs.set_option("train_exact_order", 1 if exact else 0)
        seg = vs.Segment.upload(rows)
        got = vs.PqTrainer.train(None, d, M, K, iters, 42, segment=seg)
        if exact:
            want = oracle.pq_train(rows, d, M, K, iters, 42)
            assert np.array_equal(got.view(np.uint32), want.view(np.uint32))
        else:
            def distortion(c):
                codes = oracle.pq_encode_batch(c, rows, threads=4)
                recon = np.stack([c[s, codes[:, s]] for s in range(M)], axis=1).reshape(n, d)
                return float(((rows - recon) ** 2).sum(axis=1).mean())

            want = oracle.pq_train(rows, d, M, K, iters, 42)
            assert np.isfinite(got).all() and abs(distortion(got) - distortion(want)) <= 0.02 * distortion(want)
        seg.free()
    finally:
        vs.set_option("train_exact_order", 1)
        vs.init(0)


def test_group_rebind_and_errors(vs, oracle):
    from vectorsearch_b200 import _lib as L

    rows = oracle.gen_rows(1, 0, 1000, 16)
    vs.init_multi(_devices(2))
    try:
        seg = vs.Segment.upload(rows)
        with pytest.raises(L.VsError):          # no PQ attached
            seg.adc_topk(rows[0], 10)
        with pytest.raises(ValueError):          # PqTrainer.java:32-34
            vs.PqTrainer.train(None, 16, 3, 4, 1, 42, segment=seg)
        tiny = vs.Segment.upload(rows[:1])       # fewer rows than GPUs: one shard is empty
        gi, gs = tiny.bruteforce_topk(rows[0], 5)
        assert gi.tolist() == [0] and gs[0] == 0.0
        with pytest.raises(ValueError):
            vs.PqTrainer.train(None, 16, 4, 4, 1, 42, segment=tiny)
        h = seg.handle
    finally:
        vs.init(0)                               # re-binding frees the group's sharded segments
    assert vs.device_count() == 1
    n = C.c_int64()
    assert L.load().vs_segment_info(h, C.byref(n), None, None, None, None) == L.VS_EHANDLE


# ---- an empty shard takes part in the exchange ----------------------------------------------------------------------
def test_exchange_with_an_empty_rank(vs, oracle):
    """a rank whose shard has no rows must publish an all-empty list through the same exchange as its
    peers (threads as ranks, communicators connected by address)."""
    from vectorsearch_b200 import _lib as L

    lib = vs.load()
    world, n, d, k = 3, 5000, 32, 10
    rows = oracle.gen_rows(42, 0, n, d)
    q = oracle.gen_floats(43, 0, d)
    bounds = [0, 3000, 3000, n]                  # rank 1 is empty
    cent = oracle.pq_train(rows[:2000], d, 4, 16, 2, 42)
    codes = oracle.pq_encode_batch(cent, rows)
    segs, comms, bases = [], [], (C.c_uint64 * world)()
    try:
        for r in range(world):
            seg = vs.Segment.upload(rows[bounds[r]:bounds[r + 1]].reshape(-1, d), id_base=bounds[r])
            if seg.n:
                seg.attach_pq(cent, codes[bounds[r]:bounds[r + 1]])
            segs.append(seg)
            comm, hd = C.c_uint64(0), (C.c_uint8 * 64)()
            L.check(lib.vs_peer_create(r, world, 1 << 16, 4, C.byref(comm), hd))
            comms.append(comm.value)
            b = C.c_uint64(0)
            L.check(lib.vs_peer_base(comm.value, C.byref(b)))
            bases[r] = b.value
        for r in range(world):
            L.check(lib.vs_peer_connect_ptrs(comms[r], bases))
        out, errs = [None] * world, []

        def run(r):
            try:
                ids, sc, cn = np.zeros((1, k), np.int64), np.zeros((1, k), np.float64), np.zeros(1, np.int32)
                L.check(lib.vs_bruteforce_topk_exchange(segs[r].handle, comms[r], q.ctypes.data_as(L.f32p), 1, k, 0,
                                                        ids.ctypes.data_as(L.i64p), sc.ctypes.data_as(L.f64p), cn.ctypes.data_as(L.i32p)))
                ai, asc, acn = np.zeros((1, k), np.int64), np.zeros((1, k), np.float64), np.zeros(1, np.int32)
                L.check(lib.vs_adc_rerank_topk_exchange(segs[r].handle, comms[r], q.ctypes.data_as(L.f32p), 1, 50, k, 0, 0,
                                                        ai.ctypes.data_as(L.i64p), asc.ctypes.data_as(L.f64p), acn.ctypes.data_as(L.i32p)))
                out[r] = (ids[0], sc[0], ai[0], asc[0])
            except Exception as e:  # noqa: BLE001
                errs.append(e)

        th = [threading.Thread(target=run, args=(r,)) for r in range(world)]
        for t in th:
            t.start()
        for t in th:
            t.join(timeout=120)
        assert not errs, errs
        oi, os_, _ = oracle.bruteforce_topk(rows, q, k)
        ci, _ = oracle.adc_topn(oracle.build_lut(cent, q), codes, 50)
        ri, rs, _ = oracle.rerank_topk(rows, q, ci, k)
        for r in range(world):
            assert np.array_equal(out[r][0], oi) and _same(out[r][1], os_)
            assert np.array_equal(out[r][2], ri) and _same(out[r][3], rs)
    finally:
        for c in comms:
            lib.vs_peer_destroy(c)
        for seg in segs:
            seg.free()


# ---- vs_segment_free racing a query -----------------------------------------------------------------------------------
def test_segment_free_while_queries_run(vs, oracle):
    """a query holds its own reference, so freeing the handle on another thread either lets the
    query finish on live memory or makes the NEXT call fail with VS_EHANDLE -- never a use-after-free."""
    from vectorsearch_b200 import _lib as L

    rows = oracle.gen_rows(5, 0, 200_000, 64)
    q = oracle.gen_floats(6, 0, 64)
    want, _, _ = oracle.bruteforce_topk(rows, q, 10, threads=4)
    for rep in range(4):
        seg = vs.Segment.upload(rows)
        stop, bad = threading.Event(), []

        def worker():
            while not stop.is_set():
                try:
                    ids, _ = seg.bruteforce_topk(q, 10)
                    if not np.array_equal(ids, want):
                        bad.append(ids)
                except L.VsError as e:
                    if e.code != L.VS_EHANDLE:
                        bad.append(e)
                    return

        th = [threading.Thread(target=worker) for _ in range(3)]
        for t in th:
            t.start()
        import time

        time.sleep(0.02 * (rep + 1))
        L.check(L.load().vs_segment_free(seg.handle))
        seg.handle = 0
        for t in th:
            t.join(timeout=60)
        stop.set()
        assert not bad, bad


# ---- wire formats -----------------------------------------------------------------------------------------------------
def _proto_classes():
    """The reference's messages (src/main/proto/vectorsearch.proto:108-142) declared to the real protobuf runtime."""
    from google.protobuf import descriptor_pb2, descriptor_pool, message_factory

    fdp = descriptor_pb2.FileDescriptorProto()
    fdp.name, fdp.package, fdp.syntax = "vs_wire_test.proto", "vswire", "proto3"
    T = descriptor_pb2.FieldDescriptorProto
    cb = fdp.message_type.add()
    cb.name = "PQCodebook"
    for name, num, typ, lab in (("m", 1, T.TYPE_INT32, T.LABEL_OPTIONAL), ("k", 2, T.TYPE_INT32, T.LABEL_OPTIONAL),
                                ("centroids", 3, T.TYPE_BYTES, T.LABEL_REPEATED)):
        f = cb.field.add()
        f.name, f.number, f.type, f.label = name, num, typ, lab
    vr = fdp.message_type.add()
    vr.name = "VectorRecord"
    for name, num, typ in (("seg_id", 1, T.TYPE_INT32), ("vec_id", 2, T.TYPE_INT32), ("embedding", 3, T.TYPE_BYTES),
                           ("deleted", 4, T.TYPE_BOOL), ("payload", 5, T.TYPE_BYTES)):
        f = vr.field.add()
        f.name, f.number, f.type, f.label = name, num, typ, T.LABEL_OPTIONAL
    pool = descriptor_pool.DescriptorPool()
    pool.Add(fdp)
    return (message_factory.GetMessageClass(pool.FindMessageTypeByName("vswire.PQCodebook")),
            message_factory.GetMessageClass(pool.FindMessageTypeByName("vswire.VectorRecord")))


def test_upload_from_stored_bytes_and_records(vs, oracle):
    """Rows consumed as the reference stores them: FloatPacker bytes at a stride, and serialized VectorRecord
    messages (embedding -> row, deleted -> skipped, vec_id reported)."""
    _, VectorRecord = _proto_classes()
    n, d, k = 3000, 24, 7
    rows = oracle.gen_rows(42, 0, n, d)
    q = oracle.gen_floats(43, 0, d)
    packed = b"".join(oracle.floats_to_bytes(r) for r in rows[:50]) + rows[50:].tobytes()
    a = vs.Segment.upload_bytes(packed, n, d)
    stride = d * 4 + 13                          # embeddings embedded in larger records
    blob = bytearray(stride * n)
    for i in range(n):
        blob[i * stride:i * stride + d * 4] = rows[i].tobytes()
    b = vs.Segment.upload_bytes(bytes(blob), n, d, stride=stride, id_base=77)
    rng = np.random.default_rng(3)
    deleted = rng.random(n) < 0.1
    recs = []
    for i in range(n):
        r = VectorRecord(seg_id=9, vec_id=100 + i, embedding=oracle.floats_to_bytes(rows[i]))
        if deleted[i]:
            r.deleted = True
        if i % 3 == 0:
            r.payload = bytes(rng.integers(0, 256, size=int(rng.integers(0, 300)), dtype=np.uint8))
        recs.append(r.SerializeToString())
    c, vec_ids = vs.Segment.upload_records(recs, d)
    try:
        assert np.array_equal(vec_ids, 100 + np.arange(n))
        assert np.array_equal(a.rows().view(np.uint32), rows.view(np.uint32))
        assert np.array_equal(b.rows().view(np.uint32), rows.view(np.uint32))
        assert np.array_equal(c.rows().view(np.uint32), rows.view(np.uint32))
        oi, os_, _ = oracle.bruteforce_topk(rows, q, k)
        for seg, base in ((a, 0), (b, 77)):
            gi, gs = seg.bruteforce_topk(q, k)
            assert np.array_equal(gi, oi + base) and _same(gs, os_)
        oi, os_, _ = oracle.bruteforce_topk(rows, q, k, skip=deleted.astype(np.uint8))
        gi, gs = c.bruteforce_topk(q, k)
        assert np.array_equal(gi, oi) and _same(gs, os_)
        with pytest.raises(ValueError):          # wrong dimension: the reference would mis-read, we refuse
            vs.Segment.upload_records(recs[:3], d + 1)
        with pytest.raises(ValueError):          # truncated message
            vs.Segment.upload_records([recs[0][:-3]], d)
    finally:
        for s in (a, b, c):
            s.free()


def test_codebook_wire_format_and_sealing_from_bytes(vs, oracle):
    """vs_codebook_encode emits the bytes of PQCodebook.toByteArray() (SegmentBuildService.java:325-338); decode
    follows SegmentCaches.decodeCodebook (:141-162); a segment seals from the stored message."""
    PQCodebook, _ = _proto_classes()
    n, d, M, K = 4000, 32, 4, 16
    rows = oracle.gen_rows(42, 0, n, d)
    cent = oracle.pq_train(rows, d, M, K, 2, 42)
    msg = PQCodebook(m=M, k=K)
    for s in range(M):
        msg.centroids.append(b"".join(oracle.floats_to_bytes(cent[s, ci]) for ci in range(K)))
    want = msg.SerializeToString()
    assert vs.codebook_encode(cent) == want
    assert np.array_equal(vs.codebook_decode(want).view(np.uint32), cent.view(np.uint32))
    big = oracle.gen_rows(7, 0, 16 * 256 * 8, 1).reshape(16, 256, 8)  # the production shape: 2-byte varints
    msg = PQCodebook(m=16, k=256)
    for s in range(16):
        msg.centroids.append(big[s].tobytes())
    assert vs.codebook_encode(big) == msg.SerializeToString()
    assert np.array_equal(vs.codebook_decode(msg.SerializeToString()), big)
    for bad in (want[:-5], b"\x08\x04\x10\x10", PQCodebook(m=M, k=K, centroids=[b"1234"] * M).SerializeToString()):
        with pytest.raises(ValueError):
            vs.codebook_decode(bad)
    seg = vs.Segment.upload(rows)
    try:
        seg.attach_pq_codebook(want)             # codes encoded on the device
        assert (seg.M, seg.K) == (M, K)
        assert np.array_equal(seg.codes(), oracle.pq_encode_batch(cent, rows, threads=4))
    finally:
        seg.free()


def test_residency_table(vs, oracle):
    """(segment id, SegmentMeta.State) -> resident handle: upload on PENDING, attach PQ on SEALED, invalidate on
    compaction, least-recently-used eviction under a byte budget."""
    from vectorsearch_b200 import _lib as L
    from vectorsearch_b200.ops import STATE_PENDING, STATE_SEALED

    R = vs.Residency
    d = 16
    rows = oracle.gen_rows(42, 0, 3000, d)
    cent = oracle.pq_train(rows[:1000], d, 4, 8, 1, 42)
    try:
        R.set_budget(0)
        assert R.get(5, STATE_PENDING) is None
        s5 = vs.Segment.upload(rows[:1000])
        R.put(5, STATE_PENDING, s5)
        got = R.get(5, STATE_PENDING)
        assert got.handle == s5.handle and got.n == 1000
        with pytest.raises(L.VsError) as e:      # sealed since: the PENDING copy has no codes
            R.get(5, STATE_SEALED)
        assert e.value.code == L.VS_ESTATE
        seg, ok = R.peek(5, STATE_SEALED)
        assert seg.handle == s5.handle and not ok
        seg.attach_pq(cent)                      # upgrade in place, register the new state
        R.put(5, STATE_SEALED, seg)
        assert R.get(5, STATE_SEALED).M == 4 and R.stats()["segments"] == 1
        # replacing a segment frees the stale copy
        s5b = vs.Segment.upload(rows[:500])
        R.put(5, STATE_PENDING, s5b)
        assert L.load().vs_segment_info(s5.handle, None, None, None, None, None) == L.VS_EHANDLE
        # budget: room for two of the three 1000-row segments
        for sid in (6, 7):
            R.put(sid, STATE_PENDING, vs.Segment.upload(rows[:1000]))
        assert R.stats()["segments"] == 3
        R.get(5, STATE_PENDING)                  # 5 is now more recent than 6
        R.set_budget(2 * 1000 * d * 4 + 500 * d * 4 - 1)
        R.put(8, STATE_PENDING, vs.Segment.upload(rows[:1000]))
        assert R.get(6, STATE_PENDING) is None and R.get(8, STATE_PENDING) is not None
        # compaction rebuild (MaintenanceService.java:388-390): drop the copy
        h8 = R.get(8, STATE_PENDING).handle
        R.invalidate(8)
        assert R.get(8, STATE_PENDING) is None
        assert L.load().vs_segment_info(h8, None, None, None, None, None) == L.VS_EHANDLE
        R.invalidate(8)                          # idempotent
    finally:
        R.set_budget(0)
        for sid in (5, 6, 7, 8):
            R.invalidate(sid)
    assert R.stats() == {"segments": 0, "bytes": 0}


# ---- expansion scoring --------------------------------------------------------------------------------------------------
def test_adc_gather_matches_oracle(vs, oracle):
    """pqApproxDistance of frontier ids against resident codes (J/fdb/FdbVectorIndex.java:950-963): bit-exact,
    codes >= K skipped (:1061), ids without a code reported invalid."""
    n, d, M, K = 20_000, 64, 8, 200
    rows = oracle.gen_rows(42, 0, n, d)
    cent = oracle.pq_train(rows[:3000], d, M, K, 2, 42)
    codes = oracle.pq_encode_batch(cent, rows, threads=4)
    codes[100:200, 3] = 255                      # stored codes beyond K contribute nothing
    seg = vs.Segment.upload(rows, id_base=500)
    try:
        seg.attach_pq(cent, codes)
        rng = np.random.default_rng(1)
        for qi in range(3):
            q = oracle.gen_floats(43, qi * d, d)
            lut = oracle.build_lut(cent, q)
            with seg.adc_query(q) as aq:
                for m in (1, 64, 300, 6000):
                    ids = rng.integers(-50, n + 50, size=m).astype(np.int64)
                    ids[: min(m, 100)] = np.arange(100, 100 + min(m, 100))
                    dist, ok = aq.gather(ids + 500)
                    want_ok = (ids >= 0) & (ids < n)
                    assert np.array_equal(ok, want_ok)
                    want = np.array([oracle.pq_approx_distance(lut, codes[i], K) if o else np.nan for i, o in zip(ids, want_ok)])
                    assert _same(dist, want)
            d1, o1 = seg.adc_gather(q, np.arange(500, 564))
            assert o1.all() and _same(d1, np.array([oracle.pq_approx_distance(lut, codes[i], K) for i in range(64)]))
    finally:
        seg.free()


# ---- graph construction distances ----------------------------------------------------------------------------------------
@pytest.mark.parametrize("n,d,degree,l_build,alpha", [(3000, 64, 16, 0, 1.0), (2500, 128, 32, 0, 1.0), (700, 24, 8, 0, 1.0),
                                                      (40, 16, 64, 0, 1.0), (2000, 64, 12, 40, 1.2), (1500, 32, 8, 24, 1.0),
                                                      (900, 100, 10, 0, 1.0)])
def test_knn_graph_matches_graph_builder(vs, oracle, n, d, degree, l_build, alpha):
    """vs_knn_graph = GraphBuilder.buildL2Neighbors / buildPrunedNeighbors (J/graph/GraphBuilder.java:41-109) over the
    segment's rows: identical lists (order and ties), including nodes whose candidate list had to be redone exactly."""
    rows = oracle.gen_rows(31, 0, n, d)
    if n >= 700:
        rows[100:160] = rows[7]          # 61 identical rows: runs of equal distance longer than the nomination margin
        rows[300] = rows[301]            # a pair at distance 0
    seg = vs.Segment.upload(rows)
    try:
        got = seg.knn_graph(degree, l_build, alpha)
        want = oracle.knn_graph(rows, degree, l_build, alpha, threads=8)
        assert len(got) == n
        bad = [i for i in range(n) if not np.array_equal(got[i], want[i])]
        assert not bad, (bad[:5], got[bad[0]], want[bad[0]])
    finally:
        seg.free()


def test_knn_graph_edge_cases(vs, oracle):
    from vectorsearch_b200 import _lib as L

    one = vs.Segment.upload(oracle.gen_rows(1, 0, 1, 8))
    assert [a.tolist() for a in one.knn_graph(4)] == [[]]
    one.free()
    rows = oracle.gen_rows(2, 0, 600, 32)
    rows[17] = np.nan                        # NaN distances sort last (Double.compare) and still fill short lists
    seg = vs.Segment.upload(rows)
    try:
        got = seg.knn_graph(6)
        want = oracle.knn_graph(rows, 6, threads=4)
        assert all(np.array_equal(a, b) for a, b in zip(got, want))
        seg.set_skip(np.zeros(600, np.uint8))
        with pytest.raises(L.VsError):       # a skip mask has no meaning for GraphBuilder
            seg.knn_graph(6)
    finally:
        seg.free()
