"""GPU parity: libvsgpu (through its C ABI) against the CPU oracle on identical seeded inputs.

Bar (BASELINE.json north_star): ids and PQ codes bit-exact; scores are doubles carrying the
reference's own arithmetic, so they are compared bit-exactly too (tolerance 0).
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def vs():
    import vectorsearch_b200 as v

    v.init(0)
    yield v
    v.set_simd_lanes(16)


def _same(a, b):
    """Bit-exact equality; NaN matches NaN (Java does not distinguish NaN payloads in ordering)."""
    a, b = np.asarray(a), np.asarray(b)
    if a.shape != b.shape:
        return False
    if a.dtype != np.float64:
        return np.array_equal(a, b)
    na, nb = np.isnan(a), np.isnan(b)
    return np.array_equal(na, nb) and np.array_equal(a[~na].view(np.uint64), b[~nb].view(np.uint64))


# ---- Distances: T/util/DistancesTest.java shapes, all lane models ---------------------------------------
@pytest.mark.parametrize("lanes", [16, 8, 4])
@pytest.mark.parametrize("dim", [1, 3, 7, 16, 17, 128, 768, 1000])
def test_pair_ops_bit_exact(vs, oracle, dim, lanes):
    vs.set_simd_lanes(lanes)
    oracle.set_lanes(lanes)
    try:
        a = oracle.gen_floats(42 + dim, 0, dim)
        b = oracle.gen_floats(42 + dim, dim, dim)
        D = vs.Distances
        assert D.l2_squared(a, b) == oracle.l2_squared(a, b)
        assert D.l2(a, b) == oracle.l2(a, b)
        assert D.dot(a, b) == oracle.dot(a, b)
        assert D.norm(a) == oracle.norm(a)
        assert D.cosine(a, b) == oracle.cosine(a, b)
    finally:
        vs.set_simd_lanes(16)
        oracle.set_lanes(16)


def test_known_answers(vs):
    D = vs.Distances  # DistancesTest.java:36-47, :100-125
    assert D.l2_squared([1, 2, 3], [4, 6, 3]) == 25.0
    assert D.l2([1, 2, 3], [4, 6, 3]) == 5.0
    assert D.cosine([0, 0, 0], [1, 2, 3]) == 0.0
    assert D.dot([1, 0, 0, 0], [0, 1, 0, 0]) == 0.0
    assert D.l2_squared([0, 1, 2, 3, 4], [9, 9, 2, 3, 9], 2, 2, 2) == 0.0  # offset overload :77-94
    cent = np.array([[[0, 0], [1, 1]], [[0, 0], [2, 2]]], dtype=np.float32)  # PqEncoderTest.java:12-23
    assert vs.PqEncoder.encode(cent, [0.1, 0.1, 1.9, 2.1]).tolist() == [0, 1]
    with pytest.raises(ValueError):  # PqTrainerTest.java: dimension % m != 0
        vs.PqTrainer.train(np.zeros((3, 3), np.float32), 3, 2, 2, 1, 1)
    with pytest.raises(IndexError):
        vs.PqTrainer.train(np.zeros((0, 4), np.float32), 4, 2, 2, 1, 1)


def test_jmh_state_values(vs, oracle):
    from tests.test_oracle_golden import jmh_pq_state

    cent, vec, lut, codes = jmh_pq_state(oracle)
    assert vs.PqEncoder.encode(cent, vec).tolist() == [81, 78, 7, 3, 2, 166, 155, 121, 147, 65, 197, 171,
                                                      102, 180, 200, 175]
    assert abs(vs.pq_lut_distance(lut, codes) - 90.38559) < 5e-5
    assert _same(vs.build_lut(cent, vec), oracle.build_lut(cent, vec))


def test_generator_matches_java_random(vs, oracle):
    seg = vs.Segment.generate(42, 1000, 3001, 24)
    try:
        assert np.array_equal(seg.rows(), oracle.gen_rows(42, 1000, 3001, 24))
    finally:
        seg.free()


# ---- brute force ----------------------------------------------------------------------------------------
def _check_bruteforce(vs, oracle, rows, q, k, metric, skip=None, id_base=0):
    seg = vs.Segment.upload(rows, skip=skip, id_base=id_base)
    try:
        ids, sc = seg.bruteforce_topk(q, k, metric)
    finally:
        seg.free()
    oi, os_, _ = oracle.bruteforce_topk(rows, q, k, metric, skip=skip, threads=4)
    assert np.array_equal(ids, oi + id_base), (ids[:10], oi[:10])
    assert _same(sc, os_)


@pytest.mark.parametrize("metric", [0, 1])
@pytest.mark.parametrize("n,d,k", [(20000, 128, 10), (5000, 768, 50), (3000, 100, 10), (4000, 6, 7),
                                   (1000, 16, 1000), (7, 128, 10), (33, 32, 1), (50000, 64, 100)])
def test_bruteforce_matches_oracle(vs, oracle, n, d, k, metric):
    rows = oracle.gen_rows(42, 0, n, d)
    q = oracle.gen_floats(43, 0, d)
    _check_bruteforce(vs, oracle, rows, q, k, metric)


@pytest.mark.parametrize("lanes", [8, 4])
@pytest.mark.parametrize("metric", [0, 1])
def test_bruteforce_other_lane_models(vs, oracle, lanes, metric):
    vs.set_simd_lanes(lanes)
    oracle.set_lanes(lanes)
    try:
        rows = oracle.gen_rows(5, 0, 6000, 72)
        q = oracle.gen_floats(6, 0, 72)
        _check_bruteforce(vs, oracle, rows, q, 20, metric)
    finally:
        vs.set_simd_lanes(16)
        oracle.set_lanes(16)


@pytest.mark.parametrize("metric", [0, 1])
def test_bruteforce_ties_nan_skip(vs, oracle, metric):
    rows = oracle.gen_rows(11, 0, 4096, 32)
    rows[100:400] = rows[7]          # 300 exact duplicates: ties resolve to the lowest row
    rows[2000] = rows[7]
    rows[50, 3] = np.nan             # NaN score sorts first (Double.compare)
    rows[3000, 0] = np.nan
    rows[9] = 0.0                    # zero norm: cosine 0.0
    q = rows[7].copy()
    _check_bruteforce(vs, oracle, rows, q, 40, metric)
    skip = np.zeros(4096, np.uint8)
    skip[[7, 50, 101, 102, 4095]] = 1  # deleted / gid missing
    _check_bruteforce(vs, oracle, rows, q, 40, metric, skip=skip, id_base=1_000_000_007)
    _check_bruteforce(vs, oracle, rows, np.zeros(32, np.float32), 5, metric)  # zero query


@pytest.mark.parametrize("d,lanes", [(100, 16), (20, 8), (70, 16), (6, 4)])
def test_bruteforce_nan_only_in_the_fp64_tail(vs, oracle, d, lanes):
    """with d % lanes != 0 the L2 pre-filter sees the fp32 lane part over [0, ub) only; a row whose ONLY NaN
    sits in the fp64 tail [ub, d) must still surface (its score is NaN, which Double.compare sorts first)."""
    vs.set_simd_lanes(lanes)
    oracle.set_lanes(lanes)
    try:
        n = 30_000
        rows = oracle.gen_rows(42, 0, n, d)
        ub = d - d % lanes
        rows[17_000, ub:] = np.nan               # the whole tail
        rows[25_001, d - 1] = np.nan             # its last element only
        rows[29_999, ub] = np.inf                # an infinite distance is NOT NaN: it sorts last, not first
        q = oracle.gen_floats(43, 0, d)
        seg = vs.Segment.upload(rows)
        try:
            for k in (1, 10, 40):
                ids, sc = seg.bruteforce_topk(q, k)
                oi, os_, _ = oracle.bruteforce_topk(rows, q, k)
                assert np.array_equal(ids, oi) and _same(sc, os_), (k, ids[:4], oi[:4])
            assert set(seg.bruteforce_topk(q, 2)[0].tolist()) == {17_000, 25_001}
        finally:
            seg.free()
    finally:
        vs.set_simd_lanes(16)
        oracle.set_lanes(16)


def test_bruteforce_query_batch(vs, oracle):
    rows = oracle.gen_rows(42, 0, 8000, 128)
    qs = oracle.gen_rows(43, 0, 5, 128)
    seg = vs.Segment.upload(rows)
    try:
        ids, sc, cn = seg.bruteforce_topk(qs, 10)
        for i in range(5):
            oi, os_, _ = oracle.bruteforce_topk(rows, qs[i], 10)
            assert cn[i] == 10 and np.array_equal(ids[i], oi) and _same(sc[i], os_)
        # VectorIndexTest.java:599-609 + query merge :432-437
    finally:
        seg.free()
    seg = vs.Segment.upload(np.array([[1, 0, 0], [0, 1, 0], [0, 0, 1], [1, 1, 0]], np.float32))
    try:
        ids, sc = seg.bruteforce_topk([1, 0, 0], 4)
        mi, ms = vs.merge_topk(ids, sc, 2)
        assert mi.tolist() == [0, 3] and ms[0] == 0.0 and ms[1] == -1.0
    finally:
        seg.free()


# ---- batched queries: tensor-core nomination + exact re-score (batch.cu) ----------------------------------
@pytest.fixture(params=["fp16", "tf32", "fp16-warpselect", "fp16-g16", "tf32-g32", "fp16-g64", "fp16-pairs", "tf32-pairs"])
def force_batch(vs, request):
    """Route every query batch of >= 2 queries through batch.cu, whatever the segment size, nominating
    on the fp16 operand copies (default) or on the fp32 rows read as tf32; "warpselect" also forces the
    one-warp-per-query selection kernel that large batches use."""
    vs.set_option("batch_min_queries", 2)
    vs.set_option("batch_min_rows", 1)
    vs.set_option("batch_fp16", 0 if request.param == "tf32" else 1)
    vs.set_option("batch_warp_min_queries", 2 if "warpselect" in request.param else 0)
    vs.set_option("batch_group", int(request.param.split("-g")[1]) if "-g" in request.param else 0)  # rows per group
    vs.set_option("batch_pairs", 1 if "pairs" in request.param else 0)  # cta_group::2 nomination kernel
    yield
    vs.set_option("batch_group", 0)
    vs.set_option("batch_pairs", 2)  # automatic: long vectors only
    vs.set_option("batch_min_queries", 3)
    vs.set_option("batch_min_rows", 16384)
    vs.set_option("batch_fp16", 1)
    vs.set_option("batch_warp_min_queries", 0)


def _check_batch(vs, oracle, rows, qs, k, metric, skip=None, id_base=0, threads=4):
    seg = vs.Segment.upload(rows, skip=skip, id_base=id_base)
    try:
        ids, sc, cn = seg.bruteforce_topk(qs, k, metric)
        for i in range(qs.shape[0]):
            oi, os_, _ = oracle.bruteforce_topk(rows, qs[i], k, metric, skip=skip, threads=threads)
            c = len(oi)
            assert cn[i] == c, (i, cn[i], c)
            assert np.array_equal(ids[i, :c], oi + id_base), f"query {i}: ids differ: {ids[i, :8]} vs {oi[:8] + id_base}"
            assert _same(sc[i, :c], os_), f"query {i}: scores differ"
    finally:
        seg.free()


@pytest.mark.parametrize("metric", [0, 1])
@pytest.mark.parametrize("n,d,k,nq", [(40000, 128, 10, 37), (20000, 768, 50, 130), (9000, 72, 10, 5), (6000, 100, 100, 3),
                                      (300, 32, 1, 2), (7, 128, 10, 4), (70000, 64, 33, 260), (1000, 128, 1000, 3),
                                      (33000, 128, 10, 300), (500, 96, 5, 129)])
def test_batch_bruteforce_matches_oracle(vs, oracle, force_batch, n, d, k, nq, metric):
    rows = oracle.gen_rows(42, 0, n, d)
    qs = oracle.gen_rows(43, 0, nq, d)
    _check_batch(vs, oracle, rows, qs, k, metric)


@pytest.mark.parametrize("lanes", [8, 4])
@pytest.mark.parametrize("metric", [0, 1])
def test_batch_other_lane_models(vs, oracle, force_batch, lanes, metric):
    vs.set_simd_lanes(lanes)
    oracle.set_lanes(lanes)
    try:
        rows = oracle.gen_rows(5, 0, 12000, 72)
        qs = oracle.gen_rows(6, 0, 9, 72)
        _check_batch(vs, oracle, rows, qs, 20, metric)
    finally:
        vs.set_simd_lanes(16)
        oracle.set_lanes(16)


@pytest.mark.parametrize("metric", [0, 1])
def test_batch_ties_skip_and_fallbacks(vs, oracle, force_batch, metric):
    n, d = 30000, 64
    rows = oracle.gen_rows(11, 0, n, d)
    rows[100:400] = rows[7]           # 300 exact duplicates: ties resolve to the lowest row
    rows[20000] = rows[7]
    rows[9] = 0.0                     # zero norm: cosine 0.0
    rows[5000:5004] *= 3.0            # same direction, larger norm
    qs = oracle.gen_rows(12, 0, 6, d)
    qs[0] = rows[7]
    qs[1] = 0.0                       # zero query: every cosine is 0.0, L2 = |x|
    skip = np.zeros(n, np.uint8)
    skip[[7, 101, 102, n - 1]] = 1
    _check_batch(vs, oracle, rows, qs, 40, metric)
    _check_batch(vs, oracle, rows, qs, 40, metric, skip=skip, id_base=1_000_000_007)
    # candidate-list overflow: 20000 duplicates of the best row -> that query falls back to the exact scan
    rows2 = rows.copy()
    rows2[2000:22000] = rows2[7]
    _check_batch(vs, oracle, rows2, qs, 10, metric)
    # non-finite query -> exact scan for that query only; NaN row -> the whole segment stays per-query
    qs2 = qs.copy()
    qs2[3, 5] = np.nan
    qs2[4, 0] = np.inf
    _check_batch(vs, oracle, rows, qs2, 10, metric)
    rows3 = rows.copy()
    rows3[50, 3] = np.nan
    rows3[3000, 0] = np.inf
    _check_batch(vs, oracle, rows3, qs, 10, metric)


def test_batch_skip_update_invalidates_coefficients(vs, oracle, force_batch):
    rows = oracle.gen_rows(21, 0, 5000, 32)
    qs = oracle.gen_rows(22, 0, 4, 32)
    seg = vs.Segment.upload(rows)
    try:
        ids0, _, _ = seg.bruteforce_topk(qs, 5)
        skip = np.zeros(5000, np.uint8)
        skip[ids0[:, 0]] = 1
        seg.set_skip(skip)
        ids1, sc1, _ = seg.bruteforce_topk(qs, 5)
        for i in range(4):
            oi, os_, _ = oracle.bruteforce_topk(rows, qs[i], 5, 0, skip=skip)
            assert np.array_equal(ids1[i], oi) and _same(sc1[i], os_)
        seg.set_skip(None)
        ids2, _, _ = seg.bruteforce_topk(qs, 5)
        assert np.array_equal(ids2, ids0)
    finally:
        seg.free()


def test_c2_full_size_batch_1024(vs, oracle):
    """C2, query batch 1024: batched results equal the per-query kernel's (all 1024) and the oracle's (a sample)."""
    n, d, nq = 1_000_000, 128, 1024
    seg = vs.Segment.generate(42, 0, n, d)
    try:
        qs = oracle.gen_rows(43, 0, nq, d)
        ids, sc, cn = seg.bruteforce_topk(qs, 10)
        assert np.all(cn == 10) and np.all(np.diff(sc, axis=1) <= 0)
        for i in range(0, nq, 16):
            si, ss = seg.bruteforce_topk(qs[i], 10)     # single query: scan.cu
            assert np.array_equal(ids[i], si) and _same(sc[i], ss)
        rows = oracle.gen_rows(42, 0, n, d)
        for i in (0, 511, 1023):
            oi, os_, _ = oracle.bruteforce_topk(rows, qs[i], 10, threads=8)
            assert np.array_equal(ids[i], oi) and _same(sc[i], os_)
    finally:
        seg.free()


@pytest.mark.parametrize("mode", ["fp16", "tf32"])
@pytest.mark.parametrize("metric", [0, 1])
@pytest.mark.parametrize("n,d,scale", [(20000, 128, 1.0), (6000, 768, 1.0), (8000, 96, 300.0), (8000, 64, 1e-4)])
def test_batch_nomination_bound(vs, oracle, mode, metric, n, d, scale):
    """The tensor-core stage may deviate from the exact a(q, x) by at most the slack the selection adds
    (that is what makes the nominated set a superset of the true top-k).  Measured here against float64."""
    import ctypes as C

    from vectorsearch_b200 import _lib as L

    vs.set_option("batch_fp16", 1 if mode == "fp16" else 0)
    try:
        rows = (oracle.gen_rows(7, 0, n, d) * np.float32(scale)).astype(np.float32)
        rows[::7] *= np.float32(0.25)                       # mixed norms
        qs = (oracle.gen_rows(8, 0, 16, d) * np.float32(scale)).astype(np.float32)
        qs[3] *= np.float32(1e-3)                           # a query much smaller than the rows
        seg = vs.Segment.upload(rows)
        try:
            lib = vs.load()
            ng, g = C.c_int64(), C.c_int32()
            L.check(lib.vs_debug_batch_groupmins(seg.handle, qs.ctypes.data_as(L.f32p), 16, metric, None, 0, C.byref(ng), C.byref(g), None))
            gm = np.zeros((16, ng.value), np.float32)
            slack = np.zeros(16, np.float64)
            L.check(lib.vs_debug_batch_groupmins(seg.handle, qs.ctypes.data_as(L.f32p), 16, metric, gm.ctypes.data_as(L.f32p),
                                                 gm.size, C.byref(ng), C.byref(g), slack.ctypes.data_as(L.f64p)))
        finally:
            seg.free()
        X, Q = rows.astype(np.float64), qs.astype(np.float64)
        dots = Q @ X.T
        xx = (X * X).sum(1)
        a = (xx[None, :] - 2 * dots) if metric == 0 else -dots / np.sqrt(xx)[None, :]
        pad = (-n) % g.value
        a = np.pad(a, ((0, 0), (0, pad)), constant_values=np.inf).reshape(16, -1, g.value).min(2)
        err = np.abs(gm.astype(np.float64) - a)
        worst = float((err / slack[:, None]).max())
        assert worst <= 1.0, f"nomination error exceeds the slack: {worst:.3f} of the bound"
        assert worst >= 1e-4 or d < 64, "suspiciously exact: is the tensor-core stage running?"
        print(f"nomination error / slack: max {worst:.4f} ({mode}, metric {metric}, n={n}, d={d}, scale={scale})")
    finally:
        vs.set_option("batch_fp16", 1)


def test_sharded_coordinator_pipelined_streams(vs, oracle):
    """The multi-GPU coordinator alternates independent queries between two streams (one libvsgpu scratch set
    per stream); with one rank the collective drops out and the rest is testable here.  Results must equal
    the oracle's for every query, whatever overlapped."""
    import torch

    from vectorsearch_b200.sharded import ShardedSegment

    n, d, nq = 300_000, 128, 24
    rows = oracle.gen_rows(42, 0, n, d)
    qs = oracle.gen_rows(43, 0, nq, d)
    seg = vs.Segment.upload(rows, id_base=5_000_000)
    try:
        sh = ShardedSegment(seg, 0, 1)
        q_dev = torch.from_numpy(qs).cuda()
        torch.cuda.synchronize()
        kept = []
        for i in range(nq):
            ids, sc, cn, stream = sh.bruteforce_topk_pipelined(q_dev[i:i + 1], 1, 10)
            with torch.cuda.stream(stream):
                kept.append((ids.clone(), sc.clone(), cn.clone()))
        sh.drain()
        torch.cuda.synchronize()
        for i in range(nq):
            oi, os_, _ = oracle.bruteforce_topk(rows, qs[i], 10, threads=4)
            assert kept[i][2].item() == 10
            assert np.array_equal(kept[i][0].cpu().numpy()[0], oi + 5_000_000) and _same(kept[i][1].cpu().numpy()[0], os_)
    finally:
        seg.free()


@pytest.mark.parametrize("metric", [0, 1])
@pytest.mark.parametrize("shards", [1, 2, 3])
def test_adc_rerank_across_shards(vs, oracle, shards, metric):
    """C4 on several GPUs, emulated with row-range shards resident on this one: every shard packs its ADC
    candidates with their exact scores, the packs are concatenated in rank order (what the all-gather
    produces) and merged.  Must equal the reference's single-segment result: global first n_cand by
    approximate distance, re-ranked exactly (FdbVectorIndex.java:769,820-828,997-1043)."""
    import torch

    from vectorsearch_b200 import _lib as L
    from vectorsearch_b200.sharded import merge_adc_rerank_host, shard_range

    n, d, M, K, n_cand, k, nq = 60000, 64, 8, 64, 100, 10, 3
    rows = oracle.gen_rows(42, 0, n, d)
    rows[1000:1040] = rows[17]                      # duplicate rows: equal approximate AND exact distances
    rows[40000:40030] = rows[17]
    qs = oracle.gen_rows(43, 0, nq, d)
    qs[0] = rows[17]
    skip = np.zeros(n, np.uint8)
    skip[[17, 1001, 40001, 59999]] = 1              # deleted rows keep their candidate slot, are not scored
    cent = oracle.pq_train(rows[:5000], d, M, K, 3, 42)
    codes = oracle.pq_encode_batch(cent, rows, threads=4)
    lib = vs.load()
    st = torch.cuda.current_stream().cuda_stream
    segs, packs = [], []
    try:
        q_dev = torch.from_numpy(qs).cuda()
        for r in range(shards):
            lo, hi = shard_range(n, r, shards)
            seg = vs.Segment.upload(rows[lo:hi], skip=skip[lo:hi], id_base=lo)
            seg.attach_pq(cent, codes[lo:hi])
            segs.append(seg)
            pack = torch.empty((nq, 4, n_cand), dtype=torch.int64, device="cuda")
            L.check(lib.vs_adc_rerank_packed_dev(seg.handle, q_dev.data_ptr(), nq, n_cand, metric, 0, pack.data_ptr(), st))
            packs.append(pack)
        gath = torch.stack(packs).contiguous()
        ids = torch.empty((nq, k), dtype=torch.int64, device="cuda")
        sc = torch.empty((nq, k), dtype=torch.float64, device="cuda")
        cn = torch.empty((nq,), dtype=torch.int32, device="cuda")
        L.check(lib.vs_merge_adc_rerank_packed_dev(gath.data_ptr(), shards, nq, n_cand, k, ids.data_ptr(), sc.data_ptr(), cn.data_ptr(), st))
        torch.cuda.synchronize()
        ids, sc, cn, gath = ids.cpu().numpy(), sc.cpu().numpy(), cn.cpu().numpy(), gath.cpu().numpy()
        for i in range(nq):
            ci, _ = oracle.adc_topn(oracle.build_lut(cent, qs[i]), codes, n_cand)
            ri, rs, _ = oracle.rerank_topk(rows, qs[i], ci, k, metric, skip=skip)
            c = len(ri)
            assert cn[i] == c
            assert np.array_equal(ids[i, :c], ri) and _same(sc[i, :c], rs), (i, ids[i], ri)
            hi_, hs_ = merge_adc_rerank_host(gath[:, i], k)   # the host restatement the gloo test uses
            assert np.array_equal(hi_, ri) and _same(hs_, rs)
    finally:
        for seg in segs:
            seg.free()


@pytest.mark.parametrize("world,nq,k", [(2, 1, 10), (3, 5, 7), (8, 64, 10), (4, 3, 100)])
def test_peer_exchange_merge_equals_gather_then_merge(vs, world, nq, k):
    """The peer-memory exchange (push into every rank's buffer + flag wait inside the merge kernel) must give what
    "all-gather, then vs_merge_packed_dev" gives.  Ranks are emulated by `world` communicators on this GPU, each
    on its own stream, connected by address; more rounds than a stream's ring is deep, so every slot is reused, and a
    second stream per rank every third round (its own ring)."""
    import ctypes as C

    import torch

    from vectorsearch_b200 import _lib as L

    lib = vs.load()
    depth, rounds = 12, 11  # ring 0 is the communicator's own stream; two caller streams per rank
    comms, bases = [], (C.c_uint64 * world)()
    streams = [torch.cuda.Stream() for _ in range(world)]
    streams2 = [torch.cuda.Stream() for _ in range(world)]
    rng = np.random.default_rng(world * 1000 + nq)
    try:
        for r in range(world):
            comm, hd = C.c_uint64(0), (C.c_uint8 * 64)()
            L.check(lib.vs_peer_create(r, world, 1 << 20, depth, C.byref(comm), hd))
            comms.append(comm.value)
            b = C.c_uint64(0)
            L.check(lib.vs_peer_base(comm.value, C.byref(b)))
            bases[r] = b.value
        for r in range(world):
            L.check(lib.vs_peer_connect_ptrs(comms[r], bases))
        for rnd in range(rounds):
            # per-rank packed lists: ids (some empty slots = -1), scores with ties across ranks, NaN and -0.0
            ids = rng.integers(0, 1 << 40, size=(world, nq, k)).astype(np.int64)
            sc = rng.integers(-3, 4, size=(world, nq, k)).astype(np.float64) / 2
            sc[rng.random(sc.shape) < 0.05] = np.nan
            sc[rng.random(sc.shape) < 0.05] = -0.0
            sc = -np.sort(-sc, axis=2)          # each rank's list is sorted (as a scan's is); NaN placement is free
            ids[rng.random(ids.shape) < 0.1] = -1
            pack = np.concatenate([ids, sc.view(np.int64)], axis=2)  # [world][nq][2k]
            d_pack = torch.from_numpy(pack).cuda()
            want_i = torch.empty((nq, k), dtype=torch.int64, device="cuda")
            want_s = torch.empty((nq, k), dtype=torch.float64, device="cuda")
            want_c = torch.empty((nq,), dtype=torch.int32, device="cuda")
            st0 = torch.cuda.current_stream().cuda_stream
            L.check(lib.vs_merge_packed_dev(d_pack.data_ptr(), world, nq, k, 1, want_i.data_ptr(), want_s.data_ptr(),
                                            want_c.data_ptr(), st0))
            torch.cuda.synchronize()
            outs = []
            for r in range(world):
                oi = torch.full((nq, k), -7, dtype=torch.int64, device="cuda")
                os_ = torch.zeros((nq, k), dtype=torch.float64, device="cuda")
                oc = torch.zeros((nq,), dtype=torch.int32, device="cuda")
                outs.append((oi, os_, oc))
            torch.cuda.synchronize()
            vs.set_option("peer_fused", rnd % 2)  # one query: publish inside the merge kernel / as a kernel of its own
            # ONE host thread issues every rank's exchange here, so the calls must not block: poll inside the kernels (what
            # ranks on separate GPUs do) instead of the host-side wait that communicators sharing a device default to --
            # safe in this test, nothing large runs beside the polling kernels
            vs.set_option("peer_spin_shared", 1)
            for r in range(world):               # rank r's merge spins until the later-launched ranks have published
                oi, os_, oc = outs[r]
                st = (streams2 if rnd % 3 == 2 else streams)[r].cuda_stream
                L.check(lib.vs_exchange_merge_packed_dev(comms[r], d_pack[r].data_ptr(), nq, k, 1, oi.data_ptr(),
                                                         os_.data_ptr(), oc.data_ptr(), st))
            torch.cuda.synchronize()
            for r in range(world):
                oi, os_, oc = outs[r]
                assert torch.equal(oc, want_c), (rnd, r)
                assert torch.equal(oi, want_i), (rnd, r)
                assert torch.equal(os_.view(torch.int64), want_s.view(torch.int64)), (rnd, r)
        # a third caller stream does not fit a communicator of depth 12 (its own ring + two): refused, not corrupted
        oi, os_, oc = outs[0]
        rc = lib.vs_exchange_merge_packed_dev(comms[0], d_pack[0].data_ptr(), nq, k, 1, oi.data_ptr(), os_.data_ptr(),
                                              oc.data_ptr(), torch.cuda.Stream().cuda_stream)
        assert rc != 0
    finally:
        vs.set_option("peer_fused", 1)
        vs.set_option("peer_spin_shared", 0)
        torch.cuda.synchronize()
        for c in comms:
            lib.vs_peer_destroy(c)


@pytest.mark.parametrize("world,nq", [(2, 1), (3, 4)])
def test_sharded_host_call_through_peer_exchange(vs, oracle, world, nq):
    """vs_bruteforce_topk_exchange: the whole sharded query as one host-buffer call per rank (H2D, scan, peer
    exchange, merge into pinned host memory, one synchronisation).  Ranks are host threads with one shard each on
    this GPU; every rank must return the reference's result over the concatenated rows, ties included."""
    import ctypes as C
    import threading

    from vectorsearch_b200 import _lib as L
    from vectorsearch_b200.sharded import shard_range

    lib = vs.load()
    n, d, k = 30011, 48, 10
    rows = oracle.gen_rows(42, 0, n, d)
    rows[20000:20006] = rows[5]                       # ties across shards: the lower global row wins
    qs = oracle.gen_rows(43, 0, nq, d)
    qs[0] = rows[5]
    segs, comms, bases = [], [], (C.c_uint64 * world)()
    try:
        for r in range(world):
            lo, hi = shard_range(n, r, world)
            segs.append(vs.Segment.upload(rows[lo:hi], id_base=lo))
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
                res = []
                for rep in range(6):                  # more calls than the ring is deep
                    ids = np.zeros((nq, k), np.int64)
                    sc = np.zeros((nq, k), np.float64)
                    cn = np.zeros(nq, np.int32)
                    L.check(lib.vs_bruteforce_topk_exchange(segs[r].handle, comms[r], qs.ctypes.data_as(L.f32p), nq, k, 0,
                                                            ids.ctypes.data_as(L.i64p), sc.ctypes.data_as(L.f64p),
                                                            cn.ctypes.data_as(L.i32p)))
                    res.append((ids, sc, cn))
                out[r] = res
            except Exception as e:  # noqa: BLE001
                errs.append(e)

        threads = [threading.Thread(target=run, args=(r,)) for r in range(world)]
        for th in threads:
            th.start()
        for th in threads:
            th.join(timeout=120)
        assert not errs, errs
        for i in range(nq):
            oi, os_, _ = oracle.bruteforce_topk(rows, qs[i], k)
            for r in range(world):
                for ids, sc, cn in out[r]:
                    assert cn[i] == len(oi)
                    assert np.array_equal(ids[i], oi), (r, i, ids[i], oi)
                    assert _same(sc[i], os_)
    finally:
        for c in comms:
            lib.vs_peer_destroy(c)
        for seg in segs:
            seg.free()


@pytest.mark.parametrize("world,metric", [(2, 0), (3, 1)])
def test_sharded_adc_rerank_host_call_through_peer_exchange(vs, oracle, world, metric):
    """vs_adc_rerank_topk_exchange, ranks as host threads: equals the single-segment ADC + re-rank of the reference
    (global first n_cand by approximate distance, re-ranked exactly)."""
    import ctypes as C
    import threading

    from vectorsearch_b200 import _lib as L
    from vectorsearch_b200.sharded import shard_range

    lib = vs.load()
    n, d, M, K, n_cand, k, nq = 40000, 64, 8, 64, 100, 10, 3
    rows = oracle.gen_rows(42, 0, n, d)
    rows[1000:1040] = rows[17]
    rows[30000:30030] = rows[17]
    qs = oracle.gen_rows(43, 0, nq, d)
    qs[0] = rows[17]
    cent = oracle.pq_train(rows[:5000], d, M, K, 3, 42)
    codes = oracle.pq_encode_batch(cent, rows, threads=4)
    segs, comms, bases = [], [], (C.c_uint64 * world)()
    try:
        for r in range(world):
            lo, hi = shard_range(n, r, world)
            seg = vs.Segment.upload(rows[lo:hi], id_base=lo)
            seg.attach_pq(cent, codes[lo:hi])
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
                res = []
                for rep in range(5):
                    ids = np.zeros((nq, k), np.int64)
                    sc = np.zeros((nq, k), np.float64)
                    cn = np.zeros(nq, np.int32)
                    L.check(lib.vs_adc_rerank_topk_exchange(segs[r].handle, comms[r], qs.ctypes.data_as(L.f32p), nq, n_cand, k,
                                                            metric, 0, ids.ctypes.data_as(L.i64p), sc.ctypes.data_as(L.f64p),
                                                            cn.ctypes.data_as(L.i32p)))
                    res.append((ids, sc, cn))
                out[r] = res
            except Exception as e:  # noqa: BLE001
                errs.append(e)

        threads = [threading.Thread(target=run, args=(r,)) for r in range(world)]
        for th in threads:
            th.start()
        for th in threads:
            th.join(timeout=120)
        assert not errs, errs
        for i in range(nq):
            ci, _ = oracle.adc_topn(oracle.build_lut(cent, qs[i]), codes, n_cand)
            ri, rs, _ = oracle.rerank_topk(rows, qs[i], ci, k, metric)
            for r in range(world):
                for ids, sc, cn in out[r]:
                    c = len(ri)
                    assert cn[i] == c
                    assert np.array_equal(ids[i, :c], ri) and _same(sc[i, :c], rs), (r, i, ids[i], ri)
    finally:
        for c in comms:
            lib.vs_peer_destroy(c)
        for seg in segs:
            seg.free()


def _quantisation_error(oracle, cent, rows):
    codes = oracle.pq_encode_batch(cent, rows)
    M, K, sd = cent.shape
    rec = np.concatenate([cent[s][codes[:, s]] for s in range(M)], axis=1)
    return float(((rows - rec) ** 2).sum(1).mean())


@pytest.mark.parametrize("exact", [True, False])
@pytest.mark.parametrize("world", [1, 2, 3])
def test_pq_train_sharded_ranks_emulated_with_threads(vs, oracle, world, exact):
    """C3 on several GPUs, emulated here: `world` host threads, each owning a row-range shard resident on
    this GPU, call vs_pq_train_sharded concurrently; the collective hook is a thread barrier that sums the
    ranks' buffers.  exact_order: bit-identical to the reference on any number of ranks.  One all-reduce per
    iteration: identical on all ranks, and as good a codebook (quantisation error within 2 % of the
    reference's) -- but a different Lloyd trajectory as soon as one row changes cluster."""
    import threading

    import torch

    from vectorsearch_b200.sharded import ShardedSegment, shard_range

    n, d, M, K, iters = 30000, 32, 4, 64, 5
    rows = oracle.gen_rows(42, 0, n, d)
    rows[100:160] = rows[3]          # duplicates: duplicate initial draws / empty clusters get re-initialised
    want = oracle.pq_train(rows, d, M, K, iters, 42)
    segs = []
    try:
        for r in range(world):
            lo, hi = shard_range(n, r, world)
            segs.append((vs.Segment.upload(rows[lo:hi]), lo))
        barrier = threading.Barrier(world)
        pending, lock = [], threading.Lock()

        def allreduce(buf):
            with lock:
                pending.append(buf)
            barrier.wait()
            if barrier.wait() == 0:            # one thread sums, everybody waits for it
                total = torch.stack(pending).sum(0) if buf.dtype != torch.float32 else None
                if total is None:              # fp32: add in rank order, like a ring would not -- any order is legal
                    total = pending[0].clone()
                    for b in pending[1:]:
                        total += b
                for b in pending:
                    b.copy_(total)
                torch.cuda.synchronize()
                pending.clear()
            barrier.wait()

        out, errs = [None] * world, []

        def run(r):
            try:
                torch.cuda.set_device(0)
                sh = ShardedSegment(segs[r][0], r, world)
                out[r] = sh.pq_train(n, segs[r][1], M, K, iters, 42, allreduce=allreduce if world > 1 else None, exact_order=exact)
            except Exception as e:
                errs.append(e)
                barrier.abort()

        threads = [threading.Thread(target=run, args=(r,)) for r in range(world)]
        for th in threads:
            th.start()
        for th in threads:
            th.join(timeout=120)
        assert not errs, errs
        for r in range(1, world):
            assert np.array_equal(out[r].view(np.uint32), out[0].view(np.uint32)), "ranks disagree"
        if world == 1 or exact:
            assert np.array_equal(out[0].view(np.uint32), want.view(np.uint32))
        else:
            e_got, e_ref = _quantisation_error(oracle, out[0], rows), _quantisation_error(oracle, want, rows)
            assert abs(e_got - e_ref) <= 0.02 * e_ref, (e_got, e_ref)
    finally:
        for seg, _ in segs:
            seg.free()


def test_empty_segment_and_bad_args(vs):
    seg = vs.Segment.upload(np.zeros((0, 8), np.float32))
    try:
        ids, sc = seg.bruteforce_topk(np.zeros(8, np.float32), 5)
        assert ids.size == 0
        with pytest.raises(ValueError):
            seg.bruteforce_topk(np.zeros(8, np.float32), 0)
        with pytest.raises(vs.VsError):
            seg.adc_topk(np.zeros(8, np.float32), 5)  # no PQ attached
    finally:
        seg.free()


# ---- PQ encode / ADC / re-rank ----------------------------------------------------------------------------
def _pq_fixture(oracle, n, d, M, K, seed=3):
    rows = oracle.gen_rows(seed, 0, n, d)
    sub = d // M
    # centroids sampled from the data (as PqTrainer's init does) so codes are well spread
    r = oracle.random(seed + 1)
    cent = np.zeros((M, K, sub), np.float32)
    for s in range(M):
        for c in range(K):
            cent[s, c] = rows[r.next_int(n), s * sub:(s + 1) * sub]
    return rows, cent


@pytest.mark.parametrize("n,d,M,K", [(20000, 128, 16, 256), (3000, 64, 8, 16), (2000, 768, 16, 256),
                                     (1500, 96, 32, 100), (500, 20, 4, 7), (900, 24, 3, 300)])
def test_pq_encode_matches_oracle(vs, oracle, n, d, M, K):
    rows, cent = _pq_fixture(oracle, n, d, M, K)
    want = oracle.pq_encode_batch(cent, rows, threads=8)
    got = vs.PqEncoder.encode_batch(cent, rows)
    assert np.array_equal(got, want)
    seg = vs.Segment.upload(rows)
    try:
        seg.attach_pq(cent)  # codes == NULL: encoded on the device from the resident rows
        assert np.array_equal(seg.codes(), want)
        assert np.array_equal(vs.PqEncoder.encode_batch(cent, segment=seg), want)
    finally:
        seg.free()


@pytest.mark.parametrize("tensor_cores", [2, 1, 0])
@pytest.mark.parametrize("n,d,M,K", [(20001, 128, 16, 256), (5000, 64, 8, 200), (777, 32, 4, 3), (63, 8, 1, 256),
                                     (9000, 40, 5, 33)])
def test_pq_encode_subdim8_tensor_core_and_ffma_nomination(vs, oracle, tensor_cores, n, d, M, K):
    """subDim 8 (the production shape) nominates with tcgen05 (fp16 hi/lo operand pairs, pq_tc.cu), with mma.sync
    3xTF32 or with the FFMA kernel; all decide
    near-ties in the reference arithmetic, so both are bit-identical to the oracle -- near-duplicate centroids,
    centroids trained on the data (rows AT a centroid) and ragged K / n included."""
    vs.set_option("pq_tensor_cores", tensor_cores)
    try:
        rows = oracle.gen_rows(42, 0, n, d)
        cent = oracle.pq_train(rows[:min(n, 3000)], d, M, K, 3, 42)
        if K > 8:
            cent[:, 5] = cent[:, 2]                                      # exact duplicates
            cent[:, 6] = cent[:, 2] * np.float32(1 + 2e-7)               # one ulp away: inside every band
            cent[0, 7] = np.nextafter(cent[0, 2], np.float32(10), dtype=np.float32)
        rows[10, :8] = cent[0, min(2, K - 1)]                            # a row exactly at a centroid
        rows[11] = 0.0
        want = oracle.pq_encode_batch(cent, rows, threads=8)
        assert np.array_equal(vs.PqEncoder.encode_batch(cent, rows), want)
        rows[20, 3] = np.nan
        rows[21] = 3e30
        rows[22] = -1e19
        cent[0, 0, 1] = np.nan
        want = oracle.pq_encode_batch(cent, rows, threads=8)
        assert np.array_equal(vs.PqEncoder.encode_batch(cent, rows), want)
    finally:
        vs.set_option("pq_tensor_cores", 2)


def test_pq_encode_duplicate_centroids_and_nan(vs, oracle):
    rows, cent = _pq_fixture(oracle, 4000, 64, 8, 64)
    cent[:, 40] = cent[:, 3]      # duplicates: strict '<' keeps the lower index
    cent[:, 41] = cent[:, 3]
    cent[2, 5, 1] = np.nan        # NaN never wins
    rows[17, 9] = np.nan          # all distances NaN in that subspace -> code 0
    rows[18] = 1e30               # fp32 estimate overflows -> all-exact path
    want = oracle.pq_encode_batch(cent, rows, threads=8)
    assert np.array_equal(vs.PqEncoder.encode_batch(cent, rows), want)


@pytest.mark.parametrize("n,d,M,K,n_cand", [(30000, 128, 16, 256, 100), (5000, 64, 8, 16, 37),
                                            (4000, 96, 32, 100, 1000), (600, 20, 5, 7, 50), (40, 128, 16, 256, 100)])
def test_adc_scan_matches_oracle(vs, oracle, n, d, M, K, n_cand):
    rows, cent = _pq_fixture(oracle, n, d, M, K)
    codes = oracle.pq_encode_batch(cent, rows, threads=8)
    q = oracle.gen_floats(77, 0, d)
    lut = oracle.build_lut(cent, q)
    assert _same(vs.build_lut(cent, q), lut)
    oi, oa = oracle.adc_topn(lut, codes, n_cand, threads=4)
    assert _same(vs.pq_approx_distance(lut, codes[:257]), np.array([oracle.pq_approx_distance(lut, c) for c in codes[:257]]))
    seg = vs.Segment.upload(rows, id_base=5)
    try:
        seg.attach_pq(cent, codes)
        ids, ap = seg.adc_topk(q, n_cand)
        assert np.array_equal(ids, oi + 5) and _same(ap, oa)
        # fused ADC -> exact re-rank (C4 shape: top-100 -> top-10)
        k = min(10, n_cand)
        for metric in (0, 1):
            ri, rs, _ = oracle.rerank_topk(rows, q, oi, k, metric)
            gi, gs = seg.adc_rerank_topk(q, n_cand, k, metric)
            assert np.array_equal(gi, ri + 5) and _same(gs, rs)
    finally:
        seg.free()


def test_adc_duplicate_codes_and_large_code_values(vs, oracle):
    rows, cent = _pq_fixture(oracle, 6000, 64, 8, 50)
    codes = oracle.pq_encode_batch(cent, rows, threads=8)
    codes[1000:1500] = codes[3]       # massive ties -> lowest rows first
    codes[20, 2] = 200                # >= K: skipped by pqApproxDistance (:1061)
    codes[21] = 255
    q = rows[3].copy()
    lut = oracle.build_lut(cent, q)
    oi, oa = oracle.adc_topn(lut, codes, 300, threads=3)
    seg = vs.Segment.upload(rows)
    try:
        seg.attach_pq(cent, codes)
        ids, ap = seg.adc_topk(q, 300)
        assert np.array_equal(ids, oi) and _same(ap, oa)
    finally:
        seg.free()


@pytest.fixture
def force_adc_fast(vs):
    """Route every M in {8, 16}, K <= 256 segment through the byte-LUT fast scan regardless of size."""
    vs.set_option("adc_fast_min_rows", 0)
    yield
    vs.set_option("adc_fast_min_rows", 16384)
    vs.set_option("adc_fast_cap", 4096)


@pytest.mark.parametrize("n,d,M,K,n_cand", [(70000, 128, 16, 256, 100), (9000, 64, 8, 16, 37), (5000, 128, 16, 256, 10),
                                            (3000, 32, 8, 256, 1000), (40, 128, 16, 256, 100), (1, 64, 8, 5, 3)])
def test_adc_fast_scan_matches_oracle(vs, oracle, force_adc_fast, n, d, M, K, n_cand):
    rows, cent = _pq_fixture(oracle, n, d, M, K)
    codes = oracle.pq_encode_batch(cent, rows, threads=8)
    seg = vs.Segment.upload(rows, id_base=11)
    try:
        seg.attach_pq(cent, codes)
        qs = np.stack([oracle.gen_floats(78 + i, 0, d) for i in range(3)])
        for q in qs:
            oi, oa = oracle.adc_topn(oracle.build_lut(cent, q), codes, n_cand, threads=4)
            ids, ap = seg.adc_topk(q, n_cand)
            assert np.array_equal(ids, oi + 11) and _same(ap, oa)
        # a query batch through the same launch (one histogram / candidate list per query)
        bi, ba, bc = seg.adc_topk(qs, n_cand)
        for j, q in enumerate(qs):
            oi, oa = oracle.adc_topn(oracle.build_lut(cent, q), codes, n_cand, threads=4)
            assert bc[j] == len(oi) and np.array_equal(bi[j, :bc[j]], oi + 11) and _same(ba[j, :bc[j]], oa)
    finally:
        seg.free()


def test_adc_fast_scan_fallbacks(vs, oracle, force_adc_fast):
    rows, cent = _pq_fixture(oracle, 20000, 64, 8, 40)
    codes = oracle.pq_encode_batch(cent, rows, threads=8)
    codes[1000:9000] = codes[3]       # 8000 exact ties inside the top: lowest rows first
    codes[20, 2] = 200                # >= K: contributes 0 (:1061)
    codes[21] = 255
    q = rows[3].copy()
    seg = vs.Segment.upload(rows)
    try:
        seg.attach_pq(cent, codes)
        oi, oa = oracle.adc_topn(oracle.build_lut(cent, q), codes, 300, threads=3)
        ids, ap = seg.adc_topk(q, 300)
        assert np.array_equal(ids, oi) and _same(ap, oa)
        # candidate list far too small: the query is evaluated exactly over every row instead
        vs.set_option("adc_fast_cap", 8)
        ids, ap = seg.adc_topk(q, 300)
        assert np.array_equal(ids, oi) and _same(ap, oa)
        ids, ap = seg.adc_topk(q, 300)    # the scratch was left clean by the fallback
        assert np.array_equal(ids, oi) and _same(ap, oa)
        vs.set_option("adc_fast_cap", 4096)
        # NaN query: every LUT entry is NaN, every distance NaN, ties -> rows 0..k-1
        qn = q.copy()
        qn[:] = np.nan
        oi, oa = oracle.adc_topn(oracle.build_lut(cent, qn), codes, 20, threads=3)
        ids, ap = seg.adc_topk(qn, 20)
        assert np.array_equal(ids, oi) and _same(ap, oa)
        # constant table (all centroids equal): zero range, all rows tie
        flat = np.zeros_like(cent)
        seg.attach_pq(flat, codes)
        oi, oa = oracle.adc_topn(oracle.build_lut(flat, q), codes, 50, threads=3)
        ids, ap = seg.adc_topk(q, 50)
        assert np.array_equal(ids, oi) and _same(ap, oa)
        # and a healthy query right after the degenerate ones
        seg.attach_pq(cent, codes)
        oi, oa = oracle.adc_topn(oracle.build_lut(cent, q), codes, 300, threads=3)
        ids, ap = seg.adc_topk(q, 300)
        assert np.array_equal(ids, oi) and _same(ap, oa)
    finally:
        seg.free()


@pytest.mark.parametrize("metric", [0, 1])
def test_rerank_matches_oracle(vs, oracle, metric):
    rows = oracle.gen_rows(21, 0, 5000, 128)
    rows[10] = rows[11]
    q = oracle.gen_floats(22, 0, 128)
    r = oracle.random(5)
    cand = np.array([r.next_int(5000) for _ in range(300)] + [10, 11, 11, 10, -1, 5000, 99999], np.int64)
    skip = np.zeros(5000, np.uint8)
    skip[cand[5]] = 1
    seg = vs.Segment.upload(rows, skip=skip)
    try:
        for k in (1, 10, 400):
            oi, os_, _ = oracle.rerank_topk(rows, q, cand, k, metric, skip=skip)
            gi, gs = seg.rerank_topk(q, cand, k, metric, normalize_on_read=bool(k % 2))
            assert np.array_equal(gi, oi) and _same(gs, os_)
    finally:
        seg.free()


def test_merge_matches_oracle(vs, oracle):
    r = np.random.default_rng(0)
    scores = -np.abs(r.standard_normal(3000)).round(2)  # many ties
    scores[17] = np.nan
    ids = r.integers(0, 1 << 40, 3000)
    for k in (1, 10, 257):
        oi, os_ = oracle.merge_topk(ids, scores, k)
        gi, gs = vs.merge_topk(ids, scores, k)
        assert np.array_equal(gi, oi) and _same(gs, os_)


# ---- PQ training -------------------------------------------------------------------------------------------
@pytest.mark.parametrize("n,d,M,K,iters", [(6000, 32, 4, 16, 5), (3000, 128, 16, 256, 2), (300, 16, 2, 64, 5),
                                           (50, 8, 2, 40, 3), (2000, 24, 3, 10, 0)])
def test_pq_train_bit_exact(vs, oracle, n, d, M, K, iters):
    rows = oracle.gen_rows(31, 0, n, d)
    want = oracle.pq_train(rows, d, M, K, iters, 42)
    got = vs.PqTrainer.train(rows, d, M, K, iters, 42)
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))
    seg = vs.Segment.upload(rows)
    try:
        got2 = vs.PqTrainer.train(None, d, M, K, iters, 42, segment=seg)
        assert np.array_equal(got2.view(np.uint32), want.view(np.uint32))
    finally:
        seg.free()


def test_pq_train_with_duplicate_rows(vs, oracle):
    rows = oracle.gen_rows(32, 0, 400, 16)
    rows[100:300] = rows[5]  # distinct indices, identical sub-vectors: unpredicted empty clusters -> replay waves
    want = oracle.pq_train(rows, 16, 4, 32, 5, 42)
    got = vs.PqTrainer.train(rows, 16, 4, 32, 5, 42)
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))


# ---- full-size configurations (BASELINE.json configs) ----------------------------------------------------
def test_c2_full_size_l2_top10(vs, oracle):
    """C2: exact L2 top-10 over 1M x 128, rows generated on the device from the Java LCG stream."""
    n, d = 1_000_000, 128
    seg = vs.Segment.generate(42, 0, n, d)
    try:
        rows = oracle.gen_rows(42, 0, n, d)
        assert np.array_equal(seg.rows(n - 1000, 1000), rows[-1000:])
        for qi in range(3):
            q = oracle.gen_floats(43, qi * d, d)
            ids, sc = seg.bruteforce_topk(q, 10)
            oi, os_, _ = oracle.bruteforce_topk(rows, q, 10, threads=8)
            assert np.array_equal(ids, oi) and _same(sc, os_)
        # size-independent properties: idempotence, sortedness, a row queries itself to distance 0
        ids2, sc2 = seg.bruteforce_topk(rows[123456], 10)
        assert ids2[0] == 123456 and sc2[0] == 0.0 and np.all(np.diff(sc2) <= 0)
    finally:
        seg.free()


@pytest.mark.gpu
def test_scratch_slots_survive_destroyed_communicator_streams(vs, oracle):
    """A host thread keeps one scratch set per stream it has worked on (four at most).  Communicators own a stream
    each; creating, using and destroying more of them than there are slots must not leave a slot keyed by a dead
    stream behind (the recycling path synchronises the slot's stream)."""
    import ctypes as C

    from vectorsearch_b200 import _lib as L

    lib = vs.load()
    n, d, k, nq = 20011, 32, 5, 3
    rows = oracle.gen_rows(42, 0, n, d)
    qs = oracle.gen_rows(43, 0, nq, d)
    seg = vs.Segment.upload(rows)
    try:
        want = [oracle.bruteforce_topk(rows, qs[i], k)[0] for i in range(nq)]
        for rep in range(7):
            comm, hd = C.c_uint64(0), (C.c_uint8 * 64)()
            L.check(lib.vs_peer_create(0, 1, 1 << 16, 4, C.byref(comm), hd))
            try:
                bases = (C.c_uint64 * 1)()
                b = C.c_uint64(0)
                L.check(lib.vs_peer_base(comm.value, C.byref(b)))
                bases[0] = b.value
                L.check(lib.vs_peer_connect_ptrs(comm.value, bases))
                ids = np.zeros((nq, k), np.int64)
                sc = np.zeros((nq, k), np.float64)
                cn = np.zeros(nq, np.int32)
                L.check(lib.vs_bruteforce_topk_exchange(seg.handle, comm.value, qs.ctypes.data_as(L.f32p), nq, k, 0,
                                                        ids.ctypes.data_as(L.i64p), sc.ctypes.data_as(L.f64p),
                                                        cn.ctypes.data_as(L.i32p)))
                for i in range(nq):
                    assert np.array_equal(ids[i], want[i])
            finally:
                lib.vs_peer_destroy(comm.value)
            got = seg.bruteforce_topk(qs, k, 0)   # the thread's own stream in between
            for i in range(nq):
                assert np.array_equal(np.asarray(got[0])[i], want[i])
    finally:
        seg.free()


# ---- one or two queries: nomination on the fp16 copy (scan_half_kernel, batch.cu) ----------------------------------------
@pytest.fixture(params=[0, 1], ids=["ctas-auto", "one-cta-per-sm"])
def force_half_scan(vs, request):
    """Route single queries (and pairs) of every segment size through the fp16-copy scan, with the automatic CTA shape
    (two 256-thread CTAs per SM where k <= 16 and the rings fit) and with one CTA per SM."""
    vs.set_option("batch_min_rows", 1)
    vs.set_option("scan_fp16", 1)
    vs.set_option("scan_half_ctas", request.param)
    yield
    vs.set_option("scan_half_ctas", 0)
    vs.set_option("batch_min_rows", 16384)


def _launches(vs):
    return vs.kernel_launch_count()


@pytest.mark.gpu
@pytest.mark.parametrize("metric", [0, 1])
@pytest.mark.parametrize("n,d,k,nq", [(40000, 128, 10, 1), (20000, 768, 32, 2), (9000, 72, 10, 1), (6000, 100, 16, 2),
                                      (300, 32, 1, 1), (7, 128, 10, 2), (70001, 64, 17, 1), (33003, 128, 5, 2),
                                      (50000, 256, 10, 1), (2500, 128, 50, 1)])
def test_half_scan_matches_oracle(vs, oracle, force_half_scan, n, d, k, nq, metric):
    rows = oracle.gen_rows(42, 0, n, d)
    qs = oracle.gen_rows(43, 0, nq, d)
    _check_batch(vs, oracle, rows, qs, k, metric)
    _check_batch(vs, oracle, rows, qs[:1].ravel().reshape(1, d), k, metric, id_base=123_456_789_012)


@pytest.mark.gpu
@pytest.mark.parametrize("metric", [0, 1])
def test_half_scan_ties_clusters_skip_and_fallbacks(vs, oracle, force_half_scan, metric):
    n, d = 60000, 64
    rows = oracle.gen_rows(11, 0, n, d)
    rows[100:400] = rows[7]           # 300 exact duplicates side by side: full lists inside the band -> exact fallback
    rows[20000] = rows[7]
    rows[9] = 0.0                     # zero norm: cosine 0.0
    rows[5000:5004] *= 3.0
    qs = oracle.gen_rows(12, 0, 6, d)
    qs[0] = rows[7]
    qs[1] = 0.0                       # zero query
    skip = np.zeros(n, np.uint8)
    skip[[7, 101, 102, n - 1]] = 1
    for i in range(0, 6, 2):          # two queries per call, then one
        _check_batch(vs, oracle, rows, qs[i:i + 2], 10, metric)
        _check_batch(vs, oracle, rows, qs[i:i + 1], 30, metric, skip=skip, id_base=1_000_000_007)
    # a cluster of near neighbours stored contiguously (the rows closest to the query sit in ONE tile)
    rows2 = rows.copy()
    rows2[30000:30040] = qs[2] + 1e-3 * oracle.gen_rows(13, 0, 40, d)
    _check_batch(vs, oracle, rows2, qs[2:3], 10, metric)
    _check_batch(vs, oracle, rows2, qs[2:3], 32, metric)
    # non-finite query -> exact scan; NaN / inf row -> the whole segment stays on the fp32 scan
    qs2 = qs.copy()
    qs2[3, 5] = np.nan
    qs2[4, 0] = np.inf
    _check_batch(vs, oracle, rows, qs2[3:5], 10, metric)
    rows3 = rows.copy()
    rows3[50, 3] = np.nan
    rows3[3000, 0] = np.inf
    _check_batch(vs, oracle, rows3, qs[:2], 10, metric)


@pytest.mark.gpu
def test_half_scan_is_the_path_taken_and_can_be_switched_off(vs, oracle):
    """1M x 128, one query: by default the fp16-copy scan answers (query conversion + scan + fallback check = 3
    launches); with scan_fp16 = 0 the fp32 streaming scan does (1 launch).  Same bits either way."""
    seg = vs.Segment.generate(42, 0, 200_000, 128)
    try:
        q = oracle.gen_rows(43, 0, 3, 128)
        seg.bruteforce_topk(q[0], 10)          # builds the copy
        a0 = _launches(vs)
        r_half = seg.bruteforce_topk(q[1], 10)
        a1 = _launches(vs)
        vs.set_option("scan_fp16", 0)
        r_full = seg.bruteforce_topk(q[1], 10)
        a2 = _launches(vs)
        assert a1 - a0 == 2 and a2 - a1 == 1
        assert np.array_equal(np.asarray(r_half[0]), np.asarray(r_full[0]))
        assert np.array_equal(np.asarray(r_half[1]).view(np.uint64), np.asarray(r_full[1]).view(np.uint64))
        rows = oracle.gen_rows(42, 0, 200_000, 128)
        oi, os_, _ = oracle.bruteforce_topk(rows, q[1], 10, 0, threads=8)
        assert np.array_equal(np.asarray(r_half[0]).ravel(), oi) and _same(np.asarray(r_half[1]).ravel(), os_)
    finally:
        vs.set_option("scan_fp16", 1)
        seg.free()
