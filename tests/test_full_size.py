"""GPU parity at BASELINE.json's FULL sizes (configs C3, C4, C5; C2 lives in test_gpu_parity.py).

The size-dependent code -- the 100M-row histogram threshold and per-CTA candidate lists of the ADC
fast scan, the multi-slab tcgen05 encode, the streaming-operand nomination GEMM of the 768-d
cosine batch -- is exactly what small cases cannot reach, so these tests run the real shapes on the
device and compare with the oracle (all host threads; rows regenerated from the java.util.Random
stream slab by slab, so no more than a few GB of host memory is ever held).

Bar: ids, codes, centroids, distances and scores bit-exact (tolerance 0).
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


@pytest.fixture(scope="module")
def threads():
    from oracle import pyoracle

    return max(1, pyoracle.host_threads())


def _same(a, b):
    a, b = np.asarray(a), np.asarray(b)
    if a.shape != b.shape:
        return False
    if a.dtype != np.float64:
        return np.array_equal(a, b)
    na, nb = np.isnan(a), np.isnan(b)
    return np.array_equal(na, nb) and np.array_equal(a[~na].view(np.uint64), b[~nb].view(np.uint64))


def _need_bytes(vs, nbytes):
    info = vs.device_info()
    if info["free_bytes"] < nbytes:
        pytest.skip(f"needs {nbytes / 1e9:.0f} GB of free device memory, {info['free_bytes'] / 1e9:.0f} GB available")


def _oracle_bruteforce_slabs(oracle, seed, n, d, q, k, metric, threads, slab=1_000_000):
    """searchBruteForceSegment over rows [0, n) of the LCG stream, regenerated slab by slab: per-slab top-k in
    ascending row order merged with the reference's stable merge (J/fdb/FdbVectorIndex.java:432-437) -- equal
    to one pass over all rows because ties keep the lower slab = the lower row."""
    ids, scs = [], []
    for r0 in range(0, n, slab):
        cnt = min(slab, n - r0)
        rows = oracle.gen_rows(seed, r0, cnt, d)
        i, s, _ = oracle.bruteforce_topk(rows, q, k, metric, threads=threads)
        ids.append(i + r0)
        scs.append(s)
    return oracle.merge_topk(np.concatenate(ids), np.concatenate(scs), k)


# ---- C3: PQ k-means train + encode over 10M x 128, M=16, K=256, 5 iterations, seed 42 ----------------------
def test_c3_full_size_train_and_encode(vs, oracle, threads):
    """PqTrainer.train(vectors, 128, 16, 256, 5, 42) (J/pq/PqTrainer.java:28-91, production parameters of
    J/tasks/SegmentBuildService.java:180) on 10M resident rows: centroids bit-identical to the oracle's
    (order-preserving multi-threaded restatement, itself checked against the literal one on the CPU), then
    PqEncoder.encode (J/pq/PqEncoder.java:18-37) of all 10M rows on the device against the oracle on two 1M slices."""
    n, d, M, K = 10_000_000, 128, 16, 256
    _need_bytes(vs, 12 << 30)
    seg = vs.Segment.generate(42, 0, n, d)
    try:
        cent = vs.PqTrainer.train(None, d, M, K, 5, 42, segment=seg)
        rows = oracle.gen_rows(42, 0, n, d)
        want = oracle.pq_train(rows, d, M, K, 5, 42, threads=threads)
        assert np.array_equal(cent.view(np.uint32), want.view(np.uint32)), "C3 centroids differ from the oracle"
        seg.attach_pq(cent)
        for r0 in (0, n - 1_000_000):
            got = seg.codes(r0, 1_000_000)
            exp = oracle.pq_encode_batch_fast(want, rows[r0:r0 + 1_000_000], threads=threads)
            assert np.array_equal(got, exp), f"C3 codes of rows {r0}.. differ from the oracle"
        # size-independent property: a row's code is the code of the nearest centroid's own sub-vector
        codes = seg.codes(5_000_000, 4096)
        recon = np.stack([cent[s, codes[:, s]] for s in range(M)], axis=1).reshape(-1, d)
        again = vs.PqEncoder.encode_batch(cent, recon)
        assert np.array_equal(again, codes)
    finally:
        seg.free()


# ---- C4: ADC top-100 + exact re-rank to top-10 over 100M x 128 ------------------------------------------------
def test_c4_full_size_adc_rerank(vs, oracle, threads):
    """100M x 128 rows resident (51.2 GB) + 1.6 GB of codes encoded on the device with a codebook trained on the
    first 1M rows.  Checks (a) the codebook, (b) 1M-row code slices, (c) ADC top-100 ids and approximate
    distances of 3 queries against the oracle's scan over all 100M downloaded codes
    (J/fdb/FdbVectorIndex.java:754-769,820-822), (d) the re-ranked top-10 against the oracle's
    fetchExactAndScore (:997-1043) over the 100 regenerated candidate rows, (e) the fused call."""
    n, d, M, K, n_cand, k = 100_000_000, 128, 16, 256, 100, 10
    _need_bytes(vs, 60 << 30)
    train = vs.Segment.generate(42, 0, 1_000_000, d)
    try:
        cent = vs.PqTrainer.train(None, d, M, K, 5, 42, segment=train)
    finally:
        train.free()
    rows1m = oracle.gen_rows(42, 0, 1_000_000, d)
    want = oracle.pq_train(rows1m, d, M, K, 5, 42, threads=threads)
    assert np.array_equal(cent.view(np.uint32), want.view(np.uint32)), "C4 codebook differs from the oracle"
    seg = vs.Segment.generate(42, 0, n, d)
    try:
        seg.attach_pq(cent)  # 100M rows encoded on the device
        codes = seg.codes()  # 1.6 GB
        assert codes.shape == (n, M)
        exp = oracle.pq_encode_batch_fast(cent, rows1m, threads=threads)
        assert np.array_equal(codes[:1_000_000], exp), "C4 codes of the first 1M rows differ"
        r0 = 73_000_000
        exp = oracle.pq_encode_batch_fast(cent, oracle.gen_rows(42, r0, 500_000, d), threads=threads)
        assert np.array_equal(codes[r0:r0 + 500_000], exp), "C4 codes of rows 73M.. differ"
        tail = oracle.gen_rows(42, n - 100_000, 100_000, d)
        assert np.array_equal(codes[-100_000:], oracle.pq_encode_batch_fast(cent, tail, threads=threads))
        for qi in range(3):
            q = oracle.gen_floats(43, qi * d, d)
            lut = oracle.build_lut(cent, q)
            oi, oa = oracle.adc_topn(lut, codes, n_cand, threads=threads)
            gi, ga = seg.adc_topk(q, n_cand)
            assert np.array_equal(gi, oi), f"C4 ADC top-{n_cand} ids differ (query {qi})"
            assert _same(ga, oa), f"C4 ADC distances differ (query {qi})"
            # re-rank: the 100 candidate rows are regenerated from the stream; the oracle ranks them by position
            cand_rows = np.concatenate([oracle.gen_rows(42, int(i), 1, d) for i in oi])
            pos, sc, _ = oracle.rerank_topk(cand_rows, q, np.arange(n_cand), k)
            ri, rs = seg.rerank_topk(q, oi, k)
            assert np.array_equal(ri, oi[pos]) and _same(rs, sc), f"C4 re-rank differs (query {qi})"
            fi, fs = seg.adc_rerank_topk(q, n_cand, k)
            assert np.array_equal(fi, oi[pos]) and _same(fs, sc), f"C4 fused ADC + re-rank differs (query {qi})"
        # a batch of queries through one call equals the single-query calls
        qs = oracle.gen_rows(43, 3, 5, d)
        bi, bs, bc = seg.adc_rerank_topk(qs, n_cand, k)
        for j in range(5):
            si, ss = seg.adc_rerank_topk(qs[j], n_cand, k)
            assert bc[j] == k and np.array_equal(bi[j], si) and _same(bs[j], ss)
    finally:
        seg.free()


# ---- C5: cosine top-50 over a 6.25M x 768 shard, query batch 256 ------------------------------------------------
def test_c5_full_size_cosine_batch(vs, oracle, threads):
    """One GPU's row range of C5 (50M x 768 over 8 GPUs = 6.25M rows = 19.2 GB), 256 queries, cosine top-50
    (J/fdb/FdbVectorIndex.java:676-721 with Metric.COSINE): the tensor-core batch path against the oracle for 3
    queries, and against the per-query streaming kernel for 8 more."""
    n, d, nq, k = 6_250_000, 768, 256, 50
    _need_bytes(vs, 32 << 30)
    seg = vs.Segment.generate(42, 0, n, d)
    try:
        qs = oracle.gen_rows(43, 0, nq, d)
        ids, sc, cn = seg.bruteforce_topk(qs, k, metric=1)
        assert np.all(cn == k) and np.all(np.diff(sc, axis=1) <= 0)
        for i in (0, 128, 255):
            oi, os_ = _oracle_bruteforce_slabs(oracle, 42, n, d, qs[i], k, 1, threads)
            assert np.array_equal(ids[i], oi), f"C5 ids differ (query {i})"
            assert _same(sc[i], os_), f"C5 scores differ (query {i})"
        for i in range(1, nq, 37):
            si, ss = seg.bruteforce_topk(qs[i], k, metric=1)
            assert np.array_equal(ids[i], si) and _same(sc[i], ss)
    finally:
        seg.free()
