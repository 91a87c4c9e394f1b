"""Committed golden vectors (tests/golden/golden_v1.npz, made by tests/golden/make_golden.py).

CPU part: the oracle still produces them bit for bit.  GPU part: the CUDA path, called through the
C ABI, produces them bit for bit -- without the oracle in the loop.
"""
import importlib.util
from pathlib import Path

import numpy as np
import pytest

GOLD = Path(__file__).resolve().parent / "golden" / "golden_v1.npz"
_spec = importlib.util.spec_from_file_location("make_golden", GOLD.parent / "make_golden.py")
mg = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(mg)


def _bits(a):
    a = np.asarray(a)
    return a.view(np.uint64) if a.dtype == np.float64 else (a.view(np.uint32) if a.dtype == np.float32 else a)


@pytest.fixture(scope="module")
def gold():
    return dict(np.load(GOLD))


def test_oracle_reproduces_golden(oracle, gold):
    fresh = mg.cases(oracle)
    assert sorted(fresh) == sorted(gold)
    for k in gold:
        assert np.array_equal(_bits(fresh[k]), _bits(gold[k])), k


@pytest.fixture(scope="module")
def vs():
    import vectorsearch_b200 as v

    v.init(0)
    yield v
    v.set_simd_lanes(16)


@pytest.mark.gpu
def test_gpu_distances_match_golden(vs, oracle, gold):
    for lanes in mg.LANES:
        vs.set_simd_lanes(lanes)
        want = gold[f"dist_l{lanes}"]
        for i, dim in enumerate(mg.DIMS):
            ab = oracle.gen_floats(42 + dim, 0, 2 * dim)  # input generation only (java.util.Random stream)
            a, b = ab[:dim], ab[dim:]
            got = np.array([vs.Distances.l2_squared(a, b), vs.Distances.l2(a, b), vs.Distances.dot(a, b),
                            vs.Distances.norm(a), vs.Distances.cosine(a, b)])
            assert np.array_equal(_bits(got), _bits(want[i])), (lanes, dim)
    vs.set_simd_lanes(16)


@pytest.mark.gpu
def test_gpu_bruteforce_matches_golden(vs, oracle, gold):
    inp = mg.inputs(oracle)
    seg = vs.Segment.upload(inp["bf_rows"], skip=inp["bf_skip"])
    try:
        for metric, name in ((0, "l2"), (1, "cos")):
            ids, sc = seg.bruteforce_topk(inp["bf_q"], mg.BF["k"], metric)
            assert np.array_equal(ids, gold[f"bf_{name}_ids"])
            assert np.array_equal(_bits(sc), _bits(gold[f"bf_{name}_scores"]))
    finally:
        seg.free()


@pytest.mark.gpu
@pytest.mark.parametrize("operands", ["fp16", "tf32"])
def test_gpu_batched_bruteforce_matches_golden(vs, oracle, gold, operands):
    """The tensor-core batched path (batch.cu) must return the golden lists for every query of a batch --
    here the golden query repeated next to unrelated ones -- without the oracle in the loop."""
    inp = mg.inputs(oracle)
    vs.set_option("batch_min_queries", 2)
    vs.set_option("batch_min_rows", 1)
    vs.set_option("batch_fp16", 1 if operands == "fp16" else 0)
    try:
        seg = vs.Segment.upload(inp["bf_rows"], skip=inp["bf_skip"])
        try:
            q = np.asarray(inp["bf_q"], np.float32)
            qs = np.stack([q, -q, q, q * np.float32(0.5) + np.float32(0.1), q])
            for metric, name in ((0, "l2"), (1, "cos")):
                ids, sc, cn = seg.bruteforce_topk(qs, mg.BF["k"], metric)
                for i in (0, 2, 4):
                    c = len(gold[f"bf_{name}_ids"])
                    assert cn[i] == c and np.array_equal(ids[i, :c], gold[f"bf_{name}_ids"])
                    assert np.array_equal(_bits(sc[i, :c]), _bits(gold[f"bf_{name}_scores"]))
        finally:
            seg.free()
    finally:
        vs.set_option("batch_min_queries", 3)
        vs.set_option("batch_min_rows", 16384)
        vs.set_option("batch_fp16", 1)


@pytest.mark.gpu
def test_gpu_pq_path_matches_golden(vs, oracle, gold):
    inp = mg.inputs(oracle)
    P = mg.PQ
    cent = vs.PqTrainer.train(inp["pq_rows"], P["d"], P["M"], P["K"], P["iters"], P["tseed"])
    assert np.array_equal(_bits(cent), _bits(gold["pq_centroids"]))
    assert np.array_equal(vs.PqEncoder.encode_batch(cent, inp["pq_rows"]), gold["pq_codes"])
    assert np.array_equal(_bits(vs.build_lut(cent, inp["pq_q"])), _bits(gold["pq_lut"]))
    seg = vs.Segment.upload(inp["pq_rows"])
    try:
        seg.attach_pq(cent)
        assert np.array_equal(seg.codes(), gold["pq_codes"])
        for force in (16384, 0):  # generic kernel, then the byte-LUT fast scan
            vs.set_option("adc_fast_min_rows", force)
            ai, ad = seg.adc_topk(inp["pq_q"], P["n_cand"])
            assert np.array_equal(ai, gold["adc_ids"]) and np.array_equal(_bits(ad), _bits(gold["adc_dist"]))
            for metric, name in ((0, "l2"), (1, "cos")):
                ri, rs = seg.rerank_topk(inp["pq_q"], ai, P["k"], metric)
                assert np.array_equal(ri, gold[f"rr_{name}_ids"]) and np.array_equal(_bits(rs), _bits(gold[f"rr_{name}_scores"]))
                fi, fs = seg.adc_rerank_topk(inp["pq_q"], P["n_cand"], P["k"], metric)
                assert np.array_equal(fi, gold[f"rr_{name}_ids"]) and np.array_equal(_bits(fs), _bits(gold[f"rr_{name}_scores"]))
        vs.set_option("adc_fast_min_rows", 16384)
    finally:
        seg.free()
