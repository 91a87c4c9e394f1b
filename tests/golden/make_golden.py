"""Generates tests/golden/golden_v1.npz (run from the repository root: python tests/golden/make_golden.py).

The reference (panghy/vectorsearch) is Java and cannot run in the build container (no JVM), so
these vectors come from the CPU oracle (oracle/vs_oracle.c, cross-checked by the pure-Python twin
oracle/pytwin.py), which itself is pinned to the reference's own known-answer tests in
tests/test_oracle_golden.py.  They freeze the oracle's answers: a later change of the oracle OR of
the CUDA path that moves any bit shows up against this file.  Inputs are not stored: every input
is a slice of new java.util.Random(seed) (nextFloat()*2-1, B/DistanceAndPqBenchmark.java:127-133),
reproducible by the oracle (gen_rows) and on the device (vs_segment_generate).
"""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
from oracle import pyoracle  # noqa: E402

DIMS = [1, 3, 7, 16, 17, 128, 768, 1000]
LANES = [16, 8, 4]
BF = dict(seed=101, n=600, d=24, qseed=102, k=7)
PQ = dict(seed=201, n=400, d=32, M=4, K=16, iters=5, tseed=42, qseed=202, n_cand=25, k=5)


def cases(orc):
    out = {}
    # Distances on the Random(42 + dim) pairs of T/util/DistancesTest.java:53-98, every lane model
    for lanes in LANES:
        orc.set_lanes(lanes)
        vals = []
        for dim in DIMS:
            ab = orc.gen_floats(42 + dim, 0, 2 * dim)
            a, b = ab[:dim], ab[dim:]
            vals.append([orc.l2_squared(a, b), orc.l2(a, b), orc.dot(a, b), orc.norm(a), orc.cosine(a, b)])
        out[f"dist_l{lanes}"] = np.array(vals, dtype=np.float64)
    orc.set_lanes(16)
    # brute force (J/fdb/FdbVectorIndex.java:676-721): L2 and cosine, skip mask, duplicated rows (ties)
    rows = orc.gen_rows(BF["seed"], 0, BF["n"], BF["d"])
    rows[300:310] = rows[5]
    skip = np.zeros(BF["n"], np.uint8)
    skip[[5, 17, 301]] = 1
    q = rows[5].copy()
    q[0] += 0.25
    for metric, name in ((0, "l2"), (1, "cos")):
        ids, sc, _ = orc.bruteforce_topk(rows, q, BF["k"], metric, skip=skip)
        out[f"bf_{name}_ids"], out[f"bf_{name}_scores"] = ids, sc
    # PQ: train (J/pq/PqTrainer.java:28-91), encode (J/pq/PqEncoder.java:18-37), LUT (:1067-1079), ADC (:754-769),
    # re-rank (:997-1043)
    prow = orc.gen_rows(PQ["seed"], 0, PQ["n"], PQ["d"])
    cent = orc.pq_train(prow, PQ["d"], PQ["M"], PQ["K"], PQ["iters"], PQ["tseed"])
    codes = orc.pq_encode_batch(cent, prow)
    pq_q = orc.gen_floats(PQ["qseed"], 0, PQ["d"])
    lut = orc.build_lut(cent, pq_q)
    ai, ad = orc.adc_topn(lut, codes, PQ["n_cand"])
    out.update(pq_centroids=cent, pq_codes=codes, pq_lut=lut, adc_ids=ai, adc_dist=ad)
    for metric, name in ((0, "l2"), (1, "cos")):
        ri, rs, _ = orc.rerank_topk(prow, pq_q, ai, PQ["k"], metric)
        out[f"rr_{name}_ids"], out[f"rr_{name}_scores"] = ri, rs
    return out


def inputs(orc):
    """The inputs the cases above are computed from (also used by the tests)."""
    rows = orc.gen_rows(BF["seed"], 0, BF["n"], BF["d"])
    rows[300:310] = rows[5]
    skip = np.zeros(BF["n"], np.uint8)
    skip[[5, 17, 301]] = 1
    q = rows[5].copy()
    q[0] += 0.25
    prow = orc.gen_rows(PQ["seed"], 0, PQ["n"], PQ["d"])
    pq_q = orc.gen_floats(PQ["qseed"], 0, PQ["d"])
    return dict(bf_rows=rows, bf_skip=skip, bf_q=q, pq_rows=prow, pq_q=pq_q)


if __name__ == "__main__":
    orc = pyoracle.get()
    data = cases(orc)
    dst = Path(__file__).resolve().parent / "golden_v1.npz"
    np.savez_compressed(dst, **data)
    print(f"wrote {dst} ({dst.stat().st_size} bytes, {len(data)} arrays)")
