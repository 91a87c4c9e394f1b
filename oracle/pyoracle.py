"""ctypes front-end of the CPU ORACLE (oracle/vs_oracle.c).

TEST INFRASTRUCTURE ONLY.  Importers allowed: tests/, __graft_entry__.smoke(), and
bench.py's cpu_baseline / --impl reference legs.  vectorsearch_b200/ never imports this.

The reference (panghy/vectorsearch) is Java and this image has no JVM, so there is no
oracle/_ref build; parity is pinned by replaying the reference's own known-answer tests
(tests/test_oracle_golden.py).  What those tests leave open is "pinned by restatement only".
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from pathlib import Path

import numpy as np

_HERE = Path(__file__).resolve().parent
METRIC_L2 = 0
METRIC_COSINE = 1

_f32p = C.POINTER(C.c_float)
_f64p = C.POINTER(C.c_double)
_u8p = C.POINTER(C.c_uint8)
_i64p = C.POINTER(C.c_int64)


def build(native: bool = False, force: bool = False) -> Path:
    """Compile the oracle with the Makefile beside it (gcc only, no reference build system)."""
    target = "libvsoracle_native.so" if native else "libvsoracle.so"
    subprocess.run(["make", "-C", str(_HERE)] + (["-B"] if force else []) + [target], check=True, capture_output=True)
    return _HERE / target


def _load(native: bool) -> C.CDLL:
    name = "libvsoracle_native.so" if native else "libvsoracle.so"
    path = _HERE / name
    src_mtime = max((_HERE / "vs_oracle.c").stat().st_mtime, (_HERE / "vs_oracle.h").stat().st_mtime)
    if not path.exists() or path.stat().st_mtime < src_mtime:
        build(native)
    lib = C.CDLL(str(path))
    lib.vso_l2_squared.restype = C.c_double
    lib.vso_l2.restype = C.c_double
    lib.vso_dot.restype = C.c_double
    lib.vso_norm.restype = C.c_double
    lib.vso_cosine.restype = C.c_double
    lib.vso_pq_approx_distance.restype = C.c_double
    lib.vso_jr_next_float.restype = C.c_float
    lib.vso_bench_ns_per_op.restype = C.c_double
    for fn in ("vso_adc_topn", "vso_bruteforce_topk", "vso_rerank_topk", "vso_merge_topk"):
        getattr(lib, fn).restype = C.c_int64
    return lib


class _JRandomStruct(C.Structure):
    _fields_ = [("state", C.c_uint64)]


def _p(a: np.ndarray, t):
    return a.ctypes.data_as(t)


def _f32(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.float32)


class JavaRandom:
    """java.util.Random (restated in vs_oracle.c)."""

    def __init__(self, seed: int, lib: C.CDLL):
        self._lib = lib
        self._s = _JRandomStruct()
        lib.vso_jr_init(C.byref(self._s), C.c_int64(seed))

    def next_int(self, bound: int | None = None) -> int:
        if bound is None:
            return int(self._lib.vso_jr_next_int(C.byref(self._s)))
        if bound <= 0:
            raise ValueError("bound must be positive")
        return int(self._lib.vso_jr_next_int_bound(C.byref(self._s), C.c_int32(bound)))

    def next_float(self) -> float:
        return float(self._lib.vso_jr_next_float(C.byref(self._s)))

    def skip(self, n: int) -> None:
        self._lib.vso_jr_skip(C.byref(self._s), C.c_uint64(n))


class Oracle:
    """One loaded copy of the oracle.  `native=True` is the -march=native build used for timing."""

    def __init__(self, native: bool = False):
        self.lib = _load(native)
        self.native = native

    # -- configuration -----------------------------------------------------------------
    def set_lanes(self, lanes: int) -> None:
        self.lib.vso_set_lanes(C.c_int(lanes))

    def get_lanes(self) -> int:
        return int(self.lib.vso_get_lanes())

    def random(self, seed: int) -> JavaRandom:
        return JavaRandom(seed, self.lib)

    # -- Distances ---------------------------------------------------------------------
    def l2_squared(self, a, b) -> float:
        a, b = _f32(a), _f32(b)
        return float(self.lib.vso_l2_squared(_p(a, _f32p), _p(b, _f32p), C.c_int(a.size)))

    def l2(self, a, b) -> float:
        a, b = _f32(a), _f32(b)
        return float(self.lib.vso_l2(_p(a, _f32p), _p(b, _f32p), C.c_int(a.size)))

    def dot(self, a, b) -> float:
        a, b = _f32(a), _f32(b)
        return float(self.lib.vso_dot(_p(a, _f32p), _p(b, _f32p), C.c_int(a.size)))

    def norm(self, a) -> float:
        a = _f32(a)
        return float(self.lib.vso_norm(_p(a, _f32p), C.c_int(a.size)))

    def cosine(self, a, b) -> float:
        a, b = _f32(a), _f32(b)
        return float(self.lib.vso_cosine(_p(a, _f32p), _p(b, _f32p), C.c_int(a.size)))

    # -- PQ ----------------------------------------------------------------------------
    def pq_encode(self, centroids, v) -> np.ndarray:
        c = _f32(centroids)
        M, K, sub = c.shape
        v = _f32(v)
        out = np.zeros(M, dtype=np.uint8)
        self.lib.vso_pq_encode(_p(c, _f32p), M, K, sub, _p(v, _f32p), _p(out, _u8p))
        return out

    def pq_encode_batch(self, centroids, rows, threads: int = 1) -> np.ndarray:
        c = _f32(centroids)
        M, K, sub = c.shape
        rows = _f32(rows)
        n = rows.shape[0]
        out = np.zeros((n, M), dtype=np.uint8)
        self.lib.vso_pq_encode_batch(_p(c, _f32p), M, K, sub, _p(rows, _f32p), C.c_int64(n),
                                     _p(out, _u8p), C.c_int(threads))
        return out

    def pq_encode_batch_fast(self, centroids, rows, threads: int = 1) -> np.ndarray:
        """pq_encode_batch with the centroid-major evaluation (bit-identical; for BASELINE-size checks)."""
        c = _f32(centroids)
        M, K, sub = c.shape
        rows = _f32(rows)
        n = rows.shape[0]
        out = np.zeros((n, M), dtype=np.uint8)
        self.lib.vso_pq_encode_batch_fast(_p(c, _f32p), M, K, sub, _p(rows, _f32p), C.c_int64(n),
                                          _p(out, _u8p), C.c_int(threads))
        return out

    def pq_train(self, rows, D: int, M: int, K: int, iterations: int, seed: int, return_draws=False,
                 threads: int = 0):
        """threads == 0: the literal single-threaded restatement; > 0: vso_pq_train_mt (bit-identical)."""
        rows = _f32(rows).reshape(-1, D) if D > 0 else _f32(rows)
        n = rows.shape[0]
        if M <= 0 or K <= 0 or D <= 0 or D % M != 0:
            raise ValueError("Invalid PQ params (m,k,dimension)")  # IllegalArgumentException
        out = np.zeros((M, K, D // M), dtype=np.float32)
        draws = C.c_int64(0)
        if threads > 0:
            rc = self.lib.vso_pq_train_mt(_p(rows, _f32p), C.c_int64(n), D, M, K, iterations,
                                          C.c_int64(seed), _p(out, _f32p), C.byref(draws), C.c_int(threads))
        else:
            rc = self.lib.vso_pq_train(_p(rows, _f32p), C.c_int64(n), D, M, K, iterations,
                                       C.c_int64(seed), _p(out, _f32p), C.byref(draws))
        if rc == -1:
            raise ValueError("Invalid PQ params (m,k,dimension)")
        if rc == -2:
            raise IndexError("empty training set")
        return (out, int(draws.value)) if return_draws else out

    # -- ADC ---------------------------------------------------------------------------
    def build_lut(self, centroids, q) -> np.ndarray:
        c = _f32(centroids)
        M, K, sub = c.shape
        q = _f32(q)
        lut = np.zeros((M, K), dtype=np.float64)
        self.lib.vso_build_lut(_p(c, _f32p), M, K, sub, _p(q, _f32p), _p(lut, _f64p))
        return lut

    def pq_approx_distance(self, lut, codes, K: int | None = None) -> float:
        lut = np.ascontiguousarray(lut, dtype=np.float64)
        codes = np.ascontiguousarray(codes, dtype=np.uint8)
        M, KK = lut.shape
        return float(self.lib.vso_pq_approx_distance(_p(lut, _f64p), _p(codes, _u8p), M,
                                                     KK if K is None else K))

    def adc_topn(self, lut, codes, n_cand: int, threads: int = 1):
        lut = np.ascontiguousarray(lut, dtype=np.float64)
        codes = np.ascontiguousarray(codes, dtype=np.uint8)
        M, K = lut.shape
        n = codes.shape[0]
        cap = max(0, min(n_cand, n))
        ids = np.zeros(max(cap, 1), dtype=np.int64)
        ap = np.zeros(max(cap, 1), dtype=np.float64)
        c = self.lib.vso_adc_topn(_p(lut, _f64p), M, K, _p(codes, _u8p), C.c_int64(n),
                                  C.c_int64(cap), _p(ids, _i64p), _p(ap, _f64p), C.c_int(threads))
        return ids[:c].copy(), ap[:c].copy()

    # -- exact scorers -----------------------------------------------------------------
    def bruteforce_topk(self, rows, q, k: int, metric: int = METRIC_L2, skip=None, threads: int = 1):
        rows = _f32(rows)
        n, d = rows.shape
        q = _f32(q)
        cap = max(0, min(k, n))
        ids = np.zeros(max(cap, 1), dtype=np.int64)
        sc = np.zeros(max(cap, 1), dtype=np.float64)
        di = np.zeros(max(cap, 1), dtype=np.float64)
        sp = None
        if skip is not None:
            skip = np.ascontiguousarray(skip, dtype=np.uint8)
            sp = _p(skip, _u8p)
        c = self.lib.vso_bruteforce_topk(_p(rows, _f32p), C.c_int64(n), d, sp, _p(q, _f32p), metric,
                                         C.c_int64(cap), _p(ids, _i64p), _p(sc, _f64p),
                                         _p(di, _f64p), C.c_int(threads))
        return ids[:c].copy(), sc[:c].copy(), di[:c].copy()

    def rerank_topk(self, rows, q, cand, k: int, metric: int = METRIC_L2, skip=None,
                    normalize_on_read: bool = False):
        rows = _f32(rows)
        n, d = rows.shape
        q = _f32(q)
        cand = np.ascontiguousarray(cand, dtype=np.int64)
        cap = max(0, min(k, cand.size))
        ids = np.zeros(max(cap, 1), dtype=np.int64)
        sc = np.zeros(max(cap, 1), dtype=np.float64)
        di = np.zeros(max(cap, 1), dtype=np.float64)
        sp = None
        if skip is not None:
            skip = np.ascontiguousarray(skip, dtype=np.uint8)
            sp = _p(skip, _u8p)
        c = self.lib.vso_rerank_topk(_p(rows, _f32p), C.c_int64(n), d, sp, _p(q, _f32p), metric,
                                     int(bool(normalize_on_read)), _p(cand, _i64p),
                                     C.c_int64(cand.size), C.c_int64(cap), _p(ids, _i64p),
                                     _p(sc, _f64p), _p(di, _f64p))
        return ids[:c].copy(), sc[:c].copy(), di[:c].copy()

    def merge_topk(self, ids, scores, k: int):
        ids = np.ascontiguousarray(ids, dtype=np.int64)
        scores = np.ascontiguousarray(scores, dtype=np.float64)
        cap = max(0, min(k, ids.size))
        oi = np.zeros(max(cap, 1), dtype=np.int64)
        os_ = np.zeros(max(cap, 1), dtype=np.float64)
        c = self.lib.vso_merge_topk(_p(ids, _i64p), _p(scores, _f64p), C.c_int64(ids.size),
                                    C.c_int64(cap), _p(oi, _i64p), _p(os_, _f64p))
        return oi[:c].copy(), os_[:c].copy()

    # -- GraphBuilder ------------------------------------------------------------------
    def knn_graph(self, rows, degree: int, l_build: int = 0, alpha: float = 1.0, threads: int = 1):
        """buildL2Neighbors (l_build == 0) / buildPrunedNeighbors -> list of int32 arrays."""
        rows = _f32(rows)
        n, d = rows.shape
        out = np.full((n, max(degree, 1)), -1, dtype=np.int32)
        cn = np.zeros(n, dtype=np.int32)
        self.lib.vso_knn_graph(_p(rows, _f32p), C.c_int64(n), d, degree, l_build, C.c_double(alpha),
                               out.ctypes.data_as(C.POINTER(C.c_int32)), cn.ctypes.data_as(C.POINTER(C.c_int32)), C.c_int(threads))
        return [out[i, :cn[i]] for i in range(n)]

    # -- JMH-like timing (config C1) ------------------------------------------------------
    def bench_ns_per_op(self, kind: int, a=None, b=None, centroids=None, lut=None, codes=None, iters: int = 100000) -> float:
        a = _f32(a) if a is not None else np.zeros(1, np.float32)
        b = _f32(b) if b is not None else a
        c = _f32(centroids) if centroids is not None else np.zeros((1, 1, 1), np.float32)
        M, K, sub = c.shape
        lt = _f32(lut) if lut is not None else np.zeros((M, K), np.float32)
        cd = np.ascontiguousarray(codes, dtype=np.uint8) if codes is not None else np.zeros(M, np.uint8)
        return float(self.lib.vso_bench_ns_per_op(kind, _p(a, _f32p), _p(b, _f32p), C.c_int(a.size), _p(c, _f32p), M, K, sub,
                                                  _p(lt, _f32p), _p(cd, _u8p), C.c_int64(iters)))

    # -- FloatPacker -------------------------------------------------------------------
    def floats_to_bytes(self, arr) -> bytes:
        a = _f32(arr)
        out = np.zeros(a.size * 4, dtype=np.uint8)
        self.lib.vso_floats_to_bytes(_p(a, _f32p), C.c_int(a.size), _p(out, _u8p))
        return out.tobytes()

    def bytes_to_floats(self, b: bytes) -> np.ndarray:
        src = np.frombuffer(b, dtype=np.uint8).copy()
        out = np.zeros(len(b) // 4, dtype=np.float32)
        self.lib.vso_bytes_to_floats(_p(src, _u8p), C.c_int(len(b)), _p(out, _f32p))
        return out

    # -- synthetic inputs --------------------------------------------------------------
    def gen_floats(self, seed: int, first: int, count: int, kind: int = 0) -> np.ndarray:
        out = np.zeros(count, dtype=np.float32)
        self.lib.vso_gen_floats(C.c_int64(seed), C.c_int64(first), C.c_int64(count), C.c_int(kind),
                                _p(out, _f32p))
        return out

    def gen_rows(self, seed: int, first_row: int, n: int, d: int, kind: int = 0) -> np.ndarray:
        return self.gen_floats(seed, first_row * d, n * d, kind).reshape(n, d)

    def gen_codes(self, seed: int, first: int, count: int) -> np.ndarray:
        out = np.zeros(count, dtype=np.uint8)
        self.lib.vso_gen_codes(C.c_int64(seed), C.c_int64(first), C.c_int64(count), _p(out, _u8p))
        return out


_default: Oracle | None = None


def get(native: bool = False) -> Oracle:
    global _default
    if native:
        return Oracle(native=True)
    if _default is None:
        _default = Oracle()
    return _default


def host_threads() -> int:
    return len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
