"""Host-side mirror of the reference interface, over the C ABI (no torch types cross it)."""
from __future__ import annotations

import ctypes as C
import threading

import numpy as np

from . import _lib as L
from ._lib import METRIC_COSINE, METRIC_L2, check

_initialised = False


def _f32(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.float32)


# Result buffers of the query calls, one set per (thread, shape): numpy's ctypes accessors cost several microseconds per
# array and per call -- 21 us of wrapper around an 80 us query -- so the hot calls keep their output arrays and the raw
# addresses, pass plain integers through ctypes, and hand COPIES of the filled part back.
_tls = threading.local()


def _out_buffers(nq: int, k: int):
    cache = getattr(_tls, "out", None)
    if cache is None:
        cache = _tls.out = {}
    b = cache.get((nq, k))
    if b is None:
        if len(cache) > 64:
            cache.clear()
        ids = np.zeros((nq, k), dtype=np.int64)
        sc = np.zeros((nq, k), dtype=np.float64)
        cn = np.zeros(nq, dtype=np.int32)
        b = cache[(nq, k)] = (ids, sc, cn, ids.ctypes.data, sc.ctypes.data, cn.ctypes.data)
    return b


def _query_ptr(q, d: int):
    """-> (array kept alive, address, nq, single) for a query [d] or a batch [nq][d]."""
    if type(q) is not np.ndarray or q.dtype != np.float32 or not q.flags.c_contiguous:
        q = np.ascontiguousarray(q, dtype=np.float32)
    single = q.ndim == 1
    if (q.shape[0] if single else q.shape[-1]) != d or q.ndim > 2:
        raise ValueError("query dimension does not match the segment")
    return q, q.__array_interface__["data"][0], (1 if single else q.shape[0]), single


def _results(b, nq: int, single: bool):
    ids, sc, cn = b[0], b[1], b[2]
    if single:
        c = int(cn[0])
        return ids[0, :c].copy(), sc[0, :c].copy()
    return ids.copy(), sc.copy(), cn.copy()


def _p(a: np.ndarray, t):
    return a.ctypes.data_as(t)


def init(device: int = 0) -> None:
    """Bind this process to one CUDA device (vs_init).  Raises if there is none."""
    global _initialised
    check(L.load().vs_init(device))
    _initialised = True


def init_multi(devices) -> None:
    """One process, several GPUs (vs_init_multi): segments created afterwards are sharded over `devices` and
    every Segment / PqTrainer / PqEncoder call fans out inside libvsgpu."""
    global _initialised
    devs = (C.c_int32 * len(devices))(*[int(d) for d in devices])
    check(L.load().vs_init_multi(len(devices), devs))
    _initialised = True


def device_count() -> int:
    return int(L.load().vs_device_count())


def _ensure() -> C.CDLL:
    if not _initialised:
        init(0)
    return L.load()


def shutdown() -> None:
    global _initialised
    check(L.load().vs_shutdown())
    _initialised = False


def set_simd_lanes(lanes: int) -> None:
    check(L.load().vs_set_simd_lanes(lanes))


def set_option(name: str, value: int) -> None:
    """Tuning knobs of libvsgpu (vs_set_option); results never depend on them."""
    check(L.load().vs_set_option(name.encode(), int(value)))


def kernel_launch_count() -> int:
    return int(L.load().vs_kernel_launch_count())


def device_info() -> dict:
    lib = _ensure()
    sm, fr, tot = C.c_int32(), C.c_int64(), C.c_int64()
    check(lib.vs_device_info(C.byref(sm), C.byref(fr), C.byref(tot)))
    return {"sm_count": sm.value, "free_bytes": fr.value, "total_bytes": tot.value}


class Distances:
    """J/util/Distances.java -- static methods, double results in the reference's arithmetic."""

    @staticmethod
    def _pair(fn, a, b):
        a, b = _f32(a).ravel(), _f32(b).ravel()
        if a.size != b.size:
            raise ValueError("vectors must have the same length")
        out = C.c_double()
        check(fn(_p(a, L.f32p), _p(b, L.f32p), a.size, C.byref(out)))
        return out.value

    @staticmethod
    def l2(a, b) -> float:
        return Distances._pair(_ensure().vs_l2, a, b)

    @staticmethod
    def l2_squared(a, b, a_offset: int | None = None, b_offset: int = 0, length: int | None = None) -> float:
        """l2Squared(a, b) or l2Squared(a, aOffset, b, bOffset, length) (Distances.java:48-94)."""
        if a_offset is not None:
            a = _f32(a).ravel()[a_offset:a_offset + length]
            b = _f32(b).ravel()[b_offset:b_offset + length]
        return Distances._pair(_ensure().vs_l2_squared, a, b)

    @staticmethod
    def dot(a, b) -> float:
        return Distances._pair(_ensure().vs_dot, a, b)

    @staticmethod
    def norm(a) -> float:
        a = _f32(a).ravel()
        out = C.c_double()
        check(_ensure().vs_norm(_p(a, L.f32p), a.size, C.byref(out)))
        return out.value

    @staticmethod
    def cosine(a, b) -> float:
        return Distances._pair(_ensure().vs_cosine, a, b)


class PqEncoder:
    """J/pq/PqEncoder.java"""

    @staticmethod
    def encode(centroids, v) -> np.ndarray:
        c = _f32(centroids)
        M, K, sub = c.shape
        v = _f32(v).ravel()
        out = np.zeros(M, dtype=np.uint8)
        check(_ensure().vs_pq_encode(_p(c, L.f32p), M, K, sub, _p(v, L.f32p), _p(out, L.u8p)))
        return out

    @staticmethod
    def encode_batch(centroids, rows=None, segment: "Segment | None" = None, n: int | None = None) -> np.ndarray:
        """PqEncoder.encode over many rows (SegmentBuildService.java:301), host rows or a resident segment."""
        c = _f32(centroids)
        M, K, sub = c.shape
        if rows is not None:
            rows = _f32(rows).reshape(-1, M * sub)
            n = rows.shape[0]
            rp, h = _p(rows, L.f32p), 0
        else:
            n = segment.n if n is None else n
            rp, h = None, segment.handle
        out = np.zeros((n, M), dtype=np.uint8)
        check(_ensure().vs_pq_encode_batch(_p(c, L.f32p), M, K, sub, rp, h, n, _p(out, L.u8p)))
        return out


class PqTrainer:
    """J/pq/PqTrainer.java"""

    @staticmethod
    def train(vectors, dimension: int, m: int, k: int, iterations: int, seed: int,
              segment: "Segment | None" = None) -> np.ndarray:
        lib = _ensure()
        if m <= 0 or k <= 0 or dimension <= 0:
            raise ValueError("Invalid PQ params (m,k,dimension)")
        if dimension % m != 0:
            raise ValueError("dimension must be divisible by m")
        out = np.zeros((m, k, dimension // m), dtype=np.float32)
        if segment is not None:
            check(lib.vs_pq_train(None, segment.handle, segment.n, dimension, m, k, iterations, seed,
                                  _p(out, L.f32p)))
        else:
            rows = _f32(vectors)
            rows = rows.reshape(-1, dimension) if rows.size else rows.reshape(0, dimension)
            check(lib.vs_pq_train(_p(rows, L.f32p), 0, rows.shape[0], dimension, m, k, iterations, seed,
                                  _p(out, L.f32p)))
        return out


def pq_lut_distance(lut, codes) -> float:
    """pqLutDistance of the JMH suite (float LUT, float sum), DistanceAndPqBenchmark.java:116-123."""
    lut = _f32(lut)
    M, K = lut.shape
    codes = np.ascontiguousarray(codes, dtype=np.uint8)
    out = C.c_float()
    check(_ensure().vs_pq_lut_distance(_p(lut, L.f32p), M, K, _p(codes, L.u8p), C.byref(out)))
    return out.value


def build_lut(centroids, q) -> np.ndarray:
    """FdbVectorIndex.buildLut (:1067-1079) -> double[M][K]."""
    c = _f32(centroids)
    M, K, sub = c.shape
    q = _f32(q).ravel()
    lut = np.zeros((M, K), dtype=np.float64)
    check(_ensure().vs_build_lut(_p(c, L.f32p), M, K, sub, _p(q, L.f32p), _p(lut, L.f64p)))
    return lut


def pq_approx_distance(lut, codes, k_cent: int | None = None) -> np.ndarray:
    """FdbVectorIndex.pqApproxDistance (:1057-1065) over n code rows."""
    lut = np.ascontiguousarray(lut, dtype=np.float64)
    M, K = lut.shape
    if k_cent is not None and k_cent != K:
        lut = np.ascontiguousarray(lut[:, :k_cent])
        K = k_cent
    codes = np.ascontiguousarray(codes, dtype=np.uint8).reshape(-1, M)
    out = np.zeros(codes.shape[0], dtype=np.float64)
    check(_ensure().vs_pq_approx_distance(_p(lut, L.f64p), M, K, _p(codes, L.u8p), codes.shape[0],
                                          _p(out, L.f64p)))
    return out


def merge_topk(ids, scores, k: int):
    """query() merge (:432-437): stable sort by score descending of concatenated lists, first k."""
    ids = np.ascontiguousarray(ids, dtype=np.int64).ravel()
    scores = np.ascontiguousarray(scores, dtype=np.float64).ravel()
    oi = np.zeros(k, dtype=np.int64)
    os_ = np.zeros(k, dtype=np.float64)
    cnt = C.c_int32()
    check(_ensure().vs_merge_topk(_p(ids, L.i64p), _p(scores, L.f64p), ids.size, k, _p(oi, L.i64p),
                                  _p(os_, L.f64p), C.byref(cnt)))
    return oi[:cnt.value], os_[:cnt.value]


class Segment:
    """A row range of vectors resident in HBM (ACTIVE/PENDING), plus codebook + codes once SEALED."""

    def __init__(self, handle: int):
        self.handle = handle
        n, d, M, K, base = C.c_int64(), C.c_int32(), C.c_int32(), C.c_int32(), C.c_int64()
        check(L.load().vs_segment_info(handle, C.byref(n), C.byref(d), C.byref(M), C.byref(K), C.byref(base)))
        self.n, self.d, self.id_base = n.value, d.value, base.value
        self.M, self.K = M.value, K.value  # 0 until a codebook is attached

    # -- residency -----------------------------------------------------------------------------
    @classmethod
    def upload(cls, rows, skip=None, id_base: int = 0) -> "Segment":
        lib = _ensure()
        rows = _f32(rows)
        if rows.ndim != 2:
            raise ValueError("rows must be [n][d]")
        n, d = rows.shape
        sp = None
        if skip is not None:
            skip = np.ascontiguousarray(skip, dtype=np.uint8)
            if skip.size != n:
                raise ValueError("skip mask must have one byte per row")
            sp = _p(skip, L.u8p)
        h = C.c_uint64()
        check(lib.vs_segment_upload(_p(rows, L.f32p), n, d, sp, id_base, C.byref(h)))
        return cls(h.value)

    @classmethod
    def upload_bytes(cls, data: bytes, n: int, d: int, stride: int | None = None, skip=None, id_base: int = 0) -> "Segment":
        """Rows from the reference's stored bytes: FloatPacker little-endian fp32, `stride` bytes between the
        embeddings of consecutive records (vs_segment_upload_strided)."""
        lib = _ensure()
        stride = d * 4 if stride is None else stride
        buf = np.frombuffer(data, dtype=np.uint8)
        if n > 0 and buf.size < (n - 1) * stride + d * 4:
            raise ValueError("byte string too short for n records")
        sp = None
        if skip is not None:
            skip = np.ascontiguousarray(skip, dtype=np.uint8)
            sp = _p(skip, L.u8p)
        h = C.c_uint64()
        check(lib.vs_segment_upload_strided(buf.ctypes.data_as(C.c_void_p), n, d, stride, sp, id_base, C.byref(h)))
        return cls(h.value)

    @classmethod
    def upload_records(cls, records, d: int, id_base: int = 0):
        """Rows from serialized VectorRecord messages (vectorsearch.proto:108-127), as a segment's range read returns
        them.  Returns (segment, vec_ids): `deleted` records become skipped rows (vs_segment_upload_records)."""
        lib = _ensure()
        offs = np.zeros(len(records) + 1, dtype=np.int64)
        for i, r in enumerate(records):
            offs[i + 1] = offs[i] + len(r)
        buf = np.frombuffer(b"".join(records), dtype=np.uint8) if records else np.zeros(1, np.uint8)
        vec_ids = np.zeros(len(records), dtype=np.int32)
        h = C.c_uint64()
        check(lib.vs_segment_upload_records(buf.ctypes.data_as(C.c_void_p), _p(offs, L.i64p), len(records), d, id_base,
                                            _p(vec_ids, L.i32p), C.byref(h)))
        return cls(h.value), vec_ids

    @classmethod
    def generate(cls, seed: int, first_row: int, n: int, d: int, id_base: int = 0) -> "Segment":
        h = C.c_uint64()
        check(_ensure().vs_segment_generate(seed, first_row, n, d, id_base, C.byref(h)))
        return cls(h.value)

    def set_skip(self, skip) -> None:
        sp = None
        if skip is not None:
            skip = np.ascontiguousarray(skip, dtype=np.uint8)
            if skip.size != self.n:
                raise ValueError("skip mask must have one byte per row")
            sp = _p(skip, L.u8p)
        check(L.load().vs_segment_set_skip(self.handle, sp))

    def rows(self, first: int = 0, count: int | None = None) -> np.ndarray:
        count = self.n - first if count is None else count
        out = np.zeros((count, self.d), dtype=np.float32)
        check(L.load().vs_segment_download_rows(self.handle, first, count, _p(out, L.f32p)))
        return out

    def attach_pq(self, centroids, codes=None) -> None:
        c = _f32(centroids)
        M, K, sub = c.shape
        if M * sub != self.d:
            raise ValueError("Invalid PQ params (m,k,dimension)")
        cp = None
        if codes is not None:
            codes = np.ascontiguousarray(codes, dtype=np.uint8)
            if codes.shape != (self.n, M):
                raise ValueError("codes must be [n][M]")
            cp = _p(codes, L.u8p)
        check(L.load().vs_segment_attach_pq(self.handle, _p(c, L.f32p), M, K, cp))
        self.M, self.K = M, K

    def attach_pq_codebook(self, codebook: bytes, codes=None) -> None:
        """attach_pq from the stored PQCodebook message (SegmentBuildService.buildCodebookBytes)."""
        buf = np.frombuffer(codebook, dtype=np.uint8)
        cp = None
        if codes is not None:
            codes = np.ascontiguousarray(codes, dtype=np.uint8)
            cp = _p(codes, L.u8p)
        check(L.load().vs_segment_attach_pq_codebook(self.handle, buf.ctypes.data_as(C.c_void_p), buf.size, cp))
        M, K = C.c_int32(), C.c_int32()
        check(L.load().vs_segment_info(self.handle, None, None, C.byref(M), C.byref(K), None))
        self.M, self.K = M.value, K.value

    def codes(self, first: int = 0, count: int | None = None) -> np.ndarray:
        count = self.n - first if count is None else count
        out = np.zeros((count, self.M), dtype=np.uint8)
        check(L.load().vs_segment_download_codes(self.handle, first, count, _p(out, L.u8p)))
        return out

    def free(self) -> None:
        if self.handle:
            check(L.load().vs_segment_free(self.handle))
            self.handle = 0

    # -- queries ----------------------------------------------------------------------------------
    @staticmethod
    def _queries(q, d):
        q = _f32(q)
        if q.ndim == 1:
            q = q.reshape(1, -1)
        if q.shape[1] != d:
            raise ValueError("query dimension does not match the segment")
        return q

    @staticmethod
    def _trim(ids, scores, counts, single):
        if single:
            return ids[0, :counts[0]], scores[0, :counts[0]]
        return ids, scores, counts

    def bruteforce_topk(self, q, k: int, metric: int = METRIC_L2):
        """searchBruteForceSegment scoring + sort + subList (:676-721).  q [d] -> (ids, scores);
        q [nq][d] -> (ids [nq][k], scores [nq][k], counts [nq]); score = -l2 or cosine similarity."""
        q, qp, nq, single = _query_ptr(q, self.d)
        b = _out_buffers(nq, k)
        check(L.load().vs_bruteforce_topk(self.handle, qp, nq, k, metric, b[3], b[4], b[5]))
        return _results(b, nq, single)

    def adc_topk(self, q, n_cand: int):
        """buildLut + ADC scan + ascending stable sort + first n_cand (:741,:754-769,:820-822)."""
        single = np.ndim(q) == 1
        q = self._queries(q, self.d)
        nq = q.shape[0]
        ids = np.zeros((nq, n_cand), dtype=np.int64)
        ap = np.zeros((nq, n_cand), dtype=np.float64)
        cn = np.zeros(nq, dtype=np.int32)
        check(L.load().vs_adc_topk(self.handle, _p(q, L.f32p), nq, n_cand, _p(ids, L.i64p), _p(ap, L.f64p),
                                   _p(cn, L.i32p)))
        return self._trim(ids, ap, cn, single)

    def rerank_topk(self, q, cand_ids, k: int, metric: int = METRIC_L2, normalize_on_read: bool = False):
        """fetchExactAndScore (:997-1043): candidates scored in the given order, ties keep it."""
        q = _f32(q).ravel()
        cand = np.ascontiguousarray(cand_ids, dtype=np.int64).ravel()
        ids = np.zeros(k, dtype=np.int64)
        sc = np.zeros(k, dtype=np.float64)
        cn = C.c_int32()
        check(L.load().vs_rerank_topk(self.handle, _p(q, L.f32p), _p(cand, L.i64p), cand.size, k, metric,
                                      int(bool(normalize_on_read)), _p(ids, L.i64p), _p(sc, L.f64p),
                                      C.byref(cn)))
        return ids[:cn.value], sc[:cn.value]

    def adc_rerank_topk(self, q, n_cand: int, k: int, metric: int = METRIC_L2, normalize_on_read: bool = False):
        """ADC top n_cand followed by exact re-rank to k in one call (config C4)."""
        q, qp, nq, single = _query_ptr(q, self.d)
        b = _out_buffers(nq, k)
        check(L.load().vs_adc_rerank_topk(self.handle, qp, nq, n_cand, k, metric, 1 if normalize_on_read else 0, b[3], b[4], b[5]))
        return _results(b, nq, single)

    # -- graph construction ----------------------------------------------------------------------------
    def knn_graph(self, degree: int, l_build: int = 0, alpha: float = 1.0):
        """GraphBuilder.buildL2Neighbors(vectors, degree) (l_build == 0) or buildPrunedNeighbors(vectors, degree,
        l_build, alpha) over this segment's rows -> list of int32 arrays (neighbour rows per node)."""
        nb = np.full((self.n, degree), -1, dtype=np.int32)
        cn = np.zeros(self.n, dtype=np.int32)
        check(L.load().vs_knn_graph(self.handle, degree, l_build, float(alpha), _p(nb, L.i32p), _p(cn, L.i32p)))
        return [nb[i, :cn[i]] for i in range(self.n)]

    # -- BEST_FIRST expansion scoring -------------------------------------------------------------------
    def adc_query(self, q) -> "AdcQuery":
        """LUT of `q` against this sealed segment, kept on the device (buildLut once per query and segment,
        FdbVectorIndex.java:741); score id lists with .gather() per expansion step (:950-963)."""
        return AdcQuery(self, q)

    def adc_gather(self, q, ids):
        """One-shot: pqApproxDistance of the listed ids -> (distances, valid)."""
        q = _f32(q).ravel()
        ids = np.ascontiguousarray(ids, dtype=np.int64).ravel()
        out = np.zeros(ids.size, dtype=np.float64)
        valid = np.zeros(ids.size, dtype=np.uint8)
        check(L.load().vs_adc_gather(self.handle, _p(q, L.f32p), _p(ids, L.i64p), ids.size, _p(out, L.f64p), _p(valid, L.u8p)))
        return out, valid.astype(bool)


class AdcQuery:
    """vs_adc_query_begin / _gather / _end as a context manager."""

    def __init__(self, segment: Segment, q):
        q = _f32(q).ravel()
        if q.size != segment.d:
            raise ValueError("query dimension does not match the segment")
        h = C.c_uint64()
        check(L.load().vs_adc_query_begin(segment.handle, _p(q, L.f32p), C.byref(h)))
        self.handle = h.value

    def gather(self, ids):
        ids = np.ascontiguousarray(ids, dtype=np.int64).ravel()
        out = np.zeros(ids.size, dtype=np.float64)
        valid = np.zeros(ids.size, dtype=np.uint8)
        check(L.load().vs_adc_query_gather(self.handle, _p(ids, L.i64p), ids.size, _p(out, L.f64p), _p(valid, L.u8p)))
        return out, valid.astype(bool)

    def close(self) -> None:
        if self.handle:
            check(L.load().vs_adc_query_end(self.handle))
            self.handle = 0

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()


# ---- wire formats (vectorsearch.proto) and residency -------------------------------------------------------------
def codebook_encode(centroids) -> bytes:
    """float[M][K][subDim] -> serialized PQCodebook, the bytes SegmentBuildService.buildCodebookBytes stores."""
    c = _f32(centroids)
    M, K, sub = c.shape
    n = C.c_int64()
    check(L.load().vs_codebook_encode(_p(c, L.f32p), M, K, sub, None, 0, C.byref(n)))
    out = np.zeros(n.value, dtype=np.uint8)
    check(L.load().vs_codebook_encode(_p(c, L.f32p), M, K, sub, out.ctypes.data_as(C.c_void_p), out.size, C.byref(n)))
    return out.tobytes()


def codebook_decode(data: bytes) -> np.ndarray:
    """Serialized PQCodebook -> float[M][K][subDim] (SegmentCaches.decodeCodebook)."""
    buf = np.frombuffer(data, dtype=np.uint8) if len(data) else np.zeros(1, np.uint8)
    M, K, sub = C.c_int32(), C.c_int32(), C.c_int32()
    check(L.load().vs_codebook_decode(buf.ctypes.data_as(C.c_void_p), len(data), None, 0, C.byref(M), C.byref(K), C.byref(sub)))
    out = np.zeros((M.value, K.value, sub.value), dtype=np.float32)
    check(L.load().vs_codebook_decode(buf.ctypes.data_as(C.c_void_p), len(data), _p(out, L.f32p), out.size, None, None, None))
    return out


STATE_ACTIVE, STATE_PENDING, STATE_SEALED, STATE_COMPACTING, STATE_WRITING = 0, 1, 2, 3, 4


class Residency:
    """The residency table of libvsgpu keyed by (segment id, SegmentMeta.State); it owns the handles it holds."""

    @staticmethod
    def put(seg_id: int, state: int, segment: Segment) -> None:
        check(L.load().vs_residency_put(seg_id, state, segment.handle))

    @staticmethod
    def get(seg_id: int, state: int):
        """-> Segment if resident in `state`; None if not resident; raises VsError(VS_ESTATE) if resident in another state."""
        h = C.c_uint64()
        rc = L.load().vs_residency_get(seg_id, state, C.byref(h))
        if rc == L.VS_EHANDLE:
            return None
        check(rc)
        return Segment(h.value)

    @staticmethod
    def peek(seg_id: int, state: int):
        """-> (Segment or None, in_requested_state)."""
        h = C.c_uint64()
        rc = L.load().vs_residency_get(seg_id, state, C.byref(h))
        if rc == L.VS_EHANDLE:
            return None, False
        if rc == L.VS_ESTATE:
            return Segment(h.value), False
        check(rc)
        return Segment(h.value), True

    @staticmethod
    def invalidate(seg_id: int) -> None:
        check(L.load().vs_residency_invalidate(seg_id))

    @staticmethod
    def set_budget(nbytes: int) -> None:
        check(L.load().vs_residency_set_budget(nbytes))

    @staticmethod
    def stats() -> dict:
        n, b = C.c_int64(), C.c_int64()
        check(L.load().vs_residency_stats(C.byref(n), C.byref(b)))
        return {"segments": n.value, "bytes": b.value}
