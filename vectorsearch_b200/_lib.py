"""ctypes binding of libvsgpu.so (include/vsgpu.h).

The library is the product: there is no Python or CPU fallback.  Importing this module loads the
in-tree shared object and raises if it is missing; every compute call raises if no CUDA device was
bound with vs_init.
"""
from __future__ import annotations

import ctypes as C
import subprocess
from pathlib import Path

_HERE = Path(__file__).resolve().parent
LIB_PATH = _HERE / "libvsgpu.so"

VS_OK, VS_EINVAL, VS_ENOMEM, VS_ECUDA, VS_EHANDLE, VS_ESTATE, VS_EEMPTY = 0, -1, -2, -3, -4, -5, -6
METRIC_L2, METRIC_COSINE = 0, 1

f32p = C.POINTER(C.c_float)
f64p = C.POINTER(C.c_double)
u8p = C.POINTER(C.c_uint8)
i64p = C.POINTER(C.c_int64)
i32p = C.POINTER(C.c_int32)
u64p = C.POINTER(C.c_uint64)
i32, i64, u64, vp = C.c_int32, C.c_int64, C.c_uint64, C.c_void_p

# name -> (restype, argtypes); mirrors include/vsgpu.h one to one
SIGNATURES = {
    "vs_version": (i32, []),
    "vs_last_error": (C.c_char_p, []),
    "vs_init": (i32, [i32]),
    "vs_init_multi": (i32, [i32, i32p]),
    "vs_device_count": (i32, []),
    "vs_shutdown": (i32, []),
    "vs_set_simd_lanes": (i32, [i32]),
    "vs_get_simd_lanes": (i32, []),
    "vs_device_info": (i32, [i32p, i64p, i64p]),
    "vs_l2": (i32, [f32p, f32p, i32, f64p]),
    "vs_l2_squared": (i32, [f32p, f32p, i32, f64p]),
    "vs_dot": (i32, [f32p, f32p, i32, f64p]),
    "vs_norm": (i32, [f32p, i32, f64p]),
    "vs_cosine": (i32, [f32p, f32p, i32, f64p]),
    "vs_pq_encode": (i32, [f32p, i32, i32, i32, f32p, u8p]),
    "vs_pq_lut_distance": (i32, [f32p, i32, i32, u8p, f32p]),
    "vs_build_lut": (i32, [f32p, i32, i32, i32, f32p, f64p]),
    "vs_pq_approx_distance": (i32, [f64p, i32, i32, u8p, i64, f64p]),
    "vs_segment_upload": (i32, [f32p, i64, i32, u8p, i64, u64p]),
    "vs_segment_upload_strided": (i32, [vp, i64, i32, i64, u8p, i64, u64p]),
    "vs_segment_upload_records": (i32, [vp, i64p, i64, i32, i64, i32p, u64p]),
    "vs_segment_generate": (i32, [i64, i64, i64, i32, i64, u64p]),
    "vs_segment_set_skip": (i32, [u64, u8p]),
    "vs_segment_info": (i32, [u64, i64p, i32p, i32p, i32p, i64p]),
    "vs_segment_download_rows": (i32, [u64, i64, i64, f32p]),
    "vs_segment_attach_pq": (i32, [u64, f32p, i32, i32, u8p]),
    "vs_segment_download_codes": (i32, [u64, i64, i64, u8p]),
    "vs_segment_free": (i32, [u64]),
    "vs_codebook_encode": (i32, [f32p, i32, i32, i32, vp, i64, i64p]),
    "vs_codebook_decode": (i32, [vp, i64, f32p, i64, i32p, i32p, i32p]),
    "vs_segment_attach_pq_codebook": (i32, [u64, vp, i64, u8p]),
    "vs_residency_put": (i32, [i64, i32, u64]),
    "vs_residency_get": (i32, [i64, i32, u64p]),
    "vs_residency_invalidate": (i32, [i64]),
    "vs_residency_set_budget": (i32, [i64]),
    "vs_residency_stats": (i32, [i64p, i64p]),
    "vs_adc_query_begin": (i32, [u64, f32p, u64p]),
    "vs_adc_query_gather": (i32, [u64, i64p, i64, f64p, u8p]),
    "vs_adc_query_end": (i32, [u64]),
    "vs_adc_gather": (i32, [u64, f32p, i64p, i64, f64p, u8p]),
    "vs_bruteforce_topk": (i32, [u64, vp, i32, i32, i32, vp, vp, vp]),  # (addresses: the hot calls pass plain integers)
    "vs_adc_topk": (i32, [u64, f32p, i32, i32, i64p, f64p, i32p]),
    "vs_rerank_topk": (i32, [u64, f32p, i64p, i32, i32, i32, i32, i64p, f64p, i32p]),
    "vs_adc_rerank_topk": (i32, [u64, vp, i32, i32, i32, i32, i32, vp, vp, vp]),
    "vs_merge_topk": (i32, [i64p, f64p, i64, i32, i64p, f64p, i32p]),
    "vs_knn_graph": (i32, [u64, i32, i32, C.c_double, i32p, i32p]),
    "vs_pq_train": (i32, [f32p, u64, i64, i32, i32, i32, i32, i64, f32p]),
    "vs_pq_train_sharded": (i32, [u64, i64, i64, i32, i32, i32, i32, i32, i32, i64, vp, vp, vp, vp, f32p]),
    "vs_pq_train_sharded_peer": (i32, [u64, u64, i64, i64, i32, i32, i32, i32, i64, f32p]),
    "vs_pq_encode_batch": (i32, [f32p, i32, i32, i32, f32p, u64, i64, u8p]),
    "vs_bruteforce_topk_dev": (i32, [u64, vp, i32, i32, i32, vp, vp, vp, vp]),
    "vs_adc_topk_dev": (i32, [u64, vp, i32, i32, vp, vp, vp, vp]),
    "vs_adc_rerank_topk_dev": (i32, [u64, vp, i32, i32, i32, i32, i32, vp, vp, vp, vp]),
    "vs_bruteforce_topk_packed_dev": (i32, [u64, vp, i32, i32, i32, vp, vp, vp]),
    "vs_merge_packed_dev": (i32, [vp, i32, i32, i32, i32, vp, vp, vp, vp]),
    "vs_merge_topk_dev": (i32, [vp, vp, i64, i32, vp, vp, vp, vp]),
    "vs_adc_rerank_packed_dev": (i32, [u64, vp, i32, i32, i32, i32, vp, vp]),
    "vs_merge_adc_rerank_packed_dev": (i32, [vp, i32, i32, i32, i32, vp, vp, vp, vp]),
    "vs_adc_topk_packed_dev": (i32, [u64, vp, i32, i32, vp, vp, vp]),
    "vs_rerank_packed_dev": (i32, [u64, vp, vp, i32, i32, i32, vp, vp]),
    "vs_peer_create": (i32, [i32, i32, i64, i32, C.POINTER(u64), vp]),
    "vs_peer_connect": (i32, [u64, vp]),
    "vs_peer_base": (i32, [u64, C.POINTER(u64)]),
    "vs_peer_connect_ptrs": (i32, [u64, C.POINTER(u64)]),
    "vs_peer_release_stream": (i32, [u64, vp]),
    "vs_peer_destroy": (i32, [u64]),
    "vs_exchange_merge_packed_dev": (i32, [u64, vp, i32, i32, i32, vp, vp, vp, vp]),
    "vs_bruteforce_topk_exchange": (i32, [u64, u64, vp, i32, i32, i32, vp, vp, vp]),
    "vs_adc_rerank_topk_exchange": (i32, [u64, u64, vp, i32, i32, i32, i32, i32, vp, vp, vp]),
    "vs_exchange_merge_adc_rerank_packed_dev": (i32, [u64, vp, i32, i32, i32, vp, vp, vp, vp]),
    "vs_bruteforce_topk_exchange_dev": (i32, [u64, u64, vp, i32, i32, i32, vp, vp, vp, vp]),
    "vs_adc_rerank_topk_exchange_dev": (i32, [u64, u64, vp, i32, i32, i32, i32, i32, vp, vp, vp, vp]),
    "vs_adc_topk_exchange": (i32, [u64, u64, f32p, i32, i32, i64p, f64p, i32p]),
    "vs_rerank_topk_exchange": (i32, [u64, u64, f32p, i64p, i32, i32, i32, i32, i64p, f64p, i32p]),
    "vs_kernel_launch_count": (i64, []),
    "vs_set_option": (i32, [C.c_char_p, i64]),
    "vs_debug_batch_groupmins": (i32, [u64, f32p, i32, i32, f32p, i64, i64p, i32p, f64p]),
}


ALLREDUCE_FN = C.CFUNCTYPE(C.c_int32, C.c_void_p, C.c_int32, C.c_int64)  # vs_allreduce_fn


class VsError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"libvsgpu error {code}: {msg}")
        self.code = code


def build(verbose: bool = False) -> Path:
    """Compile libvsgpu.so for sm_100a with nvcc (Makefile in csrc/)."""
    r = subprocess.run(["make", "-C", str(_HERE / "csrc"), "-j8"], capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("libvsgpu build failed:\n" + r.stdout[-4000:] + r.stderr[-4000:])
    if verbose:
        print(r.stdout[-2000:])
    return LIB_PATH


_lib = None


def load() -> C.CDLL:
    """Load the in-tree CUDA library.  Raises if it has not been built: there is no fallback."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise ImportError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                          "(nvcc, sm_100a).  vectorsearch_b200 has no CPU fallback.")
    lib = C.CDLL(str(LIB_PATH))
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the library does not export a declared symbol
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc: int) -> None:
    if rc == VS_OK:
        return
    msg = load().vs_last_error().decode("utf-8", "replace")
    if rc == VS_EINVAL:
        raise ValueError(msg)  # IllegalArgumentException
    if rc == VS_EEMPTY:
        raise IndexError(msg)  # IndexOutOfBoundsException
    if rc == VS_ENOMEM:
        raise MemoryError(msg)
    raise VsError(rc, msg)  # IllegalStateException
