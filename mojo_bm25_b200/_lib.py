"""ctypes binding of libbm25_b200.so (the C ABI of include/bm25_b200.h).

The library is the product: if it is missing or cannot be loaded this module raises -- there is
no Python/CPU fallback for the BM25 query path.
"""
import ctypes
import os

from . import build as _build

OK, ERR_INVALID, ERR_CUDA, ERR_NO_DEVICE, ERR_UNSUPPORTED, ERR_OOM = 0, 1, 2, 3, 4, 5
WEIGHTS_FP32, WEIGHTS_BF16 = 0, 1
MAX_K = 65536
SMALL_K = 6144

SYMBOLS = [
    "bm25_index_create", "bm25_index_create_device", "bm25_index_destroy", "bm25_index_get_info",
    "bm25_index_set_option", "bm25_index_compress", "bm25_index_get_timing", "bm25_search", "bm25_search_host", "bm25_scores_dense",
    "bm25_scores_dense_host", "bm25_merge_topk", "bm25_posting_bytes", "bm25_kernel_launches",
    "bm25_last_error", "bm25_version",
]


class IndexInfo(ctypes.Structure):
    _fields_ = [
        ("n_terms", ctypes.c_int64), ("n_docs", ctypes.c_int64), ("nnz", ctypes.c_int64),
        ("doc_id_base", ctypes.c_int64), ("device_bytes", ctypes.c_int64),
        ("device", ctypes.c_int32), ("tile_docs", ctypes.c_int32), ("n_tiles", ctypes.c_int32),
        ("all_positive", ctypes.c_int32), ("was_sorted", ctypes.c_int32), ("sm_count", ctypes.c_int32),
        ("weight_format", ctypes.c_int32), ("posting_bytes", ctypes.c_int32),
    ]


class Bm25Error(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"libbm25_b200 error {code}: {msg}")
        self.code = code
        self.msg = msg


_lib = None


def load(rebuild_if_stale: bool = True):
    """Load (building in-tree first if needed) and declare the C ABI.  Raises on failure."""
    global _lib
    if _lib is not None:
        return _lib
    so = _build.SO
    alt = os.environ.get("BM25_B200_LIB")  # developer A/B switch: another build of the same C ABI
    if alt:
        so, rebuild_if_stale = alt, False
    if rebuild_if_stale and (not os.path.exists(so) or (_build.is_stale() and os.path.exists(_build.nvcc_path()))):
        so = _build.build()
    if not os.path.exists(so):
        raise RuntimeError(f"{so} is missing: build it with `python -m mojo_bm25_b200.build` "
                           "(the BM25 path has no CPU fallback)")
    lib = ctypes.CDLL(so)
    vp, i64, i32 = ctypes.c_void_p, ctypes.c_int64, ctypes.c_int
    lib.bm25_index_create.argtypes = [vp, vp, vp, i64, i64, i64, i32, i64, ctypes.POINTER(vp)]
    lib.bm25_index_create_device.argtypes = [vp, vp, vp, i64, i64, i64, i32, i64, i32, ctypes.POINTER(vp)]
    lib.bm25_index_destroy.argtypes = [vp]
    lib.bm25_index_get_info.argtypes = [vp, ctypes.POINTER(IndexInfo)]
    lib.bm25_index_set_option.argtypes = [vp, ctypes.c_char_p, i64]
    lib.bm25_index_compress.argtypes = [vp, i32]
    lib.bm25_index_get_timing.argtypes = [vp, ctypes.POINTER(ctypes.c_float)]
    lib.bm25_search.argtypes = [vp, vp, i64, i64, i32, vp, vp, vp]
    lib.bm25_search_host.argtypes = [vp, vp, i64, i64, i32, vp, vp]
    lib.bm25_scores_dense.argtypes = [vp, vp, i64, i64, vp, vp]
    lib.bm25_scores_dense_host.argtypes = [vp, vp, i64, i64, vp]
    lib.bm25_merge_topk.argtypes = [vp, vp, i32, i64, i64, i32, i32, vp, vp, i32, vp]
    lib.bm25_posting_bytes.argtypes = [vp, vp, i64, i64, i32, ctypes.POINTER(i64)]
    lib.bm25_kernel_launches.argtypes = []
    lib.bm25_kernel_launches.restype = i64
    lib.bm25_last_error.restype = ctypes.c_char_p
    lib.bm25_version.restype = ctypes.c_char_p
    for name in SYMBOLS:
        fn = getattr(lib, name)
        if name not in ("bm25_kernel_launches", "bm25_last_error", "bm25_version"):
            fn.restype = i32
    _lib = lib
    return lib


def check(rc: int):
    """Translate a status code into the exception the reference's Python API would raise."""
    if rc == OK:
        return
    msg = (load().bm25_last_error() or b"").decode("utf-8", "replace")
    if rc == ERR_INVALID:
        raise ValueError(msg)
    if rc == ERR_OOM:
        raise MemoryError(msg)
    raise Bm25Error(rc, msg)
