"""ctypes wrapper of the C oracle (oracle/bm25_oracle.c) -- TEST INFRASTRUCTURE ONLY."""
import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(HERE, "libbm25_oracle.so")
SRC = os.path.join(HERE, "bm25_oracle.c")


def build(force: bool = False) -> str:
    if force or not os.path.exists(SO) or os.path.getmtime(SO) < os.path.getmtime(SRC):
        subprocess.check_call(["gcc", "-O3", "-fopenmp", "-shared", "-fPIC", "-o", SO, SRC])
    return SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(build())
        p = ctypes.c_void_p
        _lib.oracle_search.argtypes = [p, p, p, ctypes.c_int64, p, ctypes.c_int64, ctypes.c_int64,
                                       ctypes.c_int, p, p, ctypes.c_int]
        _lib.oracle_search.restype = ctypes.c_int
        _lib.oracle_scores_dense.argtypes = [p, p, p, ctypes.c_int64, p, ctypes.c_int64, p]
        _lib.oracle_scores_dense.restype = None
        _lib.oracle_max_threads.restype = ctypes.c_int
    return _lib


def _c(a, dt):
    return np.ascontiguousarray(a, dtype=dt)


def search(indptr, indices, data, n_docs, queries, k, n_threads=0):
    indptr, indices, data = _c(indptr, np.int32), _c(indices, np.int32), _c(data, np.float32)
    queries = _c(queries, np.int32)
    q_n, t_n = queries.shape
    ids = np.zeros((q_n, k), np.int32)
    sc = np.zeros((q_n, k), np.float32)
    rc = lib().oracle_search(indptr.ctypes.data, indices.ctypes.data, data.ctypes.data, n_docs,
                             queries.ctypes.data, q_n, t_n, k, ids.ctypes.data, sc.ctypes.data, n_threads)
    if rc != 0:
        raise ValueError(f"kth(=-{k}) out of bounds ({n_docs})")
    return ids, sc


def scores_dense(indptr, indices, data, n_docs, query):
    indptr, indices, data = _c(indptr, np.int32), _c(indices, np.int32), _c(data, np.float32)
    query = _c(query, np.int32)
    out = np.empty(n_docs, np.float32)
    lib().oracle_scores_dense(indptr.ctypes.data, indices.ctypes.data, data.ctypes.data, n_docs,
                              query.ctypes.data, query.shape[0], out.ctypes.data)
    return out


def max_threads():
    return lib().oracle_max_threads()
