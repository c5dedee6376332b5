"""Drop-in for the reference's ``gpu_bm25.common.gpu_execute_query`` (gpu_bm25/common.py:28-85).

The reference uploads the dense docs x terms score matrix, builds and JIT-compiles a MAX graph
``gather(axis=1) -> sum(-1) -> transpose -> top_k(1)`` on every call and returns
``(top index, top weight)`` as [1,1] tensors.  Here the dense matrix is converted once to CSC,
pinned in HBM, and the query runs through libbm25_b200.so; ``session`` / ``device`` are accepted
for signature compatibility and ignored (there is no MAX engine).  Returned objects are [1,1]
torch CPU tensors, which satisfy the ``.item()`` access of the reference's callers (main.py:251).
"""
from __future__ import annotations

import weakref
from typing import Tuple

import numpy as np

from ..engine import DeviceIndex

_CACHE: dict = {}


def dense_to_csc(score_matrix: np.ndarray):
    """docs x terms dense fp32 -> (indptr, indices, data) CSC with explicit zeros dropped."""
    m = np.asarray(score_matrix)
    if m.ndim != 2:
        raise ValueError("score_matrix must be 2-D [docs, terms]")
    cols, rows = np.nonzero(m.T)  # column-major order, rows ascending inside a column
    indptr = np.zeros(m.shape[1] + 1, dtype=np.int32)
    np.cumsum(np.bincount(cols, minlength=m.shape[1]), out=indptr[1:])
    return indptr, rows.astype(np.int32), m[rows, cols].astype(np.float32)


def _index_for(score_matrix: np.ndarray, device_ordinal: int) -> DeviceIndex:
    key = (id(score_matrix), score_matrix.shape, device_ordinal)
    hit = _CACHE.get(key)
    if hit is not None and hit[0]() is score_matrix:
        return hit[1]
    indptr, indices, data = dense_to_csc(score_matrix)
    index = DeviceIndex(indptr, indices, data, score_matrix.shape[0], device=device_ordinal)
    try:
        _CACHE.clear()
        _CACHE[key] = (weakref.ref(score_matrix), index)
    except TypeError:
        pass
    return index


def gpu_execute_query(score_matrix, query_vector, session=None, device=None, k: int = 1) -> Tuple[object, object]:
    import torch

    ordinal = device if isinstance(device, int) else 0
    index = _index_for(score_matrix, ordinal)
    q = np.asarray(query_vector)
    if q.ndim != 1:
        raise ValueError("query_vector must be 1-D [num_query_terms]")
    q = q.astype(np.int32).reshape(1, -1)
    if q.size and int(q.max()) >= index.n_terms:
        raise ValueError("query term id out of range")
    ids, scores = index.search(q, k)
    return torch.from_numpy(ids.astype(np.int64)), torch.from_numpy(scores)
