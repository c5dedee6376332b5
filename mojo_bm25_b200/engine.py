"""DeviceIndex: thin Python owner of a `bm25_index*` handle (include/bm25_b200.h).

Everything numerical happens inside libbm25_b200.so; this file only marshals numpy arrays / torch
CUDA tensors into raw pointers.  There is no CPU fallback.
"""
from __future__ import annotations

import ctypes
from typing import Optional, Tuple

import numpy as np

from . import _lib


def _as_c(a, dtype) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=dtype)


def _as_i32(a, what: str) -> np.ndarray:
    """int32 view/copy of an integer array; values that do not fit raise instead of wrapping
    (bm25s writes int32 arrays, scipy may hand over int64 indptr/indices)."""
    a = np.asarray(a)
    if a.dtype != np.int32:
        if a.dtype.kind not in "iu":
            raise ValueError(f"{what} must be an integer array (got {a.dtype})")
        if a.size and (int(a.max()) > np.iinfo(np.int32).max or int(a.min()) < np.iinfo(np.int32).min):
            raise ValueError(f"{what} holds values outside int32: shard the index by document range "
                             "(one handle holds fewer than 2^31 postings)")
    return np.ascontiguousarray(a, dtype=np.int32)


def _ptr(a: np.ndarray):
    return ctypes.c_void_p(a.ctypes.data)


def _pinned_outputs(qn: int, k: int) -> Tuple[np.ndarray, np.ndarray]:
    """(ids int32 [Q,k], scores fp32 [Q,k]) as numpy views of ONE fresh page-locked block [2,Q,k]:
    ``bm25_search_host`` then delivers the results with a single DMA and no staging copy.  The
    block comes from torch's caching host allocator and lives as long as either array does."""
    if qn * k == 0:
        return np.empty((qn, k), np.int32), np.empty((qn, k), np.float32)
    import torch

    block = torch.empty((2, qn, k), dtype=torch.int32, pin_memory=True).numpy()
    return block[0], block[1].view(np.float32)


def round_to_bf16(data) -> np.ndarray:
    """fp32 array -> the fp32 values a compressed (bf16) handle stores: round to nearest even on the
    upper 16 bits.  Host-side twin of the library's k_quantize, for callers who want to know exactly
    which matrix a compressed index scores with."""
    u = np.ascontiguousarray(data, dtype=np.float32).view(np.uint32).astype(np.uint64)
    r = (u + 0x7FFF + ((u >> 16) & 1)) & 0xFFFF0000
    return r.astype(np.uint32).view(np.float32)


class DeviceIndex:
    """An HBM-resident CSC BM25 index (terms = columns, documents = rows).

    Parameters mirror what `BM25v.index` receives (reference bm25_native.py:59-74) plus the
    on-disk bm25s layout (indptr / indices / data, animal_index_bm25/*.csc.index.npy).
    """

    def __init__(self, indptr, indices, data, n_docs: int, device: int = 0, doc_id_base: int = 0):
        lib = _lib.load()
        indptr = _as_i32(indptr, "indptr")
        indices = _as_i32(indices, "indices")
        data = _as_c(data, np.float32)
        if indptr.ndim != 1 or indptr.shape[0] < 1:
            raise ValueError("indptr must be a 1-D array with at least one element")
        if indices.shape != data.shape or indices.ndim != 1:
            raise ValueError("indices and data must be 1-D arrays of equal length")
        self._h = ctypes.c_void_p()
        self._keepalive = None
        _lib.check(lib.bm25_index_create(_ptr(indptr), _ptr(indices), _ptr(data), indptr.shape[0] - 1,
                                         int(n_docs), indices.shape[0], int(device), int(doc_id_base),
                                         ctypes.byref(self._h)))
        self.device = int(device)

    @classmethod
    def from_torch(cls, indptr, indices, data, n_docs: int, doc_id_base: int = 0, borrow: bool = True):
        """Build from CUDA tensors already in HBM (int32 indptr/indices, fp32 data, canonical CSC).
        The library re-buckets the index into its own memory during this call; ``borrow`` is kept
        for source compatibility and no longer keeps the tensors alive."""
        import torch

        lib = _lib.load()
        if not (indptr.is_cuda and indices.is_cuda and data.is_cuda):
            raise ValueError("from_torch expects CUDA tensors")
        if indptr.dtype != torch.int32 or indices.dtype != torch.int32 or data.dtype != torch.float32:
            raise ValueError("from_torch expects int32 indptr/indices and float32 data")
        indptr, indices, data = indptr.contiguous(), indices.contiguous(), data.contiguous()
        torch.cuda.synchronize(indices.device)
        self = cls.__new__(cls)
        self._h = ctypes.c_void_p()
        self.device = indices.device.index or 0
        self._keepalive = None
        _lib.check(lib.bm25_index_create_device(
            ctypes.c_void_p(indptr.data_ptr()), ctypes.c_void_p(indices.data_ptr()),
            ctypes.c_void_p(data.data_ptr()), indptr.numel() - 1, int(n_docs), indices.numel(),
            self.device, int(doc_id_base), 1 if borrow else 0, ctypes.byref(self._h)))
        return self

    # ------------------------------------------------------------------------------------------
    def close(self):
        h, self._h = getattr(self, "_h", None), None
        if h:
            _lib.load().bm25_index_destroy(h)
        self._keepalive = None

    def __del__(self):
        try:
            self.close()
        except Exception:  # noqa: BLE001
            pass

    def _handle(self):
        if not self._h:
            raise ValueError("index is closed")
        return self._h

    @property
    def info(self) -> _lib.IndexInfo:
        out = _lib.IndexInfo()
        _lib.check(_lib.load().bm25_index_get_info(self._handle(), ctypes.byref(out)))
        return out

    @property
    def n_docs(self) -> int:
        return int(self.info.n_docs)

    @property
    def n_terms(self) -> int:
        return int(self.info.n_terms)

    def set_option(self, name: str, value: int):
        _lib.check(_lib.load().bm25_index_set_option(self._handle(), name.encode(), int(value)))

    def compress(self, weight_format: str = "bf16") -> "DeviceIndex":
        """Convert the handle in place to the compressed posting format (``bm25_index_compress``):
        weights rounded to bf16, 4-byte postings {uint16 tile-local slot, bf16 weight} for the score
        kernel.  Results are from then on those of the CSC matrix ``round_to_bf16(data)``.  Irreversible."""
        if weight_format not in ("bf16", _lib.WEIGHTS_BF16):
            raise ValueError("weight_format must be 'bf16'")
        _lib.check(_lib.load().bm25_index_compress(self._handle(), _lib.WEIGHTS_BF16))
        return self

    def last_timing_ms(self):
        """(segments, score+topk, merge) device ms of the last search; needs set_option("timing", 1)."""
        out = (ctypes.c_float * 3)()
        _lib.check(_lib.load().bm25_index_get_timing(self._handle(), out))
        return float(out[0]), float(out[1]), float(out[2])

    # ------------------------------------------------------------------------------------------
    def search(self, queries: np.ndarray, k: int) -> Tuple[np.ndarray, np.ndarray]:
        """Host-buffer search: ``queries`` int32 [Q,T] (-1 padded) -> (int32 [Q,k], fp32 [Q,k])."""
        q = _as_c(queries, np.int32)
        if q.ndim != 2:
            raise ValueError("queries must be 2-D [Q, T]")
        qn, tn = q.shape
        ids, scores = _pinned_outputs(qn, int(k))  # fully overwritten by the library
        if tn == 0:
            q = np.full((qn, 1), -1, np.int32)
            tn = 1
        _lib.check(_lib.load().bm25_search_host(self._handle(), _ptr(q), qn, tn, int(k), _ptr(ids), _ptr(scores)))
        return ids, scores

    def search_device(self, queries, k: int, out_ids=None, out_scores=None, stream: Optional[int] = None):
        """Stream-ordered search on torch CUDA tensors (queries int32 [Q,T]); returns
        (ids int32 [Q,k], scores fp32 [Q,k]) CUDA tensors.  Asynchronous."""
        import torch

        if not queries.is_cuda or queries.dtype != torch.int32 or queries.dim() != 2:
            raise ValueError("queries must be a 2-D int32 CUDA tensor")
        queries = queries.contiguous()
        qn, tn = queries.shape
        if out_ids is None:
            out_ids = torch.empty((qn, k), dtype=torch.int32, device=queries.device)
        if out_scores is None:
            out_scores = torch.empty((qn, k), dtype=torch.float32, device=queries.device)
        if stream is None:
            stream = torch.cuda.current_stream(queries.device).cuda_stream
        _lib.check(_lib.load().bm25_search(self._handle(), ctypes.c_void_p(queries.data_ptr()), qn, tn, int(k),
                                           ctypes.c_void_p(out_ids.data_ptr()),
                                           ctypes.c_void_p(out_scores.data_ptr()), ctypes.c_void_p(stream)))
        return out_ids, out_scores

    def scores_dense(self, queries: np.ndarray) -> np.ndarray:
        """Dense per-query score vectors, fp32 [Q, n_docs] (parity/debug)."""
        q = _as_c(queries, np.int32)
        if q.ndim != 2:
            raise ValueError("queries must be 2-D [Q, T]")
        out = np.zeros((q.shape[0], self.n_docs), np.float32)
        if q.shape[1] == 0 or q.shape[0] == 0:
            return out
        _lib.check(_lib.load().bm25_scores_dense_host(self._handle(), _ptr(q), q.shape[0], q.shape[1], _ptr(out)))
        return out

    def posting_bytes(self, queries: np.ndarray, k: int) -> int:
        q = _as_c(queries, np.int32)
        out = ctypes.c_int64(0)
        _lib.check(_lib.load().bm25_posting_bytes(self._handle(), _ptr(q), q.shape[0], q.shape[1], int(k),
                                                  ctypes.byref(out)))
        return int(out.value)


class SplitDeviceIndex:
    """A CSC index with more postings than one handle addresses (offsets are int32: < 2^31 postings
    per handle) -- e.g. an int64-``indptr`` matrix -- held on ONE device as several document-range
    handles.  Built on the host: the document range is cut so that every part has at most
    ``max_postings`` postings, each part becomes a ``DeviceIndex`` with local doc ids and its
    ``doc_id_base``; ``search`` runs every part and merges with ``bm25_merge_topk`` (the same path
    the multi-GPU document sharding takes)."""

    def __init__(self, indptr, indices, data, n_docs: int, device: int = 0, max_postings: int = (1 << 31) - (1 << 24)):
        indptr = np.asarray(indptr)
        if indptr.dtype.kind not in "iu":
            raise ValueError("indptr must be an integer array")
        indptr = indptr.astype(np.int64, copy=False)
        indices = np.asarray(indices)
        data = np.ascontiguousarray(data, dtype=np.float32)
        n_docs, n_terms, nnz = int(n_docs), indptr.shape[0] - 1, int(indptr[-1])
        if indices.shape[0] != nnz or data.shape[0] != nnz:
            raise ValueError("indptr[-1] must equal the number of postings")
        if indices.size and (int(indices.max()) >= n_docs or int(indices.min()) < 0):
            raise ValueError("document id outside [0, n_docs)")
        # cut the document range where the cumulative posting count crosses multiples of max_postings
        per_doc = np.bincount(indices, minlength=n_docs).astype(np.int64)
        cum = np.cumsum(per_doc)
        cuts, start = [0], 0
        while start < n_docs:
            base = cum[start - 1] if start else 0
            end = int(np.searchsorted(cum, base + max_postings, side="right"))
            end = max(end, start + 1)  # a single document never exceeds the limit in practice
            cuts.append(min(end, n_docs))
            start = cuts[-1]
        self.parts, self.n_docs, self.n_terms, self.device = [], n_docs, n_terms, int(device)
        col_ends = indptr[1:]
        for lo, hi in zip(cuts[:-1], cuts[1:]):
            sel = np.flatnonzero((indices >= lo) & (indices < hi))
            col = np.searchsorted(col_ends, sel, side="right")  # column of every selected posting
            sub_ptr = np.zeros(n_terms + 1, dtype=np.int64)
            np.cumsum(np.bincount(col, minlength=n_terms), out=sub_ptr[1:])
            self.parts.append(DeviceIndex(sub_ptr, (indices[sel] - lo).astype(np.int32), data[sel], hi - lo,
                                          device=device, doc_id_base=lo))
        self._searchers = {}

    def close(self):
        for p in self.parts:
            p.close()
        self.parts = []

    def search(self, queries: np.ndarray, k: int) -> Tuple[np.ndarray, np.ndarray]:
        """Host-buffer search like ``DeviceIndex.search``."""
        import torch

        from .sharded import DocShardedSearcher

        q = _as_c(queries, np.int32)
        if q.ndim != 2:
            raise ValueError("queries must be 2-D [Q, T]")
        if k > self.n_docs:
            raise ValueError(f"kth(=-{k}) out of bounds ({self.n_docs})")
        if q.size and int(q.max(initial=-1)) >= self.n_terms:
            raise ValueError("The maximum token ID in the query is higher than the number of tokens in the index.")
        if q.shape[1] == 0:
            q = np.full((q.shape[0], 1), -1, np.int32)
        s = self._searchers.get(int(k))
        if s is None:
            s = self._searchers[int(k)] = DocShardedSearcher.from_index(self.parts, int(k))
        dev = torch.device("cuda", self.device)
        ids, sc = s.search(torch.from_numpy(q).to(dev))
        torch.cuda.synchronize(dev)
        return ids.cpu().numpy(), sc.cpu().numpy()


def open_index(indptr, indices, data, n_docs: int, device: int = 0, doc_id_base: int = 0):
    """``DeviceIndex`` when the postings fit one handle, else ``SplitDeviceIndex``."""
    nnz = int(np.asarray(indptr)[-1]) if len(indptr) else 0
    if nnz < (1 << 31) - (1 << 24):
        return DeviceIndex(indptr, indices, data, n_docs, device=device, doc_id_base=doc_id_base)
    if doc_id_base:
        raise ValueError("doc_id_base is not supported for an index that is split into several handles")
    return SplitDeviceIndex(indptr, indices, data, n_docs, device=device)


def merge_topk_device(ids, scores, k_out: int, stream: Optional[int] = None, list_stride: int = 0,
                      n_lists: Optional[int] = None, n_queries: Optional[int] = None, k_in: Optional[int] = None):
    """Merge all-gathered shard results into the global top-k (bm25_merge_topk).

    Dense form: ids/scores CUDA tensors [L, Q, k_in].  Packed form (one all-gather buffer
    [L][2][Q][k_in]): pass the id and score views of list 0 plus ``list_stride`` (elements),
    ``n_lists``, ``n_queries`` and ``k_in`` explicitly."""
    import torch

    if list_stride == 0:
        if ids.dim() != 3 or ids.shape != scores.shape:
            raise ValueError("ids/scores must be [n_lists, Q, k_in]")
        ids, scores = ids.contiguous(), scores.contiguous()
        n_lists, n_queries, k_in = ids.shape
    if ids.dtype != torch.int32 or scores.dtype != torch.float32:
        raise ValueError("ids must be int32 and scores float32")
    out_ids = torch.empty((n_queries, k_out), dtype=torch.int32, device=ids.device)
    out_scores = torch.empty((n_queries, k_out), dtype=torch.float32, device=ids.device)
    if stream is None:
        stream = torch.cuda.current_stream(ids.device).cuda_stream
    _lib.check(_lib.load().bm25_merge_topk(ctypes.c_void_p(ids.data_ptr()), ctypes.c_void_p(scores.data_ptr()),
                                           int(n_lists), int(list_stride), int(n_queries), int(k_in), int(k_out),
                                           ctypes.c_void_p(out_ids.data_ptr()),
                                           ctypes.c_void_p(out_scores.data_ptr()), ids.device.index or 0,
                                           ctypes.c_void_p(stream)))
    return out_ids, out_scores


def kernel_launches() -> int:
    return int(_lib.load().bm25_kernel_launches())
