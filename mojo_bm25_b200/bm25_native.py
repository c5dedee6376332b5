"""Drop-in for the reference's ``bm25_native.BM25v`` (bm25_native.py:32-158): same constructor,
``index(doc_toks, doc_lengths)`` and ``search(queries, top_k)`` signatures, return shapes/dtypes and
error behaviour -- but the per-query hot loop (bm25_native.py:129-158: CSC column gather ->
scatter-add -> top-k) runs in libbm25_b200.so on the GPU.  No CPU fallback.
"""
from __future__ import annotations

import logging
from typing import Tuple

import numpy as np

from .engine import DeviceIndex, open_index


class BM25v:
    logger = logging.getLogger(__name__)

    def __init__(self, k1: float = 1.5, b: float = 0.75, device: int = 0):
        self.k1 = k1
        self.b = b
        self.dtype = np.float32
        self.device = device
        self.doc_toks = None
        self.doc_lengths = np.zeros((0,), dtype=self.dtype)
        self.avg_doc_length = 0.0
        self.num_docs = 0
        self._index: DeviceIndex | None = None

    # bm25_native.py:59-74
    def index(self, doc_toks, doc_lengths) -> None:
        """``doc_toks``: documents x tokens matrix of precomputed BM25 weights in CSC form (a
        scipy.sparse matrix, or any object with ``indptr/indices/data/shape``)."""
        if hasattr(doc_toks, "tocsc") and getattr(doc_toks, "format", "csc") != "csc":
            doc_toks = doc_toks.tocsc()
        if not all(hasattr(doc_toks, a) for a in ("indptr", "indices", "data", "shape")):
            raise ValueError("doc_toks must be a CSC sparse matrix")
        self.doc_toks = doc_toks
        self.doc_lengths = doc_lengths
        self.avg_doc_length = float(np.mean(doc_lengths)) if len(doc_lengths) else 0.0
        self.num_docs = int(doc_toks.shape[0])
        if self._index is not None:
            self._index.close()
        # one handle, or several document-range handles when the matrix has >= 2^31 postings (int64 indptr)
        self._index = open_index(doc_toks.indptr, doc_toks.indices, doc_toks.data, self.num_docs,
                                 device=self.device)

    # bm25_native.py:76-103
    def search(self, queries, top_k: int = 100) -> Tuple[np.ndarray, np.ndarray]:
        if self._index is None:
            raise ValueError("BM25v index not built. Call index() first.")
        if len(queries) == 0:
            self.logger.info("The query is empty. This will result in a zero score for all documents.")
            return np.zeros((0, 0), dtype=self.dtype), np.zeros((0, 0), dtype=self.dtype)
        return self.get_scores(queries, top_k)

    # bm25_native.py:105-127
    def get_scores(self, queries, top_k: int):
        if not isinstance(queries, np.ndarray) or queries.ndim != 2 or queries.dtype != np.int32:
            raise ValueError("The queries must be a list of list of query token IDs.")
        max_token_id = int(queries.max(initial=0))
        if max_token_id >= self._index.n_terms:
            raise ValueError(
                f"The maximum token ID in the query ({max_token_id}) is higher than the number of "
                "tokens in the index."
            )
        return self._index.search(queries, int(top_k))
