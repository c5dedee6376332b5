"""Drop-in for the reference's dense ``bm25.BM25`` (bm25.py:6-178): ``fit`` / ``get_scores`` /
``get_top_n`` with the same arguments and return types.

``fit`` (bm25.py:30-121) stays host-side -- it is index construction, off the query hot path --
but never materialises the dense docs x terms matrix unless ``bm25_matrix`` is asked for: the
weights are produced directly as CSC columns and pinned in HBM.  ``get_scores`` (bm25.py:124-145,
column gather + row sum) and ``get_top_n`` (bm25.py:147-178, full argsort) run on the GPU.
"""
from __future__ import annotations

import math
from typing import List, Sequence

import numpy as np

from .engine import DeviceIndex


class BM25:
    def __init__(self, k1: float = 1.5, b: float = 0.75, device: int = 0):
        self.k1 = k1
        self.b = b
        self.device = device
        self.corpus_size = 0
        self.avgdl = 0
        self.doc_len: List[int] = []
        self.doc_freqs = {}
        self.idf = {}
        self.vocabulary: List[str] = []
        self.term_to_id = {}
        self._csc = None  # (indptr, indices, data64)
        self._index: DeviceIndex | None = None
        self._dense = None

    def fit(self, corpus: Sequence[Sequence[str]]) -> None:
        self.corpus_size = len(corpus)
        self.doc_len = [len(d) for d in corpus]
        self._csc, self._dense = None, None
        if self._index is not None:
            self._index.close()
            self._index = None
        if self.corpus_size == 0:
            self.avgdl, self.vocabulary, self.term_to_id, self.doc_freqs, self.idf = 0, [], {}, {}, {}
            return
        self.avgdl = np.mean(self.doc_len)
        self.vocabulary = sorted({t for d in corpus for t in d})
        self.term_to_id = {t: i for i, t in enumerate(self.vocabulary)}
        n_terms = len(self.vocabulary)
        if n_terms == 0:
            self.doc_freqs, self.idf = {}, {}
            return
        # (term, doc, tf) triples, term-major == CSC order
        flat_doc = np.repeat(np.arange(self.corpus_size), self.doc_len)
        flat_term = np.fromiter((self.term_to_id[t] for d in corpus for t in d), dtype=np.int64, count=len(flat_doc))
        pair, tf = np.unique(flat_term * self.corpus_size + flat_doc, return_counts=True)
        term, doc = pair // self.corpus_size, pair % self.corpus_size
        df = np.bincount(term, minlength=n_terms)
        n = self.corpus_size
        idf = np.array([math.log((n - int(d) + 0.5) / (int(d) + 0.5) + 1) for d in df])  # bm25.py:105
        self.doc_freqs = {self.vocabulary[j]: df[j] for j in range(n_terms)}
        self.idf = {self.vocabulary[j]: float(idf[j]) for j in range(n_terms)}
        dl = np.array(self.doc_len, dtype=np.float32)
        if self.avgdl == 0:
            norm = np.full(n, self.k1 * (1 - self.b))
        else:
            norm = self.k1 * (1 - self.b + self.b * dl / self.avgdl)  # float64, bm25.py:116
        tf32 = tf.astype(np.float32)
        weights = (tf32 * (self.k1 + 1)) / (tf32 + norm[doc]) * idf.astype(np.float32)[term]  # bm25.py:117-121
        indptr = np.zeros(n_terms + 1, dtype=np.int32)
        np.cumsum(df, out=indptr[1:])
        self._csc = (indptr, doc.astype(np.int32), weights)
        self._index = DeviceIndex(indptr, self._csc[1], weights.astype(np.float32), self.corpus_size,
                                  device=self.device)

    @property
    def bm25_matrix(self):
        """Dense docs x terms matrix (float64, as in the reference) -- built lazily, toy sizes only."""
        if self._csc is None:
            return None
        if self._dense is None:
            indptr, indices, data = self._csc
            dense = np.zeros((self.corpus_size, len(self.vocabulary)), dtype=np.float64)
            cols = np.repeat(np.arange(len(self.vocabulary)), np.diff(indptr))
            dense[indices, cols] = data
            self._dense = dense
        return self._dense

    def _query_ids(self, query):
        return [self.term_to_id[t] for t in query if t in self.term_to_id]  # OOV dropped, bm25.py:140

    def get_scores(self, query) -> np.ndarray:
        if self._index is None:
            return np.zeros(self.corpus_size)
        ids = self._query_ids(query)
        if not ids:
            return np.zeros(self.corpus_size)
        return self._index.scores_dense(np.array([ids], dtype=np.int32))[0].astype(np.float64)

    def get_top_n(self, query, corpus, n: int = 5):
        if n <= 0:
            return []
        if self._index is None or self.corpus_size == 0:
            return []
        ids = self._query_ids(query)
        k = min(n, self.corpus_size)
        q = np.array([ids if ids else [-1]], dtype=np.int32)
        top, scores = self._index.search(q, k)
        return [(np.float64(s), corpus[int(i)]) for i, s in zip(top[0], scores[0])]
