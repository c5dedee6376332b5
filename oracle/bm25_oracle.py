"""CPU oracle for the BM25 query hot path -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

This module restates, in plain numpy, the algorithm of the reference's CPU implementations of
the path (CSC posting gather -> per-document score accumulation -> top-k).  It is only ever
imported by ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl
reference`` legs of ``bench.py``.  The product package (``mojo_bm25_b200``) never imports it.

Parity status: PINNED.  The restatement is checked (tests/test_oracle_golden.py) against golden
vectors produced by running the reference's own ``bm25_native.BM25v`` and ``bm25.BM25`` in the
authoring container (tests/golden/make_golden.py -> tests/golden/*.json), including the known
answers G1-G4 recorded in SURVEY.md section 8c.

Reference locations restated here (paths relative to /root/reference):
  * bm25_native.py:76-103   BM25v.search         -> OracleBM25v.search
  * bm25_native.py:105-127  BM25v.get_scores     -> OracleBM25v._validate
  * bm25_native.py:129-158  _compute_relevance_from_scores (the hot loop)
                                                  -> scores_dense / search_csc
  * bm25_native.py:204-214  _topk (argpartition + descending sort of the k)
                                                  -> topk_partition
  * bm25.py:30-121          BM25.fit             -> OracleBM25.fit
  * bm25.py:124-145         BM25.get_scores      -> OracleBM25.get_scores
  * bm25.py:147-178         BM25.get_top_n       -> OracleBM25.get_top_n
  * animal_index_bm25/*.npy (bm25s 0.2.12 "lucene" weights) -> lucene_weight

Arithmetic notes that the CUDA path must reproduce:
  * ``csc[:, q].sum(axis=1)`` (bm25_native.py:152) is a CSC mat-vec with a ones vector: the dense
    fp32 score vector starts at +0.0 and the query's columns are added **in query-term order**,
    one fp32 add per posting; a term that occurs twice in the query is added twice.
  * -1 entries of a query row are padding (bm25_native.py:151).
  * top-k ties are resolved by whatever introselect leaves behind in the reference; the oracle
    therefore exposes the dense score vector so that checkers can be tie-aware.
"""
from __future__ import annotations

import math
from collections import Counter
from typing import Iterable, List, Sequence, Tuple

import numpy as np

__all__ = [
    "scores_dense",
    "topk_partition",
    "search_csc",
    "OracleBM25v",
    "OracleBM25",
    "lucene_weight",
    "merge_topk_lists",
    "partition_csc_by_doc_range",
    "check_topk_against_dense",
    "assert_same_topk_modulo_ties",
    "posting_bytes",
]


# --------------------------------------------------------------------------------------------
# hot loop: bm25_native.py:129-158
# --------------------------------------------------------------------------------------------
def scores_dense(indptr, indices, data, n_docs: int, query: Sequence[int]) -> np.ndarray:
    """Dense fp32 score vector of ONE query (bm25_native.py:150-152).

    ``query`` is a 1-D int array; negative ids are padding.  Columns are accumulated in query
    order with one fp32 add per posting (what scipy's csc mat-vec with a ones vector does).
    """
    out = np.zeros(int(n_docs), dtype=np.float32)
    for t in np.asarray(query).tolist():
        if t < 0:
            continue
        lo, hi = int(indptr[t]), int(indptr[t + 1])
        if hi > lo:
            # np.add.at keeps multiplicity even for a non-canonical column (duplicate rows)
            rows = indices[lo:hi]
            if hi - lo > 1 and np.any(rows[1:] <= rows[:-1]):
                np.add.at(out, rows, data[lo:hi].astype(np.float32, copy=False))
            else:
                out[rows] += data[lo:hi].astype(np.float32, copy=False)
    return out


def topk_partition(doc_scores: np.ndarray, k: int) -> Tuple[np.ndarray, np.ndarray]:
    """Top-k by partial selection then descending sort of the k (bm25_native.py:204-214).

    Raises ``ValueError`` when ``k`` exceeds the number of documents, as numpy's argpartition
    does in the reference.
    """
    n = doc_scores.shape[0]
    if k > n or k < 0:
        raise ValueError(f"kth(=-{k}) out of bounds ({n})")
    if k == 0:
        return np.zeros((0,), dtype=np.int64), np.zeros((0,), dtype=doc_scores.dtype)
    part = np.argpartition(doc_scores, n - k)[n - k:]
    vals = doc_scores[part]
    order = np.argsort(vals)[::-1]
    return part[order], vals[order]


def search_csc(indptr, indices, data, n_docs: int, queries: np.ndarray, k: int):
    """``BM25v.search`` on raw CSC arrays: ``(int32[Q,k], float32[Q,k])``."""
    queries = np.asarray(queries)
    q_n = queries.shape[0]
    top_docs = np.zeros((q_n, k), dtype=np.int32)
    top_scores = np.zeros((q_n, k), dtype=np.float32)
    for i in range(q_n):
        dense = scores_dense(indptr, indices, data, n_docs, queries[i])
        ids, vals = topk_partition(dense, k)
        top_docs[i] = ids
        top_scores[i] = vals
    return top_docs, top_scores


class OracleBM25v:
    """numpy restatement of ``bm25_native.BM25v`` (index/search) on raw CSC arrays."""

    def __init__(self, k1: float = 1.5, b: float = 0.75):
        self.k1, self.b = k1, b
        self.indptr = np.zeros((1,), np.int32)
        self.indices = np.zeros((0,), np.int32)
        self.data = np.zeros((0,), np.float32)
        self.num_docs = 0

    def index(self, indptr, indices, data, num_docs: int) -> None:
        self.indptr = np.asarray(indptr)
        self.indices = np.asarray(indices)
        self.data = np.asarray(data, dtype=np.float32)
        self.num_docs = int(num_docs)

    @property
    def num_terms(self) -> int:
        return len(self.indptr) - 1

    def _validate(self, queries) -> None:
        # bm25_native.py:108-121
        if (
            not isinstance(queries, np.ndarray)
            or queries.ndim != 2
            or queries.dtype != np.int32
        ):
            raise ValueError("The queries must be a list of list of query token IDs.")
        max_token_id = int(queries.max(initial=0))
        if max_token_id >= self.num_terms:
            raise ValueError(
                f"The maximum token ID in the query ({max_token_id}) is higher than the number "
                "of tokens in the index."
            )

    def search(self, queries, top_k: int = 100):
        # bm25_native.py:92-103
        if len(queries) == 0:
            return np.zeros((0, 0), np.float32), np.zeros((0, 0), np.float32)
        self._validate(queries)
        return search_csc(self.indptr, self.indices, self.data, self.num_docs, queries, top_k)

    def scores(self, query) -> np.ndarray:
        return scores_dense(self.indptr, self.indices, self.data, self.num_docs, query)


# --------------------------------------------------------------------------------------------
# dense BM25: bm25.py
# --------------------------------------------------------------------------------------------
class OracleBM25:
    """numpy restatement of ``bm25.BM25`` (dense docs x terms matrix)."""

    def __init__(self, k1: float = 1.5, b: float = 0.75):
        self.k1, self.b = k1, b
        self.corpus_size = 0
        self.avgdl = 0.0
        self.doc_len: List[int] = []
        self.vocabulary: List[str] = []
        self.term_to_id = {}
        self.idf = {}
        self.bm25_matrix = None

    def fit(self, corpus: Sequence[Sequence[str]]) -> None:
        # bm25.py:41-121
        self.corpus_size = len(corpus)
        if self.corpus_size == 0:
            return
        self.doc_len = [len(d) for d in corpus]
        self.avgdl = np.mean(self.doc_len)  # np.float64, as in the reference (bm25.py:60)
        self.vocabulary = sorted({t for d in corpus for t in d})
        self.term_to_id = {t: i for i, t in enumerate(self.vocabulary)}
        n_terms = len(self.vocabulary)
        if n_terms == 0:
            return
        tf = np.zeros((self.corpus_size, n_terms), dtype=np.float32)
        for i, doc in enumerate(corpus):
            for term, cnt in Counter(doc).items():
                tf[i, self.term_to_id[term]] = cnt
        df = (tf > 0).sum(axis=0)
        n = self.corpus_size
        # Lucene-style idf, bm25.py:105
        idf = [math.log((n - int(d) + 0.5) / (int(d) + 0.5) + 1) for d in df]
        self.idf = dict(zip(self.vocabulary, idf))
        dl = np.array(self.doc_len, dtype=np.float32)
        if self.avgdl == 0:
            norm = np.full_like(dl, self.k1 * (1 - self.b))
        else:
            norm = self.k1 * (1 - self.b + self.b * dl / self.avgdl)  # float64 (bm25.py:116)
        self.bm25_matrix = (tf * (self.k1 + 1)) / (tf + norm[:, None]) * np.array(
            idf, dtype=np.float32
        )[None, :]

    def get_scores(self, query: Iterable[str]) -> np.ndarray:
        # bm25.py:137-145 : OOV terms dropped, duplicates counted with multiplicity
        if self.bm25_matrix is None:
            return np.zeros(self.corpus_size)
        ids = [self.term_to_id[t] for t in query if t in self.term_to_id]
        if not ids:
            return np.zeros(self.corpus_size)
        return np.sum(self.bm25_matrix[:, ids], axis=1)

    def get_top_n(self, query, corpus, n: int = 5):
        # bm25.py:161-178
        if n <= 0:
            return []
        scores = self.get_scores(query)
        if scores.shape[0] == 0:
            return []
        k = min(n, self.corpus_size)
        top = np.argsort(scores)[::-1][:k]
        return [(scores[i], corpus[i]) for i in top]


def lucene_weight(tf, df, n_docs, dl, avgdl, k1=1.5, b=0.75):
    """bm25s 0.2.12 method="lucene" weight as stored in animal_index_bm25/data.csc.index.npy:
    ``idf * tf / (tf + k1 * (1 - b + b * dl / avgdl))`` with
    ``idf = ln((N - df + 0.5) / (df + 0.5) + 1)``."""
    idf = np.log((n_docs - df + 0.5) / (df + 0.5) + 1.0)
    return (idf * tf / (tf + k1 * (1.0 - b + b * dl / avgdl))).astype(np.float32)


# --------------------------------------------------------------------------------------------
# multi-GPU host logic mirrors -- checkers for the gloo tests
# --------------------------------------------------------------------------------------------
def partition_csc_by_doc_range(indptr, indices, data, n_docs: int, n_shards: int):
    """Row-slice a CSC matrix into ``n_shards`` contiguous document ranges.

    Returns a list of ``(indptr, indices_local, data, n_docs_local, doc_id_base)``.
    Shard ``g`` owns documents ``[g*ceil(N/n), min(N, (g+1)*ceil(N/n)))``.
    """
    indptr = np.asarray(indptr, dtype=np.int64)
    indices = np.asarray(indices)
    data = np.asarray(data)
    per = -(-int(n_docs) // n_shards)
    col_of = np.repeat(np.arange(len(indptr) - 1), np.diff(indptr))
    out = []
    for g in range(n_shards):
        lo, hi = g * per, min(int(n_docs), (g + 1) * per)
        sel = (indices >= lo) & (indices < hi)
        counts = np.bincount(col_of[sel], minlength=len(indptr) - 1)
        ptr = np.zeros(len(indptr), dtype=np.int32)
        np.cumsum(counts, out=ptr[1:])
        out.append(
            (ptr, (indices[sel] - lo).astype(np.int32), data[sel].astype(np.float32), max(hi - lo, 0), lo)
        )
    return out


def merge_topk_lists(ids: np.ndarray, scores: np.ndarray, k: int):
    """Merge ``L`` per-shard candidate lists ``ids/scores [L, Q, k_in]`` into a global top-k
    ordered by (score descending, doc id ascending).  Entries with a negative id are padding of
    shards that hold fewer than k_in documents and never win."""
    l_n, q_n, k_in = ids.shape
    flat_ids = np.transpose(ids, (1, 0, 2)).reshape(q_n, l_n * k_in)
    flat_sc = np.transpose(scores, (1, 0, 2)).reshape(q_n, l_n * k_in).copy()
    flat_sc[flat_ids < 0] = -np.inf
    out_i = np.zeros((q_n, k), np.int32)
    out_s = np.zeros((q_n, k), np.float32)
    for q in range(q_n):
        order = np.lexsort((flat_ids[q], -flat_sc[q].astype(np.float64)))[:k]
        out_i[q] = flat_ids[q][order]
        out_s[q] = flat_sc[q][order]
    return out_i, out_s


# --------------------------------------------------------------------------------------------
# tie-aware checkers
# --------------------------------------------------------------------------------------------
def check_topk_against_dense(ids, scores, dense, k: int, rtol: float = 1e-5, exact: bool = False):
    """Validate one query's top-k against the oracle's dense score vector, tie-aware.

    (1) ids are distinct and in range; (2) each reported score equals the oracle score of that
    doc; (3) the reported score list equals the k largest oracle scores in descending order.
    Together these imply a correct top-k up to permutation inside equal-score groups.
    ``exact=True`` demands bit-equality of the fp32 scores, otherwise ``rtol`` relative.
    """
    ids = np.asarray(ids)
    scores = np.asarray(scores, dtype=np.float32)
    assert ids.shape == (k,) and scores.shape == (k,), (ids.shape, scores.shape, k)
    assert len(set(ids.tolist())) == k, "duplicate doc ids in top-k"
    assert ids.min(initial=0) >= 0 and ids.max(initial=0) < dense.shape[0], "doc id out of range"
    want = np.sort(dense)[::-1][:k]
    got_by_doc = dense[ids]
    if exact:
        assert np.array_equal(scores.view(np.uint32), got_by_doc.astype(np.float32).view(np.uint32)), (
            "score != oracle score of the same doc (bitwise)"
        )
        assert np.array_equal(scores.view(np.uint32), want.astype(np.float32).view(np.uint32)), (
            "score list != k largest oracle scores (bitwise)"
        )
    else:
        tol = rtol * np.maximum(np.abs(want), 1e-30)
        assert np.all(np.abs(scores - got_by_doc) <= rtol * np.maximum(np.abs(got_by_doc), 1e-30)), (
            "score != oracle score of the same doc"
        )
        assert np.all(np.abs(scores - want) <= tol), "score list != k largest oracle scores"
    assert np.all(scores[:-1] >= scores[1:]), "scores not sorted descending"


def assert_same_topk_modulo_ties(ids, scores, ref_ids, ref_scores, rtol: float = 1e-5):
    """Compare against a reference top-k whose tie order is arbitrary (golden vectors).

    Scores must agree position-wise within ``rtol``.  Ids must agree as sets inside every group
    of (tolerance-)equal scores, except that the LAST group may be cut by the k boundary, where
    only the group sizes are compared.
    """
    ids, ref_ids = np.asarray(ids), np.asarray(ref_ids)
    scores, ref_scores = np.asarray(scores, np.float64), np.asarray(ref_scores, np.float64)
    assert ids.shape == ref_ids.shape and scores.shape == ref_scores.shape
    tol = rtol * np.maximum(np.abs(ref_scores), 1e-30)
    assert np.all(np.abs(scores - ref_scores) <= tol), (scores, ref_scores)
    k = ids.shape[-1]
    start = 0
    while start < k:
        end = start + 1
        while end < k and abs(ref_scores[end] - ref_scores[start]) <= rtol * max(abs(ref_scores[start]), 1e-30):
            end += 1
        if end < k:  # group fully inside the top-k
            assert set(ids[start:end].tolist()) == set(ref_ids[start:end].tolist()), (
                start, end, ids, ref_ids,
            )
        start = end


def posting_bytes(indptr, queries, k: int) -> int:
    """Algorithmic bytes of a batch, SURVEY.md section 8d:
    ``sum_q (8 * sum_{t in q, t >= 0} df(t) + 8 * k)``."""
    indptr = np.asarray(indptr, dtype=np.int64)
    q = np.asarray(queries)
    df = np.diff(indptr)
    valid = q >= 0
    tot = int(df[np.where(valid, q, 0)][valid].sum())
    return 8 * tot + 8 * int(k) * q.shape[0]
