"""bm25s-shaped retrieval surface: ``BM25.load(dir, load_corpus=...)`` + ``retrieve(tokens, k)``
(the calls the reference makes at bm25_test.py:28 and :42 against third-party bm25s 0.2.12).

Only the query side is provided: the on-disk index is loaded into HBM and ``retrieve`` runs the
hot path in libbm25_b200.so.  Tokenisation/stemming (bm25s.tokenize + PyStemmer) is third-party
and out of scope: ``retrieve`` takes already-tokenised queries -- lists of token strings (mapped
through vocab.index.json, unknown tokens dropped like bm25.py:140), lists/arrays of term ids, or a
bm25s ``Tokenized``-like object with ``.ids`` and ``.vocab``.
"""
from __future__ import annotations

from collections import namedtuple
from typing import Iterable, List, Optional, Sequence

import numpy as np

from . import index_io
from .engine import DeviceIndex

Results = namedtuple("Results", ["documents", "scores"])


class BM25:
    def __init__(self, k1: float = 1.5, b: float = 0.75, delta: float = 0.5, method: str = "lucene", device: int = 0):
        self.k1, self.b, self.delta, self.method = k1, b, delta, method
        self.device = device
        self.vocab_dict = {}
        self.corpus = None
        self.scores = None  # {"data", "indices", "indptr", "num_docs"} like bm25s
        self.nonoccurrence_array = None  # fp32 [V] for method bm25l / bm25+ (bm25s' attribute of the same name)
        self._index: Optional[DeviceIndex] = None

    # --- index side ---------------------------------------------------------------------------
    @classmethod
    def load(cls, save_dir: str, load_corpus: bool = False, mmap: bool = False, device: int = 0) -> "BM25":
        disk = index_io.load_index(save_dir, load_corpus=load_corpus, mmap=mmap)
        p = disk.params
        self = cls(k1=p.get("k1", 1.5), b=p.get("b", 0.75), delta=p.get("delta", 0.5),
                   method=p.get("method", "lucene"), device=device)
        self._attach(disk.indptr, disk.indices, disk.data, disk.num_docs, disk.vocab, disk.corpus)
        self.nonoccurrence_array = disk.nonoccurrence
        return self

    @classmethod
    def from_arrays(cls, indptr, indices, data, num_docs: int, vocab=None, corpus=None, device: int = 0) -> "BM25":
        self = cls(device=device)
        self._attach(np.asarray(indptr), np.asarray(indices), np.asarray(data), num_docs, vocab or {}, corpus)
        return self

    def _attach(self, indptr, indices, data, num_docs, vocab, corpus):
        self.vocab_dict = dict(vocab)
        self.corpus = corpus
        self.scores = {"data": data, "indices": indices, "indptr": indptr, "num_docs": int(num_docs)}
        if self._index is not None:
            self._index.close()
        self._index = DeviceIndex(indptr, indices, data, int(num_docs), device=self.device)

    def index(self, corpus_tokens, n_terms: Optional[int] = None) -> None:
        """bm25s-shaped ``index`` (the reference calls it at bm25_test.py:19-20): build the CSC weight
        matrix ON THE GPU from a tokenised corpus and pin it in HBM.  ``corpus_tokens``: a list of
        documents, each a list of token strings (the vocabulary is assigned in order of first
        appearance) or of term ids; or a bm25s ``Tokenized``-like object with ``.ids`` / ``.vocab``."""
        import torch

        from . import index_build

        vocab = None
        if hasattr(corpus_tokens, "ids") and hasattr(corpus_tokens, "vocab"):
            vocab, corpus_tokens = dict(corpus_tokens.vocab), corpus_tokens.ids
        docs = [list(d) for d in corpus_tokens]
        if any(d and isinstance(d[0], str) for d in docs):
            vocab = {}
            docs = [[vocab.setdefault(t, len(vocab)) for t in d] for d in docs]
        if n_terms is None:
            n_terms = len(vocab) if vocab is not None else (max((max(d) for d in docs if d), default=-1) + 1)
        flat, doc_ptr = index_build.flatten_corpus(docs)
        variant = self.method if self.method in index_build.VARIANTS else "lucene"
        indptr, indices, data, _ = index_build.build_csc(flat, doc_ptr, n_terms, k1=self.k1, b=self.b, variant=variant,
                                                         device=f"cuda:{self.device}", delta=self.delta)
        torch.cuda.synchronize(self.device)
        df = (indptr[1:] - indptr[:-1]).cpu().numpy()
        self.nonoccurrence_array = index_build.nonoccurrence(df, len(docs), variant, k1=self.k1, delta=self.delta)
        self.vocab_dict = dict(vocab) if vocab is not None else {}
        self.corpus = None
        self.scores = {"data": data.cpu().numpy(), "indices": indices.cpu().numpy(), "indptr": indptr.cpu().numpy(),
                       "num_docs": len(docs)}
        if self._index is not None:
            self._index.close()
        self._index = DeviceIndex.from_torch(indptr, indices, data, len(docs), borrow=True)

    def save(self, save_dir: str, corpus=None) -> None:
        if self.scores is None:
            raise ValueError("nothing to save: no index loaded")
        s = self.scores
        index_io.save_index(save_dir, s["indptr"], s["indices"], s["data"], self.vocab_dict, s["num_docs"],
                            k1=self.k1, b=self.b, delta=self.delta, method=self.method,
                            corpus=corpus if corpus is not None else self.corpus,
                            nonoccurrence=self.nonoccurrence_array)

    # --- query side ---------------------------------------------------------------------------
    def get_tokens_ids(self, tokens: Iterable[str]) -> List[int]:
        n_terms = self._index.n_terms if self._index is not None else 0
        out = []
        for t in tokens:
            i = self.vocab_dict.get(t)
            if i is not None and 0 <= i < n_terms:  # vocab may hold ids without a column ("" -> V)
                out.append(int(i))
        return out

    def _to_id_matrix(self, query_tokens) -> np.ndarray:
        if hasattr(query_tokens, "ids") and hasattr(query_tokens, "vocab"):  # bm25s Tokenized
            rev = {i: t for t, i in query_tokens.vocab.items()}
            query_tokens = [[rev[i] for i in row if i in rev] for row in query_tokens.ids]
        if isinstance(query_tokens, np.ndarray) and query_tokens.ndim == 2:
            return np.ascontiguousarray(query_tokens, dtype=np.int32)
        rows: List[Sequence] = list(query_tokens)
        if rows and not isinstance(rows[0], (list, tuple, np.ndarray)):
            rows = [rows]  # a single tokenised query
        id_rows = []
        for row in rows:
            row = list(row)
            if row and isinstance(row[0], str):
                id_rows.append(self.get_tokens_ids(row))
            else:
                id_rows.append([int(x) for x in row])
        width = max(1, max((len(r) for r in id_rows), default=1))
        mat = np.full((len(id_rows), width), -1, dtype=np.int32)
        for i, r in enumerate(id_rows):
            mat[i, : len(r)] = r
        return mat

    def retrieve(self, query_tokens, corpus=None, k: int = 10, return_as: str = "tuple"):
        if self._index is None:
            raise ValueError("no index: call BM25.load() first")
        q = self._to_id_matrix(query_tokens)
        if k > self._index.n_docs:
            raise ValueError(f"k of {k} is larger than the number of available scores, which is "
                             f"{self._index.n_docs} (corpus size should be larger than top-k).")
        if q.size and int(q.max(initial=-1)) >= self._index.n_terms:
            raise ValueError("query token id is outside the index vocabulary")
        ids, scores = self._index.search(q, int(k))
        if self.nonoccurrence_array is not None:
            # bm25l / bm25+: every document also gets the non-occurrence score of every query token
            # (bm25s adds ``nonoccurrence_array[query_tokens_ids].sum()`` to the whole score vector:
            # a per-query constant, so the ranking found on the device is unchanged)
            for i in range(q.shape[0]):
                row = q[i][q[i] >= 0]
                if row.size:
                    scores[i] += self.nonoccurrence_array[row].sum()
        docs = ids
        corpus = corpus if corpus is not None else self.corpus
        if corpus is not None:
            docs = np.array([[corpus[int(i)] for i in row] for row in ids], dtype=object)
        if return_as == "documents":
            return docs
        return Results(documents=docs, scores=scores)
