"""Reader / writer of the on-disk bm25s-style CSC index:

    data.csc.index.npy     <f4 [nnz]   precomputed per-(term, doc) BM25 weights
    indices.csc.index.npy  <i4 [nnz]   document ids, column (term) major
    indptr.csc.index.npy   <i4 [V+1]   column pointers
    vocab.index.json       {token: term id}
    params.index.json      {k1, b, delta, method, idf_method, dtype, int_dtype, num_docs, version, backend}
    corpus.jsonl / corpus.mmindex.json   optional documents + their byte offsets

(the files the reference's bm25_test.py:35-38 writes under animal_index_bm25/ via bm25s 0.2.12 and
that nothing in the reference reads back except bm25s.BM25.load, bm25_test.py:42).
"""
from __future__ import annotations

import json
import os
from dataclasses import dataclass, field
from typing import Dict, List, Optional

import numpy as np

DATA, INDICES, INDPTR = "data.csc.index.npy", "indices.csc.index.npy", "indptr.csc.index.npy"
VOCAB, PARAMS = "vocab.index.json", "params.index.json"
NONOCC = "nonoccurrence_array.index.npy"  # bm25s writes it for method bm25l / bm25+ only
CORPUS, CORPUS_INDEX = "corpus.jsonl", "corpus.mmindex.json"


@dataclass
class DiskIndex:
    indptr: np.ndarray
    indices: np.ndarray
    data: np.ndarray
    vocab: Dict[str, int]
    params: dict
    corpus: Optional[object] = None  # JsonlCorpus (offset-indexed) or a list
    corpus_offsets: List[int] = field(default_factory=list)
    nonoccurrence: Optional[np.ndarray] = None  # fp32 [V] per-term score of documents without the term

    @property
    def num_docs(self) -> int:
        return int(self.params["num_docs"])

    @property
    def num_terms(self) -> int:
        return int(self.indptr.shape[0] - 1)


class JsonlCorpus:
    """The documents of ``corpus.jsonl`` fetched by byte offset: ``corpus.mmindex.json`` holds the
    offset of every line ([0, 57, 129, 192] in the bundled index), so returning the documents of
    a top-k costs one seek + one line read each instead of parsing the whole file (what bm25s'
    ``load_corpus=True, mmap=True`` does with the same two files)."""

    def __init__(self, path: str, offsets: List[int]):
        self.path = path
        self.offsets = [int(o) for o in offsets]
        self._f = None

    def __len__(self) -> int:
        return len(self.offsets)

    def _file(self):
        if self._f is None:
            self._f = open(self.path, "rb")
        return self._f

    def __getitem__(self, i):
        if isinstance(i, slice):
            return [self[j] for j in range(*i.indices(len(self)))]
        i = int(i)
        if i < 0:
            i += len(self)
        if not 0 <= i < len(self):
            raise IndexError(i)
        f = self._file()
        f.seek(self.offsets[i])
        return json.loads(f.readline())

    def __iter__(self):
        return (self[i] for i in range(len(self)))

    def close(self):
        if self._f is not None:
            self._f.close()
            self._f = None

    def __del__(self):
        self.close()


def validate(indptr, indices, data, num_docs: int) -> None:
    if indptr.ndim != 1 or indices.ndim != 1 or data.ndim != 1:
        raise ValueError("index arrays must be 1-D")
    if indptr.shape[0] < 1 or indptr[0] != 0 or indptr[-1] != indices.shape[0]:
        raise ValueError("indptr must start at 0 and end at nnz")
    if indices.shape != data.shape:
        raise ValueError("indices and data differ in length")
    if np.any(np.diff(indptr) < 0):
        raise ValueError("indptr is not monotone")
    if indices.size and (indices.min() < 0 or indices.max() >= num_docs):
        raise ValueError("document id outside [0, num_docs)")


def load_index(path: str, load_corpus: bool = False, mmap: bool = False) -> DiskIndex:
    mode = "r" if mmap else None
    data = np.load(os.path.join(path, DATA), mmap_mode=mode)
    indices = np.load(os.path.join(path, INDICES), mmap_mode=mode)
    indptr = np.load(os.path.join(path, INDPTR), mmap_mode=mode)
    with open(os.path.join(path, PARAMS)) as f:
        params = json.load(f)
    with open(os.path.join(path, VOCAB)) as f:
        vocab = json.load(f)
    if "num_docs" not in params:
        raise ValueError("params.index.json lacks num_docs")
    for name, arr in (("indptr", indptr), ("indices", indices)):
        if arr.dtype != np.int32 and arr.size and (int(arr.max()) > np.iinfo(np.int32).max or int(arr.min()) < 0):
            raise ValueError(f"{name} of {path} does not fit int32; shard the index by document range")
    indptr = np.ascontiguousarray(indptr, dtype=np.int32)
    indices = np.ascontiguousarray(indices, dtype=np.int32)
    data = np.ascontiguousarray(data, dtype=np.float32)
    validate(indptr, indices, data, int(params["num_docs"]))
    corpus, offsets = None, []
    if load_corpus and os.path.exists(os.path.join(path, CORPUS)):
        idx_path = os.path.join(path, CORPUS_INDEX)
        if os.path.exists(idx_path):
            with open(idx_path) as f:
                offsets = json.load(f)
            corpus = JsonlCorpus(os.path.join(path, CORPUS), offsets)  # documents are fetched by offset
        else:  # no offset index: parse the whole file
            corpus = []
            with open(os.path.join(path, CORPUS), "rb") as f:
                for line in f:
                    if line.strip():
                        corpus.append(json.loads(line))
    nonocc = None
    if os.path.exists(os.path.join(path, NONOCC)):
        nonocc = np.ascontiguousarray(np.load(os.path.join(path, NONOCC)), dtype=np.float32)
        if nonocc.shape != (indptr.shape[0] - 1,):
            raise ValueError("nonoccurrence_array must hold one value per term")
    return DiskIndex(indptr, indices, data, vocab, params, corpus, offsets, nonocc)


def save_index(path: str, indptr, indices, data, vocab: Dict[str, int], num_docs: int, k1: float = 1.5,
               b: float = 0.75, delta: float = 0.5, method: str = "lucene", corpus: Optional[List] = None,
               version: str = "0.2.12", nonoccurrence=None) -> None:
    os.makedirs(path, exist_ok=True)
    indptr = np.ascontiguousarray(indptr, dtype="<i4")
    indices = np.ascontiguousarray(indices, dtype="<i4")
    data = np.ascontiguousarray(data, dtype="<f4")
    validate(indptr, indices, data, num_docs)
    np.save(os.path.join(path, DATA), data, allow_pickle=False)
    np.save(os.path.join(path, INDICES), indices, allow_pickle=False)
    np.save(os.path.join(path, INDPTR), indptr, allow_pickle=False)
    if nonoccurrence is not None:
        np.save(os.path.join(path, NONOCC), np.ascontiguousarray(nonoccurrence, dtype="<f4"), allow_pickle=False)
    params = dict(k1=k1, b=b, delta=delta, method=method, idf_method=method, dtype="float32", int_dtype="int32",
                  num_docs=int(num_docs), version=version, backend="numpy")
    with open(os.path.join(path, PARAMS), "w") as f:
        json.dump(params, f, indent=4)
    with open(os.path.join(path, VOCAB), "w") as f:
        json.dump(vocab, f)
    if corpus is not None:
        offsets = []
        with open(os.path.join(path, CORPUS), "wb") as f:
            for i, doc in enumerate(corpus):
                offsets.append(f.tell())
                rec = doc if isinstance(doc, dict) else {"id": i, "text": doc}
                f.write((json.dumps(rec) + "\n").encode("utf-8"))
        with open(os.path.join(path, CORPUS_INDEX), "w") as f:
            json.dump(offsets, f)
