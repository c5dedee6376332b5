"""GPU index builder: tokenised corpus -> CSC BM25 weight matrix in HBM.

Computes what the reference's ``BM25.fit`` (bm25.py:30-121) and bm25s' ``index`` compute -- term
frequencies, document frequencies, IDF, length normalisation, per-(term, doc) weights -- but on
the device and directly in the CSC layout the query path consumes (terms = columns, documents =
rows, column-major, rows ascending inside a column), without a dense docs x terms matrix and
without a host pass over the postings.  The result can be handed to ``DeviceIndex.from_torch`` as
is, or written with ``index_io.save_index`` in the on-disk bm25s format.

The sort / run-length / histogram steps are torch (CUB) calls -- this is index construction, off
the query hot path; all arithmetic is IEEE float64/float32 elementwise in the order the reference
uses, so the weights are bit-identical to the CPU formulas (tests/test_index_build.py).

Variants:
  * ``"bm25py"``  bm25.py:105,112-121:  idf * tf*(k1+1) / (tf + k1*(1-b+b*dl/avgdl)),
                  idf = ln((N-df+.5)/(df+.5)+1)
  * ``"lucene"``  bm25s 0.2.12 method="lucene" (the bundled animal_index_bm25 weights):
                  idf * tf / (tf + k1*(1-b+b*dl/avgdl)), same idf
"""
from __future__ import annotations

import math
from typing import Optional, Sequence, Tuple

import numpy as np
import torch


def flatten_corpus(corpus_ids: Sequence[Sequence[int]]) -> Tuple[np.ndarray, np.ndarray]:
    """list of per-document token-id lists -> (token_ids int32 [n_tokens], doc_ptr int64 [N+1])."""
    lens = np.fromiter((len(d) for d in corpus_ids), dtype=np.int64, count=len(corpus_ids))
    doc_ptr = np.zeros(len(corpus_ids) + 1, dtype=np.int64)
    np.cumsum(lens, out=doc_ptr[1:])
    flat = np.fromiter((t for d in corpus_ids for t in d), dtype=np.int32, count=int(doc_ptr[-1]))
    return flat, doc_ptr


def _idf_host(df: np.ndarray, n_docs: int) -> np.ndarray:
    """ln((N - df + 0.5) / (df + 0.5) + 1) in float64, term by term with math.log exactly like
    bm25.py:105 (numpy's vectorised log may differ from libm in the last place)."""
    if df.shape[0] <= 200_000:
        return np.array([math.log((n_docs - int(d) + 0.5) / (int(d) + 0.5) + 1) for d in df], dtype=np.float64)
    dfd = df.astype(np.float64)
    return np.log((n_docs - dfd + 0.5) / (dfd + 0.5) + 1.0)


def build_csc(token_ids, doc_ptr, n_terms: Optional[int] = None, k1: float = 1.5, b: float = 0.75,
              variant: str = "lucene", device: str = "cuda"):
    """Build the CSC weight matrix on ``device``.

    token_ids : int32 [n_tokens]  term id of every token occurrence, documents concatenated
    doc_ptr   : int64 [N+1]       token range of document d is [doc_ptr[d], doc_ptr[d+1])
    Returns ``(indptr int32 [V+1], indices int32 [nnz], data float32 [nnz], doc_len int32 [N])``
    as tensors on ``device`` (``nnz`` < 2**31: shard larger corpora by document range).
    """
    if variant not in ("lucene", "bm25py"):
        raise ValueError("variant must be 'lucene' or 'bm25py'")
    dev = torch.device(device)
    tok = torch.as_tensor(token_ids).to(dev, dtype=torch.int64)
    ptr = torch.as_tensor(doc_ptr).to(dev, dtype=torch.int64)
    n_docs = int(ptr.numel() - 1)
    if n_docs < 0 or (ptr.numel() and int(ptr[-1]) != tok.numel()):
        raise ValueError("doc_ptr must have N+1 entries ending at len(token_ids)")
    if n_terms is None:
        n_terms = int(tok.max().item()) + 1 if tok.numel() else 0
    if tok.numel() and (int(tok.min()) < 0 or int(tok.max()) >= n_terms):
        raise ValueError("token id outside [0, n_terms)")
    doc_len = (ptr[1:] - ptr[:-1])
    if tok.numel() == 0 or n_docs == 0:
        z = torch.zeros
        return (z(n_terms + 1, dtype=torch.int32, device=dev), z(0, dtype=torch.int32, device=dev),
                z(0, dtype=torch.float32, device=dev), doc_len.to(torch.int32))
    # (term, doc) pairs, term-major: one radix sort of a 64-bit key, then run lengths = tf
    doc_of_tok = torch.repeat_interleave(torch.arange(n_docs, device=dev, dtype=torch.int64), doc_len)
    key, _ = torch.sort(tok * n_docs + doc_of_tok)
    pair, tf = torch.unique_consecutive(key, return_counts=True)
    if pair.numel() >= 2 ** 31:
        raise ValueError("more than 2**31 postings: shard the corpus by document range")
    term = torch.div(pair, n_docs, rounding_mode="floor")
    doc = pair - term * n_docs
    df = torch.bincount(term, minlength=n_terms)
    indptr = torch.zeros(n_terms + 1, dtype=torch.int64, device=dev)
    torch.cumsum(df, 0, out=indptr[1:])
    # per-term idf (host, float64, then float32 like bm25.py:118), per-document length norm (float64)
    idf32 = torch.from_numpy(_idf_host(df.cpu().numpy(), n_docs).astype(np.float32)).to(dev)
    dl = doc_len.to(torch.float32)  # bm25.py keeps doc_len as float32 for the norm
    avgdl = float(np.mean(doc_len.cpu().numpy())) if n_docs else 0.0
    if avgdl == 0:
        norm = torch.full((n_docs,), k1 * (1 - b), dtype=torch.float64, device=dev)
    else:
        norm = k1 * (1 - b + b * dl.to(torch.float64) / avgdl)
    tf32 = tf.to(torch.float32)
    if variant == "bm25py":
        # (tf32 * (k1+1)) in float32, then float64 for the division and the idf product (numpy promotion)
        num = (tf32 * np.float32(k1 + 1)).to(torch.float64)
        w = num / (tf32.to(torch.float64) + norm[doc]) * idf32[term].to(torch.float64)
    else:
        w = idf32[term].to(torch.float64) * tf32.to(torch.float64) / (tf32.to(torch.float64) + norm[doc])
    return indptr.to(torch.int32), doc.to(torch.int32), w.to(torch.float32), doc_len.to(torch.int32)
