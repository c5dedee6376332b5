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

Variants (norm = k1*(1-b+b*dl/avgdl)):
  * ``"bm25py"``  bm25.py:105,112-121:  idf * tf*(k1+1) / (tf + norm),  idf = ln((N-df+.5)/(df+.5)+1)
  * ``"lucene"``  bm25s 0.2.12 method="lucene" (the bundled animal_index_bm25 weights, pinned):
                  idf * tf / (tf + norm), same idf
  * the other scorers bm25s' ``method`` parameter names (animal_index_bm25/params.index.json:4-5),
    after the formulas bm25s publishes -- bm25s itself is not in /root/reference, so nothing pins
    their last bit ("parity unpinned"):
      ``"robertson"``  tf/(tf+norm),                         idf = ln(max(1, (N-df+.5)/(df+.5)))
      ``"atire"``      tf*(k1+1)/(tf+norm),                  idf = ln(N/df)
      ``"bm25l"``      (k1+1)*(c+delta)/(k1+c+delta), c = tf/(1-b+b*dl/avgdl),  idf = ln((N+1)/(df+.5))
      ``"bm25+"``      tf*(k1+1)/(tf+norm) + delta,          idf = ln((N+1)/df)
    bm25l and bm25+ give a document WITHOUT the term a non-zero contribution idf*tfc(tf=0); bm25s
    keeps the matrix sparse by storing weight - that contribution and a per-term
    ``nonoccurrence_array`` that is added back per query (``nonoccurrence(...)`` below).
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


VARIANTS = ("lucene", "bm25py", "robertson", "atire", "bm25l", "bm25+")


def _idf_variant_host(df: np.ndarray, n_docs: int, variant: str) -> np.ndarray:
    """Per-term idf in float64 (math.log term by term, like bm25s' scalar idf functions)."""
    if variant in ("lucene", "bm25py"):
        return _idf_host(df, n_docs)
    out = np.zeros(df.shape[0], dtype=np.float64)
    for i, d in enumerate(df):
        d = int(d)
        if d == 0:
            continue  # a term without postings has no weight to scale
        if variant == "robertson":
            out[i] = math.log(max(1.0, (n_docs - d + 0.5) / (d + 0.5)))
        elif variant == "atire":
            out[i] = math.log(n_docs / d)
        elif variant == "bm25l":
            out[i] = math.log((n_docs + 1) / (d + 0.5))
        else:  # bm25+
            out[i] = math.log((n_docs + 1) / d)
    return out


def nonoccurrence(df, n_docs: int, variant: str, k1: float = 1.5, delta: float = 0.5) -> Optional[np.ndarray]:
    """Per-term score of a document that does NOT contain the term: float32 [V] for bm25l
    (idf*(k1+1)*delta/(k1+delta)) and bm25+ (idf*delta), None for the other variants (zero)."""
    if variant not in ("bm25l", "bm25+"):
        return None
    idf = _idf_variant_host(np.asarray(df), n_docs, variant)
    tfc0 = (k1 + 1) * delta / (k1 + delta) if variant == "bm25l" else delta
    return (idf * tfc0).astype(np.float32)


def build_csc(token_ids, doc_ptr, n_terms: Optional[int] = None, k1: float = 1.5, b: float = 0.75,
              variant: str = "lucene", device: str = "cuda", delta: float = 0.5):
    """Build the CSC weight matrix on ``device``.

    token_ids : int32 [n_tokens]  term id of every token occurrence, documents concatenated
    doc_ptr   : int64 [N+1]       token range of document d is [doc_ptr[d], doc_ptr[d+1])
    Returns ``(indptr int32 [V+1], indices int32 [nnz], data float32 [nnz], doc_len int32 [N])``
    as tensors on ``device`` (``nnz`` < 2**31: shard larger corpora by document range).
    """
    if variant not in VARIANTS:
        raise ValueError(f"variant must be one of {VARIANTS}")
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
    idf32 = torch.from_numpy(_idf_variant_host(df.cpu().numpy(), n_docs, variant).astype(np.float32)).to(dev)
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
    elif variant in ("lucene", "robertson"):
        w = idf32[term].to(torch.float64) * tf32.to(torch.float64) / (tf32.to(torch.float64) + norm[doc])
    else:
        tf64, idf64 = tf32.to(torch.float64), idf32[term].to(torch.float64)
        if variant == "atire":
            w = idf64 * (tf64 * (k1 + 1)) / (tf64 + norm[doc])
        elif variant == "bm25l":  # stored minus the non-occurrence score idf * tfc(tf = 0)
            c = tf64 / (norm[doc] / k1)
            w = idf64 * ((k1 + 1) * (c + delta) / (k1 + c + delta) - (k1 + 1) * delta / (k1 + delta))
        else:  # bm25+: (tfc + delta) - delta
            w = idf64 * ((tf64 * (k1 + 1)) / (tf64 + norm[doc]))
    return indptr.to(torch.int32), doc.to(torch.int32), w.to(torch.float32), doc_len.to(torch.int32)
