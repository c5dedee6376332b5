"""Test scaffolding: the index builder's computation restated in numpy on the host (what the dense
drop-in's ``BM25.fit`` does for the "bm25py" variant), used to check the device builder bit for
bit.  Lives under tests/ -- it is not part of the product package."""
import numpy as np

from mojo_bm25_b200.index_build import _idf_host


def build_csc_reference_numpy(token_ids, doc_ptr, n_terms: int, k1: float = 1.5, b: float = 0.75,
                              variant: str = "lucene"):
    """The same computation in numpy on the host (what mojo_bm25_b200.bm25.BM25.fit does for the
    "bm25py" variant); used to check the device builder bit for bit."""
    tok = np.asarray(token_ids, dtype=np.int64)
    ptr = np.asarray(doc_ptr, dtype=np.int64)
    n_docs = len(ptr) - 1
    doc_len = np.diff(ptr)
    if tok.size == 0 or n_docs == 0:
        return (np.zeros(n_terms + 1, np.int32), np.zeros(0, np.int32), np.zeros(0, np.float32), doc_len.astype(np.int32))
    doc_of_tok = np.repeat(np.arange(n_docs, dtype=np.int64), doc_len)
    pair, tf = np.unique(tok * n_docs + doc_of_tok, return_counts=True)
    term, doc = pair // n_docs, pair % n_docs
    df = np.bincount(term, minlength=n_terms)
    indptr = np.zeros(n_terms + 1, dtype=np.int64)
    np.cumsum(df, out=indptr[1:])
    idf32 = _idf_host(df, n_docs).astype(np.float32)
    dl = doc_len.astype(np.float32)
    avgdl = float(np.mean(doc_len))
    norm = np.full(n_docs, k1 * (1 - b)) if avgdl == 0 else k1 * (1 - b + b * dl.astype(np.float64) / avgdl)
    tf32 = tf.astype(np.float32)
    if variant == "bm25py":
        w = (tf32 * np.float32(k1 + 1)).astype(np.float64) / (tf32.astype(np.float64) + norm[doc]) * idf32[term].astype(np.float64)
    else:
        w = idf32[term].astype(np.float64) * tf32.astype(np.float64) / (tf32.astype(np.float64) + norm[doc])
    return indptr.astype(np.int32), doc.astype(np.int32), w.astype(np.float32), doc_len.astype(np.int32)
