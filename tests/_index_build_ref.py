"""Test scaffolding: the index builder's computation restated in numpy on the host (what the dense
drop-in's ``BM25.fit`` does for the "bm25py" variant), used to check the device builder bit for
bit.  Lives under tests/ -- it is not part of the product package."""
import numpy as np

from mojo_bm25_b200.index_build import _idf_host


def bm25s_method_weights_float64(tf, df, n_docs, dl, avgdl, method, k1=1.5, b=0.75, delta=0.5):
    """Scalar float64 restatement of the published bm25s scorers (score of one posting and the
    non-occurrence score of its term); an independent route to the builder's numbers."""
    import math

    norm = 1 - b + b * dl / avgdl
    if method == "robertson":
        return math.log(max(1.0, (n_docs - df + 0.5) / (df + 0.5))) * tf / (k1 * norm + tf), 0.0
    if method == "atire":
        return math.log(n_docs / df) * tf * (k1 + 1) / (tf + k1 * norm), 0.0
    if method == "bm25l":
        idf, c = math.log((n_docs + 1) / (df + 0.5)), tf / norm
        non = idf * (k1 + 1) * delta / (k1 + delta)
        return idf * (k1 + 1) * (c + delta) / (k1 + c + delta) - non, non
    if method == "bm25+":
        idf = math.log((n_docs + 1) / df)
        return idf * (tf * (k1 + 1) / (k1 * norm + tf) + delta) - idf * delta, idf * delta
    raise ValueError(method)


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
