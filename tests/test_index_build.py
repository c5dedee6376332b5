"""Index builder: the torch pipeline (run here on the CPU device, on
CUDA under -m gpu) against its numpy restatement (tests/_index_build_ref.py), against the oracle's
``OracleBM25.fit`` (bm25.py:30-121) and against the reference's bundled bm25s index
(animal_index_bm25/, written by bm25_test.py:19-38)."""
import filecmp
import os

import numpy as np
import pytest
import torch

from mojo_bm25_b200 import index_build, index_io
from _index_build_ref import build_csc_reference_numpy

FOX = ["the quick brown fox jumps over the lazy dog", "a quick brown dog outpaces a lazy fox",
       "the lazy dog sleeps", "tall trees in the forest", "the forest has tall tall trees and a fox", ""]


def _random_corpus(rng, n_docs, n_terms, mean_len):
    lens = rng.poisson(mean_len, size=n_docs)
    lens[rng.integers(0, n_docs)] = 0
    zipf = 1.0 / np.arange(1, n_terms + 1)
    return [rng.choice(n_terms, size=int(l), p=zipf / zipf.sum()).tolist() for l in lens]


def _same(a, b):
    for x, y in zip(a, b):
        x = x.cpu().numpy() if isinstance(x, torch.Tensor) else np.asarray(x)
        y = y.cpu().numpy() if isinstance(y, torch.Tensor) else np.asarray(y)
        assert x.dtype == y.dtype and x.shape == y.shape
        assert np.array_equal(x.view(np.uint32) if x.dtype == np.float32 else x, y.view(np.uint32) if y.dtype == np.float32 else y)


@pytest.mark.parametrize("variant", ["lucene", "bm25py"])
def test_builder_matches_numpy_bitwise_on_cpu(variant):
    rng = np.random.default_rng(7)
    for n_docs, n_terms, mean_len in [(1, 3, 4), (50, 20, 6), (400, 300, 30)]:
        flat, ptr = index_build.flatten_corpus(_random_corpus(rng, n_docs, n_terms, mean_len))
        got = index_build.build_csc(flat, ptr, n_terms, variant=variant, device="cpu")
        want = build_csc_reference_numpy(flat, ptr, n_terms, variant=variant)
        _same(got, want)
        indptr, indices = want[0], want[1]
        assert indptr[-1] == len(indices)
        for t in range(n_terms):  # canonical CSC: rows strictly ascending inside every column
            col = indices[indptr[t]:indptr[t + 1]]
            assert np.all(col[1:] > col[:-1])


def test_bm25py_variant_equals_the_dense_dropins_fit():
    """bm25.py:30-121 semantics: vocabulary sorted, weights as in BM25.fit (host numpy)."""
    docs = [d.lower().split() for d in FOX]
    vocab = sorted({t for d in docs for t in d})
    tid = {t: i for i, t in enumerate(vocab)}
    flat, ptr = index_build.flatten_corpus([[tid[t] for t in d] for d in docs])
    got = index_build.build_csc(flat, ptr, len(vocab), variant="bm25py", device="cpu")
    # the same arithmetic as mojo_bm25_b200/bm25.py::BM25.fit, restated without a device
    want = build_csc_reference_numpy(flat, ptr, len(vocab), variant="bm25py")
    _same(got, want)
    from oracle import bm25_oracle as orc

    o = orc.OracleBM25()
    o.fit(docs)
    dense = np.zeros((len(docs), len(vocab)), dtype=np.float64)
    indptr, indices, data = (x.numpy() for x in got[:3])
    cols = np.repeat(np.arange(len(vocab)), np.diff(indptr))
    dense[indices, cols] = data
    np.testing.assert_allclose(dense, np.asarray(o.bm25_matrix), rtol=1e-6, atol=0)


def _bundled_corpus_ids(golden_dir):
    """The stemmed token ids of the reference's 4-document animal corpus, recovered from the bundled
    index itself: every posting has tf = 1 (doc lengths 4, 6, 5, 5 = the row counts)."""
    d = index_io.load_index(os.path.join(golden_dir, "animal_index_bm25"), load_corpus=True)
    cols = np.repeat(np.arange(d.num_terms), np.diff(d.indptr))
    docs = [sorted(cols[d.indices == doc].tolist()) for doc in range(d.num_docs)]
    assert [len(x) for x in docs] == [4, 6, 5, 5]
    return d, docs


def _check_bundled_reproduction(golden_dir, tmp_path, device):
    d, docs = _bundled_corpus_ids(golden_dir)
    flat, ptr = index_build.flatten_corpus(docs)
    indptr, indices, data, dl = index_build.build_csc(flat, ptr, d.num_terms, variant="lucene", device=device)
    assert np.array_equal(indptr.cpu().numpy(), d.indptr) and np.array_equal(indices.cpu().numpy(), d.indices)
    # the 20 floats bm25s 0.2.12 wrote (data.csc.index.npy); bm25s computes in float32, the builder in
    # float64 then rounds: at most 1 ulp apart
    np.testing.assert_allclose(data.cpu().numpy(), d.data, rtol=2e-7, atol=0)
    assert dl.cpu().numpy().tolist() == [4, 6, 5, 5]
    # written back in the on-disk layout with the bundled weights: the same seven files, byte for byte
    index_io.save_index(str(tmp_path), indptr.cpu().numpy(), indices.cpu().numpy(), d.data, d.vocab, d.num_docs,
                        corpus=d.corpus)
    src = os.path.join(golden_dir, "animal_index_bm25")
    for f in sorted(os.listdir(src)):
        assert filecmp.cmp(os.path.join(src, f), os.path.join(str(tmp_path), f), shallow=False), f
    return data.cpu().numpy(), d


def test_builder_reproduces_the_bundled_bm25s_index_on_cpu(golden_dir, tmp_path):
    _check_bundled_reproduction(golden_dir, tmp_path, "cpu")


def _check_against_oracle_fit(device):
    """bm25.py:30-121 through the oracle: whitespace tokens, sorted vocabulary, (k1+1) numerator."""
    from oracle import bm25_oracle as orc

    rng = np.random.default_rng(5)
    words = [f"w{i}" for i in range(60)]
    docs = [[words[j] for j in rng.choice(60, size=int(rng.integers(1, 14)), p=(1 / np.arange(1, 61)) / np.sum(1 / np.arange(1, 61)))]
            for _ in range(80)]
    o = orc.OracleBM25()
    o.fit(docs)
    vocab = sorted({t for doc in docs for t in doc})
    tid = {t: i for i, t in enumerate(vocab)}
    flat, ptr = index_build.flatten_corpus([[tid[t] for t in doc] for doc in docs])
    indptr, indices, data, _ = (x.cpu().numpy() for x in index_build.build_csc(flat, ptr, len(vocab), variant="bm25py", device=device))
    dense = np.zeros((len(docs), len(vocab)), dtype=np.float64)
    dense[indices, np.repeat(np.arange(len(vocab)), np.diff(indptr))] = data
    ref = np.asarray(o.bm25_matrix)
    assert np.array_equal(dense != 0, ref != 0)
    np.testing.assert_allclose(dense, ref, rtol=1e-6, atol=0)  # float32 storage of float64 weights


def test_builder_matches_oracle_fit_on_cpu():
    _check_against_oracle_fit("cpu")


@pytest.mark.gpu
def test_builder_on_gpu_matches_oracle_fit_and_the_bundled_index(golden_dir, tmp_path):
    _check_against_oracle_fit("cuda")
    data, d = _check_bundled_reproduction(golden_dir, tmp_path, "cuda")
    # and the built index answers the reference's known query G1 like the bundled one
    from mojo_bm25_b200.bm25s_api import BM25

    r = BM25.from_arrays(d.indptr, d.indices, data, d.num_docs, vocab=d.vocab)
    res = r.retrieve([["fish", "purr", "cat", "like"]], k=2)
    assert res.documents.tolist() == [[0, 3]]
    np.testing.assert_allclose(res.scores, [[1.5876564, 0.48158914]], rtol=1e-6)


def test_empty_inputs():
    out = index_build.build_csc(np.zeros(0, np.int32), np.zeros(1, np.int64), 5, device="cpu")
    assert out[0].tolist() == [0] * 6 and out[1].numel() == 0
    with pytest.raises(ValueError):
        index_build.build_csc(np.array([7], np.int32), np.array([0, 1]), 5, device="cpu")


@pytest.mark.gpu
def test_builder_on_gpu_and_bm25s_shaped_index_roundtrip(tmp_path):
    from mojo_bm25_b200.bm25s_api import BM25
    from oracle import bm25_oracle as orc

    rng = np.random.default_rng(11)
    corpus = _random_corpus(rng, 3000, 500, 25)
    flat, ptr = index_build.flatten_corpus(corpus)
    got = index_build.build_csc(flat, ptr, 500, variant="lucene", device="cuda")
    want = build_csc_reference_numpy(flat, ptr, 500, variant="lucene")
    _same(got, want)
    r = BM25()
    r.index(corpus, n_terms=500)
    q = rng.integers(0, 500, size=(16, 5)).astype(np.int32)
    res = r.retrieve(q, k=10)
    indptr, indices, data = want[0], want[1], want[2]
    for i in range(len(q)):
        dense = orc.scores_dense(indptr, indices, data, 3000, q[i])
        orc.check_topk_against_dense(res.documents[i], res.scores[i], dense, 10, exact=True)
    # string tokens: vocabulary by first appearance; save -> load -> same answers
    docs = [d.lower().split() for d in FOX[:5]]
    r2 = BM25()
    r2.index(docs)
    a = r2.retrieve([["lazy", "fox"]], k=3)
    r2.save(str(tmp_path / "idx"))
    b = BM25.load(str(tmp_path / "idx")).retrieve([["lazy", "fox"]], k=3)
    assert np.array_equal(a.documents, b.documents) and np.array_equal(a.scores, b.scores)


@pytest.mark.parametrize("method", ["robertson", "atire", "bm25l", "bm25+"])
def test_bm25s_scorer_variants_follow_the_published_formulas(method):
    """SURVEY.md section 8f row 4 (bm25l / bm25+ and the other ``method`` values of
    params.index.json:4-5).  bm25s is not in /root/reference, so these are checked against an
    independent scalar float64 restatement of the published formulas -- parity with bm25s itself
    is unpinned."""
    from _index_build_ref import bm25s_method_weights_float64

    rng = np.random.default_rng(3)
    corpus = _random_corpus(rng, 120, 40, 9)
    flat, ptr = index_build.flatten_corpus(corpus)
    indptr, indices, data, dl = (x.numpy() for x in index_build.build_csc(flat, ptr, 40, variant=method, device="cpu",
                                                                          k1=1.2, b=0.6, delta=0.7))
    df = np.diff(indptr)
    non = index_build.nonoccurrence(df, 120, method, k1=1.2, delta=0.7)
    assert (non is not None) == (method in ("bm25l", "bm25+"))
    avgdl = float(np.mean(dl))
    assert np.all(data >= 0) and (method == "robertson" or np.all(data > 0))
    for t in range(40):
        for p in range(indptr[t], indptr[t + 1]):
            d = indices[p]
            tf = corpus[d].count(t)
            w, n0 = bm25s_method_weights_float64(tf, int(df[t]), 120, float(dl[d]), avgdl, method, k1=1.2, b=0.6, delta=0.7)
            assert abs(data[p] - w) <= 4e-7 * max(abs(w), 1e-3), (t, d)
            if non is not None:
                assert abs(non[t] - n0) <= 4e-7 * abs(n0)


@pytest.mark.gpu
def test_bm25l_and_bm25plus_retrieve_adds_nonoccurrence_scores(tmp_path):
    """Scores of a bm25l / bm25+ index = sparse part (device top-k) + the per-query non-occurrence
    constant; ranking equals the ranking of the full dense score vector; save/load keeps the array."""
    from mojo_bm25_b200.bm25s_api import BM25

    rng = np.random.default_rng(21)
    corpus = _random_corpus(rng, 2000, 300, 20)
    for method in ("bm25l", "bm25+"):
        r = BM25(method=method, delta=0.5)
        r.index(corpus, n_terms=300)
        assert r.nonoccurrence_array is not None and r.nonoccurrence_array.shape == (300,)
        q = rng.integers(0, 300, size=(8, 4)).astype(np.int32)
        q[2, 2:] = -1
        res = r.retrieve(q, k=10)
        s = r.scores
        for i in range(len(q)):
            toks = q[i][q[i] >= 0]
            dense = np.zeros(2000, np.float32)
            for t in toks:  # sparse part in query order, like the hot loop
                sl = slice(s["indptr"][t], s["indptr"][t + 1])
                dense[s["indices"][sl]] += s["data"][sl]
            dense = dense + r.nonoccurrence_array[toks].sum()
            order = np.lexsort((np.arange(2000), -dense))[:10]
            assert np.array_equal(res.scores[i], dense[order])           # the 10 best full scores ...
            assert np.array_equal(dense[res.documents[i]], res.scores[i])  # ... and documents that have them
        r.save(str(tmp_path / method))
        again = BM25.load(str(tmp_path / method))
        assert again.method == method and np.array_equal(again.nonoccurrence_array, r.nonoccurrence_array)
        res2 = again.retrieve(q, k=10)
        assert np.array_equal(res2.documents, res.documents) and np.array_equal(res2.scores, res.scores)
