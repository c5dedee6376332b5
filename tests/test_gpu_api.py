"""The reference-facing Python API (BM25v, BM25, bm25s-shaped load/retrieve, gpu_execute_query)
on the GPU, against golden vectors produced by the reference.  Run with -m gpu."""
import json
import os

import numpy as np
import pytest

from oracle import bm25_oracle as orc

pytestmark = pytest.mark.gpu
RTOL = 1e-5  # BASELINE.json north_star: scores within 1e-5 relative


def _bits_to_f32(b):
    return np.array(b, dtype=np.uint32).view(np.float32)


def test_bm25v_dropin_on_bundled_index(golden_dir):
    import scipy.sparse as sp

    from mojo_bm25_b200.bm25_native import BM25v

    g = json.load(open(os.path.join(golden_dir, "golden_bundled.json")))
    m = sp.csc_matrix((_bits_to_f32(g["data_bits"]), np.array(g["indices"], np.int32), np.array(g["indptr"], np.int32)),
                      shape=(4, 20))
    model = BM25v()
    with pytest.raises(ValueError):
        model.search(np.array([[1]], np.int32), top_k=1)  # not indexed yet
    model.index(m, np.array(g["doc_lengths"], np.int32))
    assert model.num_docs == 4 and abs(model.avg_doc_length - 5.0) < 1e-12
    for case in g["cases"]:
        ids, sc = model.search(np.array(case["queries"], np.int32), top_k=case["k"])
        ref_sc = _bits_to_f32(case["score_bits"]).reshape(sc.shape)
        assert ids.dtype == np.int32 and sc.dtype == np.float32
        assert np.array_equal(sc.view(np.uint32), ref_sc.view(np.uint32))
        for i in range(len(ids)):
            orc.assert_same_topk_modulo_ties(ids[i], sc[i], np.array(case["ids"][i]), ref_sc[i], rtol=RTOL)
    # error behaviour recorded from the reference
    assert set(g["errors"].values()) == {"ValueError"}
    for bad, k in [(np.array([[20]], np.int32), 2), (np.array([[1]], np.int64), 2), (np.array([1, 2], np.int32), 2),
                   (np.array([[1]], np.int32), 5), ([[1, 2]], 2)]:
        with pytest.raises(ValueError):
            model.search(bad, top_k=k)
    e_ids, e_sc = model.search(np.zeros((0, 3), np.int32), top_k=3)
    assert list(e_ids.shape) == g["empty"]["ids_shape"] and str(e_ids.dtype) == g["empty"]["ids_dtype"]
    assert list(e_sc.shape) == g["empty"]["scores_shape"] and str(e_sc.dtype) == g["empty"]["scores_dtype"]
    # csr input is converted
    model.index(m.tocsr(), np.array(g["doc_lengths"], np.int32))
    ids, sc = model.search(np.array([[17, 16, 2, 0]], np.int32), top_k=2)
    assert ids.tolist() == [[0, 3]]


@pytest.mark.parametrize("corpus", ["fox", "animal"])
def test_dense_bm25_dropin(golden_dir, corpus):
    from mojo_bm25_b200.bm25 import BM25

    g = json.load(open(os.path.join(golden_dir, "golden_dense.json")))["corpora"][corpus]
    docs = [d.lower().split() for d in g["docs"]]
    model = BM25()
    model.fit(docs)
    assert model.vocabulary == g["vocabulary"] and model.corpus_size == len(docs)
    assert abs(float(model.avgdl) - g["avgdl"]) < 1e-12
    assert str(model.bm25_matrix.dtype) == g["matrix_dtype"]
    np.testing.assert_allclose(model.bm25_matrix, np.array(g["matrix"]), rtol=1e-12, atol=0)
    for qe in g["queries"]:
        toks = qe["query"].lower().split()
        scores = model.get_scores(toks)
        assert scores.shape == (len(docs),)
        np.testing.assert_allclose(scores, np.array(qe["scores"]), rtol=RTOL, atol=0)
        for n, want in qe["top_n"].items():
            got = model.get_top_n(toks, docs, n=int(n))
            assert len(got) == len(want["scores"])
            np.testing.assert_allclose([s for s, _ in got], want["scores"], rtol=RTOL, atol=0)
            # the document returned at rank i must be one whose REFERENCE score equals the reference's
            # rank-i score (ties between equal-score documents are resolved arbitrarily by argsort)
            ref_scores = np.array(qe["scores"])
            for i, (s, d) in enumerate(got):
                cands = [j for j, doc in enumerate(docs) if doc == d]
                assert any(abs(ref_scores[j] - want["scores"][i]) <= RTOL * max(abs(want["scores"][i]), 1e-30)
                           for j in cands), (qe["query"], n, i)
            assert len({id(d) for _, d in got}) == len(got)  # distinct documents
    assert model.get_top_n(["fox"], docs, n=0) == [] and model.get_top_n(["fox"], docs, n=-3) == []
    empty = BM25()
    empty.fit([])
    assert empty.get_top_n(["x"], [], n=3) == [] and empty.get_scores(["x"]).shape == (0,)


def test_known_answers_g2_g3(golden_dir):
    from mojo_bm25_b200.bm25 import BM25

    g = json.load(open(os.path.join(golden_dir, "golden_dense.json")))["corpora"]
    fox = [d.lower().split() for d in g["fox"]["docs"]]
    m = BM25()
    m.fit(fox)
    top = m.get_top_n("quick brown fox".split(), fox, n=5)
    np.testing.assert_allclose([s for s, _ in top], [1.708471, 1.615151, 1.615151, 1.2205684, 1.1655283], rtol=1e-6)
    assert " ".join(top[0][1]) == g["fox"]["docs"][2].lower()
    animal = [d.lower().split() for d in g["animal"]["docs"]]
    m = BM25()
    m.fit(animal)
    top = m.get_top_n("does the fish purr like a cat?".split(), animal, n=10)
    assert len(top) == 4  # k clipped to the corpus size (bm25.py:172)
    np.testing.assert_allclose([s for s, _ in top], [1.416218, 1.297955, 1.252951, 0.155514], rtol=1e-5)
    assert [animal.index(d) for _, d in top] == [0, 3, 1, 2]


def test_bm25s_shaped_load_and_retrieve(golden_dir, tmp_path):
    from mojo_bm25_b200.bm25s_api import BM25

    r = BM25.load(os.path.join(golden_dir, "animal_index_bm25"), load_corpus=True)
    # bm25_test.py:23-28: "does the fish purr like a cat?" -> stems fish, purr, like, cat
    res, scores = r.retrieve([["fish", "purr", "like", "cat"]], k=2)
    assert res.shape == (1, 2) and scores.shape == (1, 2)
    assert res[0, 0]["text"].startswith("a cat") and res[0, 1]["text"].startswith("a fish")
    np.testing.assert_allclose(scores, [[1.5876564, 0.48158914]], rtol=1e-6)
    r2 = BM25.load(os.path.join(golden_dir, "animal_index_bm25"))
    ids, scores = r2.retrieve([["fish", "purr", "unknownword", ""], ["dog"]], k=4)
    assert ids.dtype == np.int32 and ids[0, :2].tolist() == [0, 3] and ids[1, 0] == 1
    ids3, _ = r2.retrieve(np.array([[17, 16, -1]], np.int32), k=1)
    assert ids3.tolist() == [[0]]
    with pytest.raises(ValueError):
        r2.retrieve([["fish"]], k=5)
    r2.save(str(tmp_path), corpus=r.corpus)
    r3 = BM25.load(str(tmp_path), load_corpus=True)
    assert r3.retrieve([["bird"]], k=1).documents[0, 0]["text"].startswith("a bird")


def test_gpu_execute_query_dropin(golden_dir):
    from mojo_bm25_b200.gpu_bm25.common import gpu_execute_query

    g = json.load(open(os.path.join(golden_dir, "golden_dense.json")))["corpora"]["fox"]
    matrix = np.array(g["matrix"]).astype(np.float32)  # main.py:244
    vocab = g["vocabulary"]
    for qe in g["queries"]:
        toks = [t for t in qe["query"].lower().split() if t in vocab]
        if not toks:
            continue
        qv = np.array([vocab.index(t) for t in toks], dtype=np.int32)
        idx, weight = gpu_execute_query(matrix, qv, None, None)
        assert tuple(idx.shape) == (1, 1) and tuple(weight.shape) == (1, 1)
        want = np.array(qe["scores"])
        assert abs(weight.item() - want.max()) <= RTOL * want.max()
        assert abs(want[idx.item()] - want.max()) <= RTOL * want.max()
