"""The CPU oracle (oracle/bm25_oracle.py) pinned against golden vectors produced by running the
reference itself (tests/golden/make_golden.py).  CPU only."""
import json
import os

import numpy as np
import pytest

from oracle import bm25_oracle as orc


def _load(golden_dir, name):
    return json.load(open(os.path.join(golden_dir, name)))


def _bits_to_f32(b):
    return np.array(b, dtype=np.uint32).view(np.float32)


def test_bundled_index_weights_follow_lucene_formula(golden_dir):
    # data = idf * tf / (tf + k1 (1 - b + b dl / avgdl)), tf = 1
    g = _load(golden_dir, "golden_bundled.json")
    indptr, indices = np.array(g["indptr"]), np.array(g["indices"])
    data = _bits_to_f32(g["data_bits"])
    dl = np.array(g["doc_lengths"], dtype=np.float64)
    df = np.diff(indptr)
    col = np.repeat(np.arange(len(df)), df)
    want = orc.lucene_weight(1.0, df[col], g["params"]["num_docs"], dl[indices], dl.mean(),
                             g["params"]["k1"], g["params"]["b"])
    np.testing.assert_allclose(data, want, rtol=2e-7)


def test_bm25v_bundled_cases(golden_dir):
    g = _load(golden_dir, "golden_bundled.json")
    indptr, indices = np.array(g["indptr"], np.int32), np.array(g["indices"], np.int32)
    data = _bits_to_f32(g["data_bits"])
    n_docs = g["params"]["num_docs"]
    model = orc.OracleBM25v()
    model.index(indptr, indices, data, n_docs)
    for case in g["cases"]:
        q = np.array(case["queries"], dtype=np.int32)
        k = case["k"]
        ids, sc = model.search(q, top_k=k)
        assert ids.dtype == np.int32 and sc.dtype == np.float32 and list(sc.shape) == case["shape"]
        ref_sc = _bits_to_f32(case["score_bits"]).reshape(sc.shape)
        ref_dense = _bits_to_f32(case["dense_bits"]).reshape(len(q), n_docs)
        assert np.array_equal(sc.view(np.uint32), ref_sc.view(np.uint32))  # bitwise
        for i in range(len(q)):
            dense = model.scores(q[i])
            assert np.array_equal(dense.view(np.uint32), ref_dense[i].view(np.uint32))
            orc.check_topk_against_dense(ids[i], sc[i], ref_dense[i], k, exact=True)
            orc.assert_same_topk_modulo_ties(ids[i], sc[i], np.array(case["ids"][i]), ref_sc[i])


def test_bm25v_known_answers_g1(golden_dir):
    # the literal known answers of SURVEY.md section 8c (G1)
    g = _load(golden_dir, "golden_bundled.json")
    model = orc.OracleBM25v()
    model.index(np.array(g["indptr"], np.int32), np.array(g["indices"], np.int32),
                _bits_to_f32(g["data_bits"]), 4)
    ids, sc = model.search(np.array([[17, 16, 2, 0]], np.int32), top_k=2)
    assert ids.tolist() == [[0, 3]]
    np.testing.assert_allclose(sc, [[1.5876564, 0.48158914]], rtol=1e-6)
    ids, sc = model.search(np.array([[19, 3, 10, -1]], np.int32), top_k=4)
    assert ids[0, 0] == 1
    np.testing.assert_allclose(sc, [[1.3254746, 0, 0, 0]], rtol=1e-6)


def test_bm25v_error_behaviour(golden_dir):
    g = _load(golden_dir, "golden_bundled.json")
    model = orc.OracleBM25v()
    model.index(np.array(g["indptr"], np.int32), np.array(g["indices"], np.int32),
                _bits_to_f32(g["data_bits"]), 4)
    assert g["errors"] == dict(token_id_out_of_range="ValueError", int64_queries="ValueError",
                               one_dim_queries="ValueError", k_gt_num_docs="ValueError")
    with pytest.raises(ValueError):
        model.search(np.array([[20]], np.int32), top_k=2)
    with pytest.raises(ValueError):
        model.search(np.array([[1]], np.int64), top_k=2)
    with pytest.raises(ValueError):
        model.search(np.array([1, 2], np.int32), top_k=2)
    with pytest.raises(ValueError):
        model.search(np.array([[1]], np.int32), top_k=5)
    ids, sc = model.search(np.zeros((0, 3), np.int32), top_k=3)
    assert list(ids.shape) == g["empty"]["ids_shape"] and str(ids.dtype) == g["empty"]["ids_dtype"]
    assert list(sc.shape) == g["empty"]["scores_shape"] and str(sc.dtype) == g["empty"]["scores_dtype"]


def test_selfcheck_g4(golden_dir):
    g = _load(golden_dir, "golden_selfcheck.json")
    model = orc.OracleBM25v()
    model.index(np.array(g["indptr"], np.int32), np.array(g["indices"], np.int32),
                np.array(g["data"], np.float32), 2)
    ids, sc = model.search(np.array(g["query"], np.int32), top_k=g["k"])
    assert ids.tolist() == g["ids"] == [[1]]
    assert sc.tolist() == [[6.0]]


def test_bm25v_random_cases_bitwise(golden_dir):
    z = np.load(os.path.join(golden_dir, "golden_random.npz"))
    for name in z["names"].tolist():
        n_docs, n_terms, k = z[f"{name}_meta"].tolist()
        model = orc.OracleBM25v()
        model.index(z[f"{name}_indptr"], z[f"{name}_indices"], z[f"{name}_data"], n_docs)
        q = z[f"{name}_queries"]
        ids, sc = model.search(q, top_k=k)
        assert np.array_equal(sc.view(np.uint32), z[f"{name}_scores"].view(np.uint32)), name
        for i in range(len(q)):
            dense = model.scores(q[i])
            assert np.array_equal(dense.view(np.uint32), z[f"{name}_dense"][i].view(np.uint32)), name
            orc.check_topk_against_dense(ids[i], sc[i], dense, k, exact=True)
            orc.assert_same_topk_modulo_ties(ids[i], sc[i], z[f"{name}_ids"][i], z[f"{name}_scores"][i])


@pytest.mark.parametrize("corpus", ["fox", "animal"])
def test_dense_bm25_matches_reference(golden_dir, corpus):
    g = _load(golden_dir, "golden_dense.json")["corpora"][corpus]
    docs = [d.lower().split() for d in g["docs"]]
    model = orc.OracleBM25()
    model.fit(docs)
    assert model.vocabulary == g["vocabulary"]
    assert str(model.bm25_matrix.dtype) == g["matrix_dtype"]
    np.testing.assert_allclose(model.bm25_matrix, np.array(g["matrix"]), rtol=1e-12, atol=0)
    for qe in g["queries"]:
        toks = qe["query"].lower().split()
        np.testing.assert_allclose(model.get_scores(toks), np.array(qe["scores"]), rtol=1e-12, atol=0)
        for n, want in qe["top_n"].items():
            got = model.get_top_n(toks, docs, n=int(n))
            assert len(got) == len(want["scores"])
            np.testing.assert_allclose([s for s, _ in got], want["scores"], rtol=1e-12, atol=0)
            # documents agree wherever the score is untied
            ws = want["scores"]
            for i, (s, d) in enumerate(got):
                tied = (i > 0 and ws[i - 1] == ws[i]) or (i + 1 < len(ws) and ws[i + 1] == ws[i])
                if not tied and i + 1 < len(ws):
                    assert " ".join(d) == want["docs"][i]


def test_known_answers_g2_fox(golden_dir):
    # "quick brown fox" -> ids [2, {6,0}, 4, 10]
    model = orc.OracleBM25()
    g = _load(golden_dir, "golden_dense.json")["corpora"]["fox"]
    docs = [d.lower().split() for d in g["docs"]]
    model.fit(docs)
    s = model.get_scores("quick brown fox".split())
    order = np.argsort(-s, kind="stable")[:5]
    assert order.tolist() == [2, 0, 6, 4, 10]
    np.testing.assert_allclose(s[order], [1.708471, 1.615151, 1.615151, 1.2205684, 1.1655283], rtol=1e-6)


def test_partition_and_merge_reproduce_global_topk():
    rng = np.random.default_rng(3)
    import scipy.sparse as sp

    m = sp.random(500, 40, density=0.2, format="csc", dtype=np.float32, random_state=np.random.RandomState(1))
    m.sort_indices()
    q = rng.integers(0, 40, size=(9, 5)).astype(np.int32)
    k = 20
    parts = orc.partition_csc_by_doc_range(m.indptr, m.indices, m.data, 500, 3)
    assert sum(p[3] for p in parts) == 500
    ids = np.zeros((3, 9, k), np.int32)
    sc = np.zeros((3, 9, k), np.float32)
    for g, (ptr, idx, dat, nd, base) in enumerate(parts):
        i, s = orc.search_csc(ptr, idx, dat, nd, q, k)
        ids[g], sc[g] = i + base, s
    mi, ms = orc.merge_topk_lists(ids, sc, k)
    for r in range(9):
        dense = orc.scores_dense(m.indptr, m.indices, m.data, 500, q[r])
        orc.check_topk_against_dense(mi[r], ms[r], dense, k, exact=True)


def test_posting_bytes_formula():
    indptr = np.array([0, 3, 3, 10])
    q = np.array([[0, 2, -1], [2, 2, 1]], np.int32)
    assert orc.posting_bytes(indptr, q, 5) == 8 * (3 + 7 + 7 + 7 + 0) + 8 * 5 * 2


def test_c_oracle_matches_numpy_oracle(golden_dir):
    from oracle import c_oracle

    z = np.load(os.path.join(golden_dir, "golden_random.npz"))
    for name in z["names"].tolist():
        n_docs, n_terms, k = z[f"{name}_meta"].tolist()
        q = z[f"{name}_queries"]
        ids, sc = c_oracle.search(z[f"{name}_indptr"], z[f"{name}_indices"], z[f"{name}_data"], n_docs, q, k)
        assert np.array_equal(sc.view(np.uint32), z[f"{name}_scores"].view(np.uint32)), name
        for i in range(len(q)):
            dense = c_oracle.scores_dense(z[f"{name}_indptr"], z[f"{name}_indices"], z[f"{name}_data"], n_docs, q[i])
            assert np.array_equal(dense.view(np.uint32), z[f"{name}_dense"][i].view(np.uint32))
            orc.check_topk_against_dense(ids[i], sc[i], dense, k, exact=True)
    with pytest.raises(ValueError):
        c_oracle.search(z["r1_indptr"], z["r1_indices"], z["r1_data"], 7, z["r1_queries"], 8)
