"""Generate the golden vectors under tests/golden/ by RUNNING THE REFERENCE ITSELF.

Run in the authoring container only (needs /root/reference, numpy, scipy):

    PYTHONBREAKPOINT=0 python tests/golden/make_golden.py

It imports the reference's ``bm25_native.BM25v`` and ``bm25.BM25`` by path (nothing is copied),
feeds them fixed inputs and stores inputs + outputs:

  * golden_bundled.json  -- G1: bundled animal_index_bm25 CSC arrays, queries, BM25v outputs
  * golden_dense.json    -- G2/G3: fox + animal corpora through bm25.BM25 (scores, top-n order)
  * golden_selfcheck.json-- G4: bm25_native.py __main__ self check
  * animal_index_bm25/   -- the bundled bm25s on-disk index itself (7 data files, 1.4 KB)
  * golden_random.npz    -- R*: seeded random CSC matrices through BM25v.search plus the
                            reference's dense per-query score vectors (bitwise fp32)

The GPU box has no /root/reference; tests read only the files written here.
"""
import json
import os
import sys

import numpy as np
import scipy.sparse as sp

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
os.environ.setdefault("PYTHONBREAKPOINT", "0")
sys.dont_write_bytecode = True
sys.path.insert(0, REF)

import bm25 as ref_bm25  # noqa: E402
import bm25_native as ref_native  # noqa: E402


def f32_list(a):
    return [float(np.float32(x)) for x in np.asarray(a).ravel()]


def bits(a):
    return [int(x) for x in np.asarray(a, dtype=np.float32).ravel().view(np.uint32)]


def run_bm25v(indptr, indices, data, n_docs, queries, k):
    m = sp.csc_matrix((data, indices, indptr), shape=(n_docs, len(indptr) - 1))
    model = ref_native.BM25v()
    model.index(m, np.ones(n_docs, dtype=np.int32))
    ids, sc = model.search(queries, top_k=k)
    dense = np.stack(
        [np.asarray(m[:, row[row >= 0]].sum(axis=1)).ravel() for row in queries]
    ).astype(np.float32)
    return ids, sc, dense


def expect_error(fn):
    try:
        fn()
    except Exception as e:  # noqa: BLE001
        return type(e).__name__
    return None


def golden_bundled():
    d = os.path.join(REF, "animal_index_bm25")
    data = np.load(os.path.join(d, "data.csc.index.npy"))
    indices = np.load(os.path.join(d, "indices.csc.index.npy"))
    indptr = np.load(os.path.join(d, "indptr.csc.index.npy"))
    params = json.load(open(os.path.join(d, "params.index.json")))
    vocab = json.load(open(os.path.join(d, "vocab.index.json")))
    n_docs = params["num_docs"]
    cases = []
    for queries, k in [
        ([[17, 16, 2, 0]], 2),
        ([[17, 16, -1, -1], [19, 3, 10, -1]], 4),
        ([[2, 2, 16]], 2),
        ([[17, 16, 2, 0]], 4),
        ([[-1, -1]], 3),
        ([[0], [1], [2], [3], [4], [5], [6], [7], [8], [9], [10], [11], [12], [13], [14], [15], [16], [17], [18], [19]], 1),
    ]:
        q = np.array(queries, dtype=np.int32)
        ids, sc, dense = run_bm25v(indptr, indices, data, n_docs, q, k)
        cases.append(
            dict(queries=queries, k=k, ids=ids.tolist(), scores=f32_list(sc), score_bits=bits(sc),
                 dense_bits=bits(dense), shape=list(sc.shape))
        )
    m = sp.csc_matrix((data, indices, indptr), shape=(n_docs, len(indptr) - 1))
    model = ref_native.BM25v()
    model.index(m, np.array([4, 6, 5, 5], dtype=np.int32))
    errors = dict(
        token_id_out_of_range=expect_error(lambda: model.search(np.array([[20]], dtype=np.int32), top_k=2)),
        int64_queries=expect_error(lambda: model.search(np.array([[1]], dtype=np.int64), top_k=2)),
        one_dim_queries=expect_error(lambda: model.search(np.array([1, 2], dtype=np.int32), top_k=2)),
        k_gt_num_docs=expect_error(lambda: model.search(np.array([[1]], dtype=np.int32), top_k=5)),
    )
    e_ids, e_sc = model.search(np.zeros((0, 3), dtype=np.int32), top_k=3)
    out = dict(
        source="reference bm25_native.BM25v on animal_index_bm25 (bm25_native.py:76-158)",
        indptr=indptr.tolist(), indices=indices.tolist(), data_bits=bits(data), data=f32_list(data),
        params=params, vocab=vocab, doc_lengths=[4, 6, 5, 5], cases=cases, errors=errors,
        empty=dict(ids_shape=list(e_ids.shape), scores_shape=list(e_sc.shape),
                   ids_dtype=str(e_ids.dtype), scores_dtype=str(e_sc.dtype)),
    )
    json.dump(out, open(os.path.join(HERE, "golden_bundled.json"), "w"), indent=1)


FOX = [
    "The quick brown fox jumps over the lazy dog",
    "Some other text",
    "The quick rabbit runs past the brown fox",
    "The quick rabbit jumps over the brown dog",
    "The quick dog chases past the lazy fox",
    "The quick dog runs through the tall trees",
    "The quick brown fox jumps over the lazy dog",
    "The brown dog sleeps under the shady tree",
    "The brown rabbit hops under the tall tree",
    "The brown fox runs through the forest trees",
    "The brown fox watches the sleeping rabbit",
    "The lazy fox watches over the sleeping dog",
    "The lazy dog watches the quick rabbit",
]
ANIMAL = [
    "a cat is a feline and likes to purr",
    "a dog is the human's best friend and loves to play",
    "a bird is a beautiful animal that can fly",
    "a fish is a creature that lives in water and swims",
]


def golden_dense():
    out = dict(source="reference bm25.BM25 (bm25.py:30-178)", corpora={})
    for name, docs, queries in [
        ("fox", FOX, ["quick brown fox", "lazy dog", "tall trees forest", "some other text", "zzz",
                      "the the fox", "", "fox zzz dog"]),
        ("animal", ANIMAL, ["does the fish purr like a cat?", "a", "dog play friend"]),
    ]:
        corpus = [d.lower().split() for d in docs]
        model = ref_bm25.BM25()
        model.fit(corpus)
        entry = dict(
            docs=docs,
            vocabulary=model.vocabulary,
            avgdl=float(model.avgdl),
            matrix_dtype=str(model.bm25_matrix.dtype),
            matrix=[[float(x) for x in row] for row in model.bm25_matrix],
            queries=[],
        )
        for q in queries:
            toks = q.lower().split()
            scores = model.get_scores(toks)
            per_n = {}
            for n in (0, 1, 5, 10, 100):
                top = model.get_top_n(toks, corpus, n=n)
                per_n[str(n)] = dict(scores=[float(s) for s, _ in top], docs=[" ".join(d) for _, d in top])
            entry["queries"].append(dict(query=q, scores=[float(s) for s in scores], top_n=per_n))
        out["corpora"][name] = entry
    json.dump(out, open(os.path.join(HERE, "golden_dense.json"), "w"), indent=1)


def golden_selfcheck():
    dense = np.array([[1.0, 2.0, 3.0], [2.0, 4.0, 1.0]], dtype=np.float32)
    m = sp.csc_matrix(dense)
    model = ref_native.BM25v()
    model.index(m, np.array([2], dtype=np.int32))
    ids, sc = model.search(np.array([[0, 1]], dtype=np.int32), top_k=1)
    out = dict(source="bm25_native.py:232-243", dense=dense.tolist(), indptr=m.indptr.tolist(),
               indices=m.indices.tolist(), data=f32_list(m.data), query=[[0, 1]], k=1,
               ids=ids.tolist(), scores=f32_list(sc))
    json.dump(out, open(os.path.join(HERE, "golden_selfcheck.json"), "w"), indent=1)


def golden_random():
    rng = np.random.default_rng(20261018)
    store = {}
    specs = [  # name, n_docs, n_terms, density, Q, T, k
        ("r0", 1, 1, 1.0, 2, 1, 1),
        ("r1", 7, 5, 0.5, 6, 3, 7),
        ("r2", 64, 50, 0.2, 16, 4, 10),
        ("r3", 300, 40, 0.3, 12, 8, 100),
        ("r4", 2000, 200, 0.05, 24, 6, 10),
        ("r5", 1500, 30, 0.6, 8, 16, 1000),
        ("r6", 4099, 64, 0.1, 10, 5, 33),
    ]
    for name, n_docs, n_terms, dens, q_n, t_n, k in specs:
        m = sp.random(n_docs, n_terms, density=dens, format="csc", dtype=np.float32,
                      random_state=np.random.RandomState(int(rng.integers(1 << 31))),
                      data_rvs=lambda n: (0.05 + 3.0 * rng.random(n)).astype(np.float32))
        m.sort_indices()
        queries = rng.integers(0, n_terms, size=(q_n, t_n)).astype(np.int32)
        pad = rng.random((q_n, t_n)) < 0.25
        queries[pad] = -1
        if q_n > 1:
            queries[0, :] = -1  # all-padding query
            queries[1, :] = queries[1, 0] if queries[1, 0] >= 0 else 0  # repeated term
        ids, sc, dense = run_bm25v(m.indptr, m.indices, m.data, n_docs, queries, k)
        store[f"{name}_indptr"] = m.indptr.astype(np.int32)
        store[f"{name}_indices"] = m.indices.astype(np.int32)
        store[f"{name}_data"] = m.data.astype(np.float32)
        store[f"{name}_meta"] = np.array([n_docs, n_terms, k], dtype=np.int64)
        store[f"{name}_queries"] = queries
        store[f"{name}_ids"] = ids.astype(np.int32)
        store[f"{name}_scores"] = sc.astype(np.float32)
        store[f"{name}_dense"] = dense
    store["names"] = np.array([s[0] for s in specs])
    np.savez_compressed(os.path.join(HERE, "golden_random.npz"), **store)


def golden_disk_index():
    """Byte-for-byte copy of the bundled on-disk index (data fixture for the loader/writer tests)."""
    import shutil

    dst = os.path.join(HERE, "animal_index_bm25")
    shutil.rmtree(dst, ignore_errors=True)
    shutil.copytree(os.path.join(REF, "animal_index_bm25"), dst)
    for root, _, files in os.walk(dst):
        for f in files:
            os.chmod(os.path.join(root, f), 0o644)
        os.chmod(root, 0o755)


if __name__ == "__main__":
    golden_disk_index()
    golden_bundled()
    golden_dense()
    golden_selfcheck()
    golden_random()
    print("golden vectors written to", HERE)
