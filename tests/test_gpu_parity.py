"""GPU parity tests proper: the CUDA path, called through the C ABI (libbm25_b200.so), against the
CPU oracle and the committed golden vectors.  Run with `-m gpu` on the B200 box.

Bars: document ids must be an exact top-k of the oracle's dense score vector (tie-aware checker),
scores are compared BITWISE (the kernels accumulate in query-term order like the reference), which
is stricter than the 1e-5 relative tolerance BASELINE.json's north_star states.
"""
import json
import os

import numpy as np
import pytest

from oracle import bm25_oracle as orc
from oracle import c_oracle

pytestmark = pytest.mark.gpu

RTOL = 1e-5  # north_star tolerance; the asserts below are bitwise (exact=True) unless noted


def _bits_to_f32(b):
    return np.array(b, dtype=np.uint32).view(np.float32)


@pytest.fixture(scope="module")
def engine():
    from mojo_bm25_b200 import engine as eng

    return eng


def _check_batch(index, indptr, indices, data, n_docs, queries, k, exact=True):
    ids, sc = index.search(queries, k)
    assert ids.dtype == np.int32 and sc.dtype == np.float32 and ids.shape == (len(queries), k)
    for i in range(len(queries)):
        dense = c_oracle.scores_dense(indptr, indices, data, n_docs, queries[i])
        orc.check_topk_against_dense(ids[i], sc[i], dense, k, rtol=RTOL, exact=exact)
        # our deterministic tie rule: equal scores -> ascending doc id
        same = sc[i][1:] == sc[i][:-1]
        assert np.all(ids[i][1:][same] > ids[i][:-1][same])
    return ids, sc


def test_bundled_index_golden(engine, golden_dir):
    g = json.load(open(os.path.join(golden_dir, "golden_bundled.json")))
    indptr, indices = np.array(g["indptr"], np.int32), np.array(g["indices"], np.int32)
    data = _bits_to_f32(g["data_bits"])
    index = engine.DeviceIndex(indptr, indices, data, n_docs=4)
    assert index.info.all_positive == 1 and index.info.n_terms == 20 and index.info.nnz == 20
    for case in g["cases"]:
        q = np.array(case["queries"], np.int32)
        k = case["k"]
        ids, sc = _check_batch(index, indptr, indices, data, 4, q, k)
        ref_sc = _bits_to_f32(case["score_bits"]).reshape(sc.shape)
        assert np.array_equal(sc.view(np.uint32), ref_sc.view(np.uint32))
        dense = index.scores_dense(q)
        assert np.array_equal(dense.view(np.uint32), _bits_to_f32(case["dense_bits"]).reshape(dense.shape).view(np.uint32))
        for i in range(len(q)):
            orc.assert_same_topk_modulo_ties(ids[i], sc[i], np.array(case["ids"][i]), ref_sc[i], rtol=RTOL)
    # literal G1 known answers
    ids, sc = index.search(np.array([[17, 16, 2, 0]], np.int32), 2)
    assert ids.tolist() == [[0, 3]]
    np.testing.assert_allclose(sc, [[1.5876564, 0.48158914]], rtol=1e-6)


def test_random_golden_bitwise(engine, golden_dir):
    z = np.load(os.path.join(golden_dir, "golden_random.npz"))
    for name in z["names"].tolist():
        n_docs, n_terms, k = z[f"{name}_meta"].tolist()
        indptr, indices, data = z[f"{name}_indptr"], z[f"{name}_indices"], z[f"{name}_data"]
        index = engine.DeviceIndex(indptr, indices, data, n_docs=n_docs)
        q = z[f"{name}_queries"]
        ids, sc = _check_batch(index, indptr, indices, data, n_docs, q, k)
        assert np.array_equal(sc.view(np.uint32), z[f"{name}_scores"].view(np.uint32)), name
        dense = index.scores_dense(q)
        assert np.array_equal(dense.view(np.uint32), z[f"{name}_dense"].view(np.uint32)), name
        for i in range(len(q)):
            orc.assert_same_topk_modulo_ties(ids[i], sc[i], z[f"{name}_ids"][i], z[f"{name}_scores"][i], rtol=RTOL)
        index.close()


def test_selfcheck_g4(engine, golden_dir):
    g = json.load(open(os.path.join(golden_dir, "golden_selfcheck.json")))
    index = engine.DeviceIndex(np.array(g["indptr"], np.int32), np.array(g["indices"], np.int32),
                               np.array(g["data"], np.float32), n_docs=2)
    ids, sc = index.search(np.array(g["query"], np.int32), 1)
    assert ids.tolist() == [[1]] and sc.tolist() == [[6.0]]


@pytest.mark.parametrize("workload,scale,k", [("tiny", 1.0, 10), ("B", 0.05, 10), ("B", 0.05, 100),
                                              ("E", 0.03, 1000), ("C", 0.005, 100), ("Bc", 0.05, 10),
                                              ("10Mc", 0.01, 100)])
def test_synthetic_workloads_against_oracle(engine, workload, scale, k):
    from mojo_bm25_b200 import synth

    idx, q, _ = synth.make_workload(workload, scale=scale)
    indptr, indices, data = idx.numpy()
    q = q.numpy()[:48]
    k = min(k, idx.n_docs)
    index = engine.DeviceIndex(indptr, indices, data, n_docs=idx.n_docs)
    _check_batch(index, indptr, indices, data, idx.n_docs, q, k)
    # the same through different tilings / tile-range splits
    for tile_docs, splits in [(2048, 1), (4096, 3), (0, 7)]:
        index.set_option("tile_docs", tile_docs)
        index.set_option("splits", splits)
        _check_batch(index, indptr, indices, data, idx.n_docs, q[:8], k)


def test_general_path_negative_and_zero_weights(engine):
    rng = np.random.default_rng(5)
    import scipy.sparse as sp

    m = sp.random(5000, 60, density=0.08, format="csc", dtype=np.float32, random_state=np.random.RandomState(2),
                  data_rvs=lambda n: rng.normal(size=n).astype(np.float32))
    m.sort_indices()
    m.data[::17] = 0.0
    index = engine.DeviceIndex(m.indptr, m.indices, m.data, n_docs=5000)
    assert index.info.all_positive == 0
    q = rng.integers(-1, 60, size=(12, 6)).astype(np.int32)
    for k in (1, 7, 300, 5000):
        _check_batch(index, m.indptr, m.indices, m.data, 5000, q, k)


def test_forced_general_path_equals_pruned_path(engine):
    from mojo_bm25_b200 import synth

    idx, q, k = synth.make_workload("tiny")
    indptr, indices, data = idx.numpy()
    q = q.numpy()
    index = engine.DeviceIndex(indptr, indices, data, n_docs=idx.n_docs)
    a = index.search(q, k)
    index.set_option("force_general", 1)
    b = index.search(q, k)
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1].view(np.uint32), b[1].view(np.uint32))


def test_unsorted_and_duplicate_postings_are_canonicalised(engine):
    # column 0: rows out of order; column 1: a duplicated row (summed, as scipy's mat-vec would)
    indptr = np.array([0, 3, 6], np.int32)
    indices = np.array([5, 1, 3, 2, 2, 0], np.int32)
    data = np.array([1.0, 2.0, 3.0, 0.5, 0.25, 4.0], np.float32)
    index = engine.DeviceIndex(indptr, indices, data, n_docs=6)
    assert index.info.was_sorted == 0
    dense = index.scores_dense(np.array([[0, 1]], np.int32))[0]
    np.testing.assert_array_equal(dense, np.array([4.0, 2.0, 0.75, 3.0, 0.0, 1.0], np.float32))
    ids, sc = index.search(np.array([[0, 1]], np.int32), 6)
    assert ids.tolist() == [[0, 3, 1, 5, 2, 4]]


def test_fewer_matches_than_k_fills_with_zero_score_docs(engine):
    indptr = np.array([0, 2, 3], np.int32)
    indices = np.array([7, 900, 3], np.int32)
    data = np.array([1.0, 2.0, 5.0], np.float32)
    index = engine.DeviceIndex(indptr, indices, data, n_docs=1000)
    ids, sc = index.search(np.array([[0, -1], [1, 0], [-1, -1]], np.int32), 5)
    assert ids[0].tolist() == [900, 7, 0, 1, 2] and sc[0].tolist() == [2.0, 1.0, 0.0, 0.0, 0.0]
    assert ids[1].tolist() == [3, 900, 7, 0, 1] and sc[1].tolist() == [5.0, 2.0, 1.0, 0.0, 0.0]
    assert ids[2].tolist() == [0, 1, 2, 3, 4] and sc[2].tolist() == [0.0] * 5


def test_error_behaviour_matches_reference(engine, golden_dir):
    g = json.load(open(os.path.join(golden_dir, "golden_bundled.json")))
    index = engine.DeviceIndex(np.array(g["indptr"], np.int32), np.array(g["indices"], np.int32),
                               _bits_to_f32(g["data_bits"]), n_docs=4)
    with pytest.raises(ValueError):  # token id >= n_terms (bm25_native.py:116-121)
        index.search(np.array([[20]], np.int32), 2)
    with pytest.raises(ValueError):  # k > num_docs (argpartition raises in the reference)
        index.search(np.array([[1]], np.int32), 5)
    with pytest.raises(ValueError):
        engine.DeviceIndex(np.array([0, 2], np.int32), np.array([0, 9], np.int32), np.array([1, 1], np.float32), n_docs=4)
    with pytest.raises(ValueError):
        engine.DeviceIndex(np.array([0, 1], np.int32), np.array([0], np.int32), np.array([np.nan], np.float32), n_docs=4)


def test_device_tensor_entry_points_and_doc_id_base(engine):
    import torch
    from mojo_bm25_b200 import synth

    idx, q, k = synth.make_workload("tiny", device="cuda")
    index = engine.DeviceIndex.from_torch(idx.indptr, idx.indices, idx.data, idx.n_docs, doc_id_base=1000, borrow=True)
    ids, sc = index.search_device(q, k)
    torch.cuda.synchronize()
    indptr, indices, data = idx.numpy()
    ids, sc, qn = ids.cpu().numpy(), sc.cpu().numpy(), q.cpu().numpy()
    for i in range(len(qn)):
        dense = c_oracle.scores_dense(indptr, indices, data, idx.n_docs, qn[i])
        orc.check_topk_against_dense(ids[i] - 1000, sc[i], dense, k, exact=True)
    # non-canonical device arrays are rejected loudly
    bad = idx.indices.clone()
    bad[:2] = bad[:2].flip(0)
    with pytest.raises(ValueError):
        engine.DeviceIndex.from_torch(idx.indptr, bad, idx.data, idx.n_docs)


def test_merge_topk_of_document_shards(engine):
    import torch
    from mojo_bm25_b200 import synth

    idx, q, _ = synth.make_workload("B", scale=0.04)
    indptr, indices, data = idx.numpy()
    qn = q.numpy()[:32]
    n_shards, k = 4, 50
    parts = orc.partition_csc_by_doc_range(indptr, indices, data, idx.n_docs, n_shards)
    all_ids = torch.empty((n_shards, len(qn), k), dtype=torch.int32, device="cuda")
    all_sc = torch.empty((n_shards, len(qn), k), dtype=torch.float32, device="cuda")
    qd = torch.from_numpy(qn).cuda()
    shards = []
    for g, (ptr, ind, dat, nd, base) in enumerate(parts):
        sh = engine.DeviceIndex(ptr, ind, dat, n_docs=nd, doc_id_base=base)
        shards.append(sh)
        sh.search_device(qd, k, out_ids=all_ids[g], out_scores=all_sc[g])
    ids, sc = engine.merge_topk_device(all_ids, all_sc, k)
    torch.cuda.synchronize()
    ids, sc = ids.cpu().numpy(), sc.cpu().numpy()
    for i in range(len(qn)):
        dense = c_oracle.scores_dense(indptr, indices, data, idx.n_docs, qn[i])
        orc.check_topk_against_dense(ids[i], sc[i], dense, k, exact=True)
    want_i, want_s = orc.merge_topk_lists(all_ids.cpu().numpy(), all_sc.cpu().numpy(), k)
    assert np.array_equal(ids, want_i) and np.array_equal(sc.view(np.uint32), want_s.view(np.uint32))


def test_full_size_config_b_properties(engine):
    """BASELINE config B at full size (1M docs, 100k terms, 1000 x 4, k=10): size-independent
    properties + oracle spot checks (the oracle is too slow for all 1000 queries)."""
    import torch
    from mojo_bm25_b200 import synth

    idx, q, k = synth.make_workload("B", device="cuda")
    index = engine.DeviceIndex.from_torch(idx.indptr, idx.indices, idx.data, idx.n_docs)
    ids, sc = index.search_device(q, k)
    ids2, sc2 = index.search_device(q.flip(0), k)  # permuting the batch permutes the results
    torch.cuda.synchronize()
    assert torch.equal(ids.flip(0), ids2) and torch.equal(sc.flip(0), sc2)
    ids, sc, qn = ids.cpu().numpy(), sc.cpu().numpy(), q.cpu().numpy()
    assert np.all(sc[:, :-1] >= sc[:, 1:]) and np.all(sc > 0)
    assert all(len(set(r.tolist())) == k for r in ids)
    # top-10 is a prefix of top-100 (idempotence under k)
    ids100, sc100 = index.search_device(q, 100)
    torch.cuda.synchronize()
    assert np.array_equal(ids100.cpu().numpy()[:, :k], ids) and np.array_equal(sc100.cpu().numpy()[:, :k], sc)
    indptr, indices, data = idx.numpy()
    for i in range(0, 1000, 97):
        dense = c_oracle.scores_dense(indptr, indices, data, idx.n_docs, qn[i])
        orc.check_topk_against_dense(ids[i], sc[i], dense, k, exact=True)
    # host-buffer entry point returns the same
    hid, hsc = index.search(qn, k)
    assert np.array_equal(hid, ids) and np.array_equal(hsc.view(np.uint32), sc.view(np.uint32))


@pytest.mark.parametrize("workload,scale,k", [("B", 0.2, 10), ("10M", 0.02, 100), ("E", 0.1, 1000), ("C", 0.01, 100)])
def test_every_launch_shape_gives_identical_results(engine, workload, scale, k):
    """The tuning knobs only change HOW the work is cut (warp tile size, warps per CTA, CTAs per
    query, candidate-buffer size -> number of overflow rounds, hot-list vs dense epilogue, threshold
    priming / sharing).  Every variant must return bit-identical ids and scores, and the default must
    match the oracle."""
    from mojo_bm25_b200 import synth

    idx, q, _ = synth.make_workload(workload, scale=scale)
    indptr, indices, data = idx.numpy()
    q = q.numpy()[:40]
    k = min(k, idx.n_docs)
    index = engine.DeviceIndex(indptr, indices, data, n_docs=idx.n_docs)
    ref_ids, ref_sc = _check_batch(index, indptr, indices, data, idx.n_docs, q[:6], k)
    ref_ids, ref_sc = index.search(q, k)
    variants = [
        dict(cap=k + 64), dict(cap=k + 64, consumer_warps=4, tile_docs=512), dict(no_hot=1), dict(no_priming=1),
        dict(no_priming=1, cap=k + 64, splits=1), dict(no_theta_share=1, splits=5), dict(consumer_warps=16, tile_docs=1024),
        dict(consumer_warps=12, tile_docs=4096, splits=2), dict(consumer_warps=1, tile_docs=128, splits=3),
        dict(waves=1), dict(waves=20, no_hot=1, no_priming=1),
        dict(heavy_min=1), dict(heavy_min=1 << 20), dict(heavy_min=256, tile_docs=1024, poison=1),
        dict(heavy_min=1 << 20, cap=k + 64, consumer_warps=3, poison=1),
        dict(no_query_sort=1), dict(no_bulk_clear=1, no_query_sort=1, splits=2),
        dict(generic_kernel=1), dict(generic_kernel=1, heavy_min=1 << 20, cap=k + 64),
        dict(q_major=1), dict(q_major=1, splits=4, cap=k + 64),
        dict(no_epoch=1), dict(no_epoch=1, cap=k + 64, poison=1), dict(poison=1, no_bulk_clear=1),
    ]
    names = ["cap", "consumer_warps", "tile_docs", "no_hot", "no_priming", "no_theta_share", "splits", "waves",
             "heavy_min", "poison", "no_query_sort", "no_bulk_clear", "generic_kernel", "q_major", "no_epoch"]
    for v in variants:
        for n in names:
            index.set_option(n, v.get(n, 0))
        ids, sc = index.search(q, k)
        assert np.array_equal(ids, ref_ids), v
        assert np.array_equal(sc.view(np.uint32), ref_sc.view(np.uint32)), v


def test_degenerate_shapes(engine):
    """Empty index, all-padding queries, T = 1, k = n_docs, a wide query row, a single document."""
    # empty index (no postings at all): every document scores 0 and the first k ids come back
    index = engine.DeviceIndex(np.zeros(6, np.int32), np.zeros(0, np.int32), np.zeros(0, np.float32), n_docs=300)
    ids, sc = index.search(np.array([[0, 4, -1]], np.int32), 7)
    assert ids.tolist() == [list(range(7))] and not sc.any()
    # one document, one term
    index = engine.DeviceIndex(np.array([0, 1], np.int32), np.array([0], np.int32), np.array([2.5], np.float32), n_docs=1)
    ids, sc = index.search(np.array([[0], [-1]], np.int32), 1)
    assert ids.tolist() == [[0], [0]] and sc.tolist() == [[2.5], [0.0]]
    # k = n_docs and a 200-slot query row (mostly padding, duplicated terms count with multiplicity)
    rng = np.random.default_rng(3)
    import scipy.sparse as sp

    m = sp.random(700, 40, density=0.2, format="csc", dtype=np.float32, random_state=np.random.RandomState(9),
                  data_rvs=lambda n: (0.05 + rng.random(n)).astype(np.float32))
    m.sort_indices()
    index = engine.DeviceIndex(m.indptr, m.indices, m.data, n_docs=700)
    q = np.full((5, 200), -1, np.int32)
    q[:, ::7] = rng.integers(0, 40, size=(5, 29))
    _check_batch(index, m.indptr.astype(np.int32), m.indices.astype(np.int32), m.data, 700, q, 700)
    _check_batch(index, m.indptr.astype(np.int32), m.indices.astype(np.int32), m.data, 700, q, 3)


def test_fuzz_random_indices_and_knobs(engine):
    """Seeded fuzz: random CSC shapes/densities (incl. heavy 'stop word' columns, empty columns,
    documents no term touches), random ragged queries with padding and duplicates, random k and
    random launch shapes -- ids and score bits must match the oracle every time."""
    rng = np.random.default_rng(int(os.environ.get("BM25_FUZZ_SEED", "20260118")))  # override to widen the fuzz
    for trial in range(40):
        n_docs = int(rng.choice([1, 2, 37, 500, 2049, 7000, 30000]))
        n_terms = int(rng.integers(1, 60))
        cols, ptr = [], [0]
        for t in range(n_terms):
            mode = rng.integers(0, 4)
            dens = [0.0, 0.002, 0.05, 0.9][mode]
            rows = np.flatnonzero(rng.random(n_docs) < dens).astype(np.int32)
            cols.append(rows)
            ptr.append(ptr[-1] + len(rows))
        indices = np.concatenate(cols) if ptr[-1] else np.zeros(0, np.int32)
        data = (0.01 + rng.random(ptr[-1]) * rng.choice([1.0, 8.0])).astype(np.float32)
        indptr = np.array(ptr, np.int32)
        index = engine.DeviceIndex(indptr, indices, data, n_docs=n_docs)
        n_q, width = int(rng.integers(1, 9)), int(rng.integers(1, 12))
        q = rng.integers(-1, n_terms, size=(n_q, width)).astype(np.int32)
        k = int(min(n_docs, rng.choice([1, 3, 10, 100, 1000])))
        for name, choices in [("tile_docs", [0, 128, 512, 4096]), ("consumer_warps", [0, 1, 3, 8, 16]),
                              ("splits", [0, 1, 2, 9]), ("cap", [0, k + 64]), ("no_hot", [0, 1]),
                              ("no_priming", [0, 1]), ("no_theta_share", [0, 1]), ("heavy_min", [0, 1, 512, 1 << 20]),
                              ("poison", [0, 1]), ("no_query_sort", [0, 1]), ("no_bulk_clear", [0, 1]),
                              ("generic_kernel", [0, 1]), ("q_major", [0, 1]), ("no_epoch", [0, 1])]:
            index.set_option(name, int(rng.choice(choices)))
        _check_batch(index, indptr, indices, data, n_docs, q, k)
        index.close()
