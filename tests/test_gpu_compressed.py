"""-m gpu parity tests of the compressed posting format (the reference's
index declares "dtype": "float32", "int_dtype": "int32", animal_index_bm25/params.index.json:1-12).

A compressed handle (bm25_index_compress: 4-byte postings = uint16 tile-local slot + bf16 weight) IS
the index whose weights are rounded to bf16: every result must be bit-identical to the oracle
(reference hot loop bm25_native.py:129-158) run on the CSC matrix ``round_to_bf16(data)``.
"""
import os

import numpy as np
import pytest

from oracle import bm25_oracle as orc
from oracle import c_oracle

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def engine():
    from mojo_bm25_b200 import engine as eng

    return eng


def _check(index, indptr, indices, data_q, n_docs, queries, k):
    ids, sc = index.search(queries, k)
    for i in range(len(queries)):
        dense = c_oracle.scores_dense(indptr, indices, data_q, n_docs, queries[i])
        orc.check_topk_against_dense(ids[i], sc[i], dense, k, exact=True)
    return ids, sc


@pytest.mark.parametrize("workload,scale,k", [("tiny", 1.0, 10), ("B", 0.05, 10), ("B", 0.05, 100),
                                              ("C", 0.005, 100), ("10Mc", 0.01, 100), ("E", 0.03, 1000)])
def test_compressed_index_equals_oracle_on_rounded_weights(engine, workload, scale, k):
    from mojo_bm25_b200 import synth

    idx, q, _ = synth.make_workload(workload, scale=scale)
    indptr, indices, data = idx.numpy()
    q = q.numpy()[:48]
    k = min(k, idx.n_docs)
    data_q = engine.round_to_bf16(data)
    index = engine.DeviceIndex(indptr, indices, data, n_docs=idx.n_docs)
    before = index.info
    assert before.weight_format == 0 and before.posting_bytes == 8
    bytes8 = index.posting_bytes(q, k)
    index.compress("bf16")
    info = index.info
    assert info.weight_format == 1 and info.posting_bytes == 4
    assert index.posting_bytes(q, k) - 8 * k * len(q) == (bytes8 - 8 * k * len(q)) // 2
    ref = _check(index, indptr, indices, data_q, idx.n_docs, q, k)
    # dense scores of the handle are those of the rounded matrix, bit for bit
    dense = index.scores_dense(q[:4])
    for i in range(4):
        want = c_oracle.scores_dense(indptr, indices, data_q, idx.n_docs, q[i])
        assert np.array_equal(dense[i].view(np.uint32), want.view(np.uint32))
    # packed and unpacked kernels of the same handle, other tilings / launch shapes: identical bits
    for opts in [dict(no_packed=1), dict(tile_docs=512, splits=3), dict(tile_docs=4096, consumer_warps=4),
                 dict(tile_docs=8192, consumer_warps=2, splits=1), dict(no_epoch=1, poison=1), dict(heavy_min=1),
                 dict(generic_kernel=1)]:
        for n in ["no_packed", "tile_docs", "splits", "consumer_warps", "no_epoch", "poison", "heavy_min", "generic_kernel"]:
            index.set_option(n, opts.get(n, 0))
        ids, sc = index.search(q, k)
        assert np.array_equal(ids, ref[0]), opts
        assert np.array_equal(sc.view(np.uint32), ref[1].view(np.uint32)), opts
    index.compress("bf16")  # idempotent
    index.close()


def test_compressed_matches_uncompressed_handle_of_rounded_matrix(engine):
    """Two routes to the same numbers: compress(fp32 handle) == fp32 handle built from the rounded matrix."""
    from mojo_bm25_b200 import synth

    idx, q, _ = synth.make_workload("B", scale=0.1)
    indptr, indices, data = idx.numpy()
    q = q.numpy()[:200]
    a = engine.DeviceIndex(indptr, indices, data, n_docs=idx.n_docs).compress()
    b = engine.DeviceIndex(indptr, indices, engine.round_to_bf16(data), n_docs=idx.n_docs)
    for k in (10, 100):
        ia, sa = a.search(q, k)
        ib, sb = b.search(q, k)
        assert np.array_equal(ia, ib) and np.array_equal(sa.view(np.uint32), sb.view(np.uint32))
    a.close()
    b.close()


def test_compressed_tile_limit_and_bad_format(engine):
    from mojo_bm25_b200 import synth

    idx, q, k = synth.make_workload("tiny")
    indptr, indices, data = idx.numpy()
    index = engine.DeviceIndex(indptr, indices, data, n_docs=idx.n_docs)
    with pytest.raises(ValueError):
        index.compress("fp8")
    index.compress()
    bidx, bq, _ = synth.make_workload("B", scale=0.05)
    big = engine.DeviceIndex(*bidx.numpy(), n_docs=bidx.n_docs).compress()
    big.set_option("tile_docs", 16384)  # tile-local slots are 16-bit byte offsets: 8192 documents at most
    with pytest.raises(Exception):
        big.search(bq.numpy()[:2], 5)
    big.set_option("tile_docs", 0)
    big.search(bq.numpy()[:2], 5)
    big.close()
    index.close()


def test_compressed_fuzz(engine):
    """Seeded fuzz over index shapes, ragged queries and launch knobs, compressed handles only."""
    rng = np.random.default_rng(int(os.environ.get("BM25_FUZZ_SEED", "20260118")) + 7)
    for trial in range(30):
        n_docs = int(rng.choice([1, 2, 37, 500, 2049, 7000, 30000]))
        n_terms = int(rng.integers(1, 40))
        cols, ptr = [], [0]
        for t in range(n_terms):
            dens = [0.0, 0.002, 0.05, 0.9][rng.integers(0, 4)]
            rows = np.flatnonzero(rng.random(n_docs) < dens).astype(np.int32)
            cols.append(rows)
            ptr.append(ptr[-1] + len(rows))
        indices = np.concatenate(cols) if ptr[-1] else np.zeros(0, np.int32)
        data = (0.01 + rng.random(ptr[-1]) * rng.choice([1.0, 8.0])).astype(np.float32)
        indptr = np.array(ptr, np.int32)
        index = engine.DeviceIndex(indptr, indices, data, n_docs=n_docs).compress()
        data_q = engine.round_to_bf16(data)
        q = rng.integers(-1, n_terms, size=(int(rng.integers(1, 9)), int(rng.integers(1, 12)))).astype(np.int32)
        k = int(min(n_docs, rng.choice([1, 3, 10, 100, 1000])))
        for name, choices in [("tile_docs", [0, 128, 512, 4096]), ("consumer_warps", [0, 1, 3, 8, 16]),
                              ("splits", [0, 1, 2, 9]), ("cap", [0, k + 64]), ("no_hot", [0, 1]),
                              ("heavy_min", [0, 1, 512, 1 << 20]), ("poison", [0, 1]), ("no_epoch", [0, 1])]:
            index.set_option(name, int(rng.choice(choices)))
        _check(index, indptr, indices, data_q, n_docs, q, k)
        index.close()


def test_compress_edge_cases(engine):
    """Empty index, non-positive weights (general path) and weights that round to zero in bf16."""
    # no postings at all
    idx = engine.DeviceIndex(np.zeros(4, np.int32), np.zeros(0, np.int32), np.zeros(0, np.float32), n_docs=5).compress()
    ids, sc = idx.search(np.array([[0, 2]], np.int32), 3)
    assert ids.tolist() == [[0, 1, 2]] and sc.tolist() == [[0.0, 0.0, 0.0]]
    idx.close()
    rng = np.random.default_rng(5)
    n_docs, n_terms = 9000, 12
    cols = [np.sort(rng.choice(n_docs, size=int(rng.integers(1, 4000)), replace=False)).astype(np.int32) for _ in range(n_terms)]
    indptr = np.concatenate([[0], np.cumsum([len(c) for c in cols])]).astype(np.int32)
    indices = np.concatenate(cols)
    q = rng.integers(-1, n_terms, size=(6, 5)).astype(np.int32)
    # (a) mixed-sign weights: the general path of a compressed handle
    data = rng.standard_normal(len(indices)).astype(np.float32)
    h = engine.DeviceIndex(indptr, indices, data, n_docs=n_docs).compress()
    assert h.info.all_positive == 0
    _check(h, indptr, indices, engine.round_to_bf16(data), n_docs, q, 50)
    h.close()
    # (b) positive weights some of which round to +0.0 in bf16 (fp32 denormals): the handle must notice
    data = (0.05 + rng.random(len(indices))).astype(np.float32)
    data[::7] = np.float32(1e-45)
    h = engine.DeviceIndex(indptr, indices, data, n_docs=n_docs)
    assert h.info.all_positive == 1
    h.compress()
    dq = engine.round_to_bf16(data)
    assert (dq == 0).any() and h.info.all_positive == 0
    _check(h, indptr, indices, dq, n_docs, q, 50)
    h.close()


def test_alternating_query_shapes_on_one_handle(engine):
    """Narrow queries run on 2048-document tiles, 64-term queries on 1792-document tiles (launch plan);
    one handle keeps both tile tables (and both packed arrays) and answers both shapes, interleaved."""
    from mojo_bm25_b200 import synth

    idx, q4, _ = synth.make_workload("B", scale=0.2)
    indptr, indices, data = idx.numpy()
    q4 = q4.numpy()[:16]
    rng = np.random.default_rng(9)
    q64 = rng.integers(0, idx.n_terms, size=(8, 64)).astype(np.int32)
    q64[:, :4] = rng.integers(0, 50, size=(8, 4))  # a few stop-word-length lists
    for compress in (False, True):
        h = engine.DeviceIndex(indptr, indices, data, n_docs=idx.n_docs)
        dq = data
        if compress:
            h.compress()
            dq = engine.round_to_bf16(data)
        for _ in range(3):
            _check(h, indptr, indices, dq, idx.n_docs, q4, 10)
            _check(h, indptr, indices, dq, idx.n_docs, q64, 300)
        h.close()


def test_compress_refuses_weights_that_overflow_bf16_and_leaves_the_handle_intact(engine):
    indptr = np.array([0, 2, 3], np.int32)
    indices = np.array([0, 1, 1], np.int32)
    data = np.array([1.0, 3.4e38, 2.0], np.float32)  # 3.4e38 rounds to +inf in bf16
    h = engine.DeviceIndex(indptr, indices, data, n_docs=2)
    with pytest.raises(ValueError):
        h.compress()
    assert h.info.weight_format == 0
    ids, sc = h.search(np.array([[0, 1]], np.int32), 2)
    assert ids.tolist() == [[1, 0]] and np.array_equal(sc, np.array([[np.float32(3.4e38) + np.float32(2.0), 1.0]], np.float32))
    h.close()
