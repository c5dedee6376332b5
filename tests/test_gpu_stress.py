"""Determinism / protocol stress of k_score_topk's candidate-buffer overflow rounds (warps reach a
named barrier from different loop positions; compute-sanitizer is closed on this GPU pool, so the
protocol is covered by repetition): the overflow-heavy launch shapes, repeated many times in one
process and across fresh handles, with workspace and shared memory poisoned (0xff) before every
launch, must return bit-identical, oracle-exact results every time."""
import numpy as np
import pytest

from oracle import bm25_oracle as orc
from oracle import c_oracle

pytestmark = pytest.mark.gpu

NAMES = ["cap", "consumer_warps", "tile_docs", "cand_smem", "poison", "splits", "heavy_min", "no_hot", "generic_kernel"]


def _variants(k):
    out = []
    for warps in (1, 4):
        for tile in (128, 512):
            for cand_smem in (0, 1):
                out.append(dict(cap=k + 64, consumer_warps=warps, tile_docs=tile, cand_smem=cand_smem, poison=1))
    out.append(dict(cap=k + 64, consumer_warps=8, tile_docs=2048, poison=1, splits=1))
    out.append(dict(cap=k + 64, consumer_warps=4, tile_docs=512, poison=1, heavy_min=1 << 20))
    out.append(dict(cap=k + 64, consumer_warps=4, tile_docs=512, poison=1, no_hot=1))
    out.append(dict(cap=k + 64, consumer_warps=4, tile_docs=512, poison=1, generic_kernel=1))
    return out


@pytest.mark.parametrize("workload,scale,k,reps", [("E", 0.1, 1000, 30), ("B", 0.2, 100, 30)])
def test_overflow_rounds_are_deterministic(workload, scale, k, reps):
    from mojo_bm25_b200 import engine, synth

    idx, q, _ = synth.make_workload(workload, scale=scale)
    indptr, indices, data = idx.numpy()
    q = q.numpy()[:40]
    ref = None
    for fresh in range(2):  # across fresh handles
        index = engine.DeviceIndex(indptr, indices, data, n_docs=idx.n_docs)
        if ref is None:
            ref = index.search(q, k)
            for i in range(0, len(q), 5):
                dense = c_oracle.scores_dense(indptr, indices, data, idx.n_docs, q[i])
                orc.check_topk_against_dense(ref[0][i], ref[1][i], dense, k, exact=True)
        for v in _variants(k):
            for n in NAMES:
                index.set_option(n, v.get(n, 0))
            for _ in range(reps if fresh == 0 else 5):
                ids, sc = index.search(q, k)
                assert np.array_equal(ids, ref[0]), v
                assert np.array_equal(sc.view(np.uint32), ref[1].view(np.uint32)), v
        index.close()


def test_general_path_with_overflow_rounds_and_shared_thresholds():
    """an index with non-positive weights at a size where candidate-buffer
    overflow rounds and the per-query threshold shared between CTAs are exercised.  A negative-IDF
    term present in 90 % of the documents: the top-k must contain zero-score documents (ordered
    by ascending id) before any negative-score document."""
    from mojo_bm25_b200 import engine

    rng = np.random.default_rng(17)
    n_docs, n_terms, k = 120_000, 12, 100
    cols, ptr = [], [0]
    dens = [0.9, 0.002, 0.3, 0.0005, 0.05, 0.9, 0.01, 0.001, 0.2, 0.0, 0.6, 0.003]
    for t in range(n_terms):
        rows = np.flatnonzero(rng.random(n_docs) < dens[t]).astype(np.int32)
        cols.append(rows)
        ptr.append(ptr[-1] + len(rows))
    indices = np.concatenate(cols)
    indptr = np.array(ptr, np.int32)
    data = np.empty(len(indices), np.float32)
    for t in range(n_terms):
        seg = slice(ptr[t], ptr[t + 1])
        n = ptr[t + 1] - ptr[t]
        if t in (0, 5):   # "negative idf": very common terms
            data[seg] = -(0.05 + 0.3 * rng.random(n)).astype(np.float32)
        elif t == 2:      # mixed signs and exact zeros
            data[seg] = rng.normal(size=n).astype(np.float32)
            data[seg][::7] = 0.0
        else:
            data[seg] = (0.01 + rng.random(n)).astype(np.float32)
    index = engine.DeviceIndex(indptr, indices, data, n_docs=n_docs)
    assert index.info.all_positive == 0
    q = np.array([[0, -1, -1], [0, 5, -1], [0, 2, 5], [5, 9, 0], [2, -1, -1], [0, 3, 7], [1, 3, 7], [0, 10, 8]], np.int32)
    for opts in [dict(), dict(cap=k + 64, splits=6), dict(consumer_warps=2, tile_docs=256, cap=k + 64, poison=1),
                 dict(splits=1, no_theta_share=1)]:
        for n in ["cap", "splits", "consumer_warps", "tile_docs", "poison", "no_theta_share"]:
            index.set_option(n, opts.get(n, 0))
        ids, sc = index.search(q, k)
        for i in range(len(q)):
            dense = c_oracle.scores_dense(indptr, indices, data, n_docs, q[i])
            orc.check_topk_against_dense(ids[i], sc[i], dense, k, exact=True)
            same = sc[i][1:] == sc[i][:-1]
            assert np.all(ids[i][1:][same] > ids[i][:-1][same])
    # query 0: one negative term in 90 % of the docs -> the best 100 are zero-score docs, lowest ids first
    ids, sc = index.search(q[:1], k)
    assert not sc.any()
    zero_docs = np.setdiff1d(np.arange(n_docs), cols[0])[:k]
    assert np.array_equal(ids[0], zero_docs)


def test_large_k_uses_the_global_memory_merge_and_the_maximum_is_a_value_error():
    """BM25v._topk accepts any k <= n_docs (bm25_native.py:204-214).  Up to BM25_SMALL_K the final
    merge sorts in shared memory, above it in global memory (k_merge_large); above BM25_MAX_K the
    call is rejected with a ValueError instead of an internal error."""
    import scipy.sparse as sp
    from mojo_bm25_b200 import _lib, engine, sharded
    from mojo_bm25_b200.bm25_native import BM25v

    rng = np.random.default_rng(2)
    n_docs = _lib.MAX_K + 3000
    m = sp.random(n_docs, 8, density=0.25, format="csc", dtype=np.float32, random_state=np.random.RandomState(0),
                  data_rvs=lambda n: (0.01 + rng.random(n)).astype(np.float32))
    m.sort_indices()
    indptr, indices, data = m.indptr.astype(np.int32), m.indices.astype(np.int32), m.data
    index = engine.DeviceIndex(indptr, indices, data, n_docs=n_docs)
    q = np.array([[0, 1, -1], [2, 3, 7], [5, -1, -1]], np.int32)
    for k in (_lib.SMALL_K, _lib.SMALL_K + 1, 10_000, 30_000, _lib.MAX_K):
        ids, sc = index.search(q, k)
        for i in range(len(q)):
            dense = c_oracle.scores_dense(indptr, indices, data, n_docs, q[i])
            orc.check_topk_against_dense(ids[i], sc[i], dense, k, exact=True)
            same = sc[i][1:] == sc[i][:-1]
            assert np.all(ids[i][1:][same] > ids[i][:-1][same])
    with pytest.raises(ValueError, match="BM25_MAX_K"):
        index.search(q, _lib.MAX_K + 1)
    model = BM25v()
    model.index(m, np.ones(n_docs, np.int32))
    with pytest.raises(ValueError):
        model.search(q, top_k=_lib.MAX_K + 1)
    # the shard merge takes the same path for a large k_out
    import torch

    k = 9000
    parts = orc.partition_csc_by_doc_range(indptr, indices, data, n_docs, 3)
    shards = [engine.DeviceIndex(p, i_, d, n_docs=nd, doc_id_base=base) for p, i_, d, nd, base in parts]
    s_ = sharded.DocShardedSearcher.from_index(shards, k)
    gi, gs = s_.search(torch.from_numpy(q).cuda())
    torch.cuda.synchronize()
    wi, ws = index.search(q, k)
    assert np.array_equal(gi.cpu().numpy(), wi) and np.array_equal(gs.cpu().numpy().view(np.uint32), ws.view(np.uint32))


@pytest.mark.parametrize("lo_exp,hi_exp", [(-2, 2), (-12, 3), (-30, 30), (-60, 36), (-110, -90), (20, 36)])
def test_exponent_epochs_are_exact_for_any_weight_range(lo_exp, hi_exp):
    """k_score_topk_s does not zero its score tile after every tile: consecutive tiles accumulate
    weight * 2^(s0 + c*e) and rely on the leftovers of earlier tiles being absorbed exactly.  The
    spacing c and the run length are derived from the query's weight range, so sweep that range
    (log-uniform weights over 10^lo .. 10^hi, far beyond what BM25 produces): results must be
    bit-identical to the oracle and to the no_epoch build of the same search, with one warp
    walking every tile of the index so that runs wrap several times."""
    from mojo_bm25_b200 import engine

    rng = np.random.default_rng(1000 + lo_exp * 7 + hi_exp)
    n_docs, n_terms = 60_000, 40
    cols, vals = [], []
    indptr = np.zeros(n_terms + 1, dtype=np.int32)
    for t in range(n_terms):
        df = int(rng.integers(1, n_docs // (1 + t % 7)))
        docs = np.sort(rng.choice(n_docs, size=df, replace=False)).astype(np.int32)
        # every term spans the whole range: the smallest and the largest weight meet in one slot
        w = (10.0 ** rng.uniform(lo_exp, hi_exp, size=df)).astype(np.float32)
        w = np.minimum(np.maximum(w, np.float32(1.2e-38)), np.float32(1e37))  # finite sums of five
        cols.append(docs)
        vals.append(w)
        indptr[t + 1] = indptr[t] + df
    indices = np.concatenate(cols)
    data = np.concatenate(vals)
    q = rng.integers(0, n_terms, size=(24, 5)).astype(np.int32)
    q[3, 2:] = -1
    q[7, :] = q[7, 0]  # the same term five times
    k = 50
    index = engine.DeviceIndex(indptr, indices, data, n_docs=n_docs)
    names = ["consumer_warps", "splits", "tile_docs", "no_epoch", "poison", "cap", "heavy_min"]
    ref = None
    for v in [dict(no_epoch=1), dict(), dict(consumer_warps=1, splits=1, poison=1),
              dict(consumer_warps=1, splits=1, tile_docs=512, poison=1),
              dict(consumer_warps=2, splits=1, tile_docs=128, cap=k + 64, poison=1),
              dict(consumer_warps=1, splits=1, heavy_min=1 << 20, poison=1)]:
        for n in names:
            index.set_option(n, v.get(n, 0))
        ids, sc = index.search(q, k)
        if ref is None:
            ref = (ids, sc)
            for i in range(len(q)):
                dense = c_oracle.scores_dense(indptr, indices, data, n_docs, q[i])
                orc.check_topk_against_dense(ids[i], sc[i], dense, k, exact=True)
        assert np.array_equal(ids, ref[0]), v
        assert np.array_equal(sc.view(np.uint32), ref[1].view(np.uint32)), v
    index.close()
