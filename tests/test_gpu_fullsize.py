"""Parity at BASELINE.json's FULL sizes (the oracle is too slow for whole batches, so a query
sample is checked bitwise against the C oracle on the same arrays) plus size-independent
properties over the whole batch.  ~1 minute on a B200 (index synthesis dominates)."""
import numpy as np
import pytest

from oracle import bm25_oracle as orc
from oracle import c_oracle

pytestmark = pytest.mark.gpu


def _sample_check(workload, n_sample, k=None, seed=0, check_props=True):
    import torch
    from mojo_bm25_b200 import engine, synth

    idx, q, k0 = synth.make_workload(workload, device="cuda", index_seed=seed)
    k = k or k0
    index = engine.DeviceIndex.from_torch(idx.indptr, idx.indices, idx.data, idx.n_docs)
    ids, sc = index.search_device(q, k)
    torch.cuda.synchronize()
    ids, sc, qn = ids.cpu().numpy(), sc.cpu().numpy(), q.cpu().numpy()
    if check_props:
        assert np.all(sc[:, :-1] >= sc[:, 1:])
        same = sc[:, 1:] == sc[:, :-1]
        assert np.all(ids[:, 1:][same] > ids[:, :-1][same])  # ties by ascending doc id
        assert all(len(set(r.tolist())) == k for r in ids[:: max(1, len(ids) // 200)])
    indptr, indices, data = idx.numpy()
    step = max(1, len(qn) // n_sample)
    for i in range(0, len(qn), step)[:n_sample]:
        dense = c_oracle.scores_dense(indptr, indices, data, idx.n_docs, qn[i])
        orc.check_topk_against_dense(ids[i], sc[i], dense, k, exact=True)
    return index, idx, q, ids, sc


def test_config_e_full_size_k1000():
    """config E: 1M docs, 1000 queries x 64 terms incl. stop-word-length lists, top-1000."""
    _sample_check("E", 16)


def test_config_c_full_size_k100():
    """config C: 8.8M passages, 10k queries x ~6 terms, top-100."""
    _sample_check("C", 16)


def test_10m_target_config_full_size_k100_and_k10_prefix():
    import torch

    index, idx, q, ids, sc = _sample_check("10M", 12)
    ids10, sc10 = index.search_device(q, 10)
    torch.cuda.synchronize()
    assert np.array_equal(ids10.cpu().numpy(), ids[:, :10]) and np.array_equal(sc10.cpu().numpy(), sc[:, :10])
    # the same handle compressed (4-byte postings, bf16 weights): the oracle on the rounded matrix
    from mojo_bm25_b200 import engine

    index.compress("bf16")
    cids, csc = index.search_device(q, 100)
    torch.cuda.synchronize()
    cids, csc, qn = cids.cpu().numpy(), csc.cpu().numpy(), q.cpu().numpy()
    assert np.all(csc[:, :-1] >= csc[:, 1:])
    indptr, indices, data = idx.numpy()
    data_q = engine.round_to_bf16(data)
    for i in range(0, len(qn), len(qn) // 8)[:8]:
        dense = c_oracle.scores_dense(indptr, indices, data_q, idx.n_docs, qn[i])
        orc.check_topk_against_dense(cids[i], csc[i], dense, 100, exact=True)


def test_config_d_one_full_shard_with_doc_id_base():
    """config D: one full 12.5M-document shard (seed = shard number 3, doc_id_base = 3 * 12.5M)."""
    import torch
    from mojo_bm25_b200 import engine, synth

    idx, q, k = synth.make_workload("D", device="cuda", index_seed=3)
    base = 3 * idx.n_docs
    index = engine.DeviceIndex.from_torch(idx.indptr, idx.indices, idx.data, idx.n_docs, doc_id_base=base)
    q = q[:512].contiguous()
    ids, sc = index.search_device(q, k)
    torch.cuda.synchronize()
    ids, sc, qn = ids.cpu().numpy(), sc.cpu().numpy(), q.cpu().numpy()
    assert ids.min() >= base and ids.max() < base + idx.n_docs
    indptr, indices, data = idx.numpy()
    for i in range(0, 512, 64):
        dense = c_oracle.scores_dense(indptr, indices, data, idx.n_docs, qn[i])
        orc.check_topk_against_dense(ids[i] - base, sc[i], dense, k, exact=True)
