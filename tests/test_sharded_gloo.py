"""Multi-process host logic of the sharded query path on CPU (gloo, world_size 2): document-range
partitioning, packed all-gather buffer, merge -- with the oracle injected as the local scorer."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import bm25_oracle as orc


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _make_case():
    import scipy.sparse as sp

    rng = np.random.default_rng(11)
    m = sp.random(1201, 50, density=0.15, format="csc", dtype=np.float32, random_state=np.random.RandomState(4),
                  data_rvs=lambda n: (0.1 + rng.random(n)).astype(np.float32))
    m.sort_indices()
    q = rng.integers(-1, 50, size=(10, 5)).astype(np.int32)
    return m.indptr.astype(np.int32), m.indices.astype(np.int32), m.data, 1201, q


def _oracle_merge(ids_view, scores_view, k, list_stride, n_lists, n_queries, k_in):
    # views of list 0 inside the packed [W][2][Q][k] buffer -> rebuild the dense [W,Q,k] arrays
    base_i = ids_view.flatten()
    flat = torch.as_strided(ids_view, (n_lists, n_queries, k_in), (list_stride, k_in, 1))
    flat_s = torch.as_strided(scores_view, (n_lists, n_queries, k_in), (list_stride, k_in, 1))
    i, s = orc.merge_topk_lists(flat.numpy().copy(), flat_s.numpy().copy(), k)
    return torch.from_numpy(i), torch.from_numpy(s)


def _worker(rank, world, port, mode, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from mojo_bm25_b200 import sharded

    indptr, indices, data, n_docs, q = _make_case()
    k = 20
    qt = torch.from_numpy(q)
    if mode == "doc":
        lo, hi = sharded.doc_range_of_rank(n_docs, world, rank)
        ptr, ind, dat, nd, base = orc.partition_csc_by_doc_range(indptr, indices, data, n_docs, world)[rank]
        assert (base, base + nd) == (lo, hi)

        def local(queries, kk, out_ids, out_scores):
            i, s = orc.search_csc(ptr, ind, dat, nd, queries.numpy(), kk)
            out_ids.copy_(torch.from_numpy(i + base))
            out_scores.copy_(torch.from_numpy(s))

        s = sharded.DocShardedSearcher(local, k, merge=_oracle_merge)
        ids, sc = s.search(qt)
    elif mode == "doc2":  # two document shards per rank (a fixed 4-shard corpus on 2 ranks)
        parts = orc.partition_csc_by_doc_range(indptr, indices, data, n_docs, 2 * world)

        def make(part):
            ptr, ind, dat, nd, base = part

            def local(queries, kk, out_ids, out_scores):
                i, s = orc.search_csc(ptr, ind, dat, nd, queries.numpy(), kk)
                out_ids.copy_(torch.from_numpy(i + base))
                out_scores.copy_(torch.from_numpy(s))

            return local

        s = sharded.DocShardedSearcher([make(parts[rank]), make(parts[rank + world])], k, merge=_oracle_merge)
        ids, sc = s.search(qt)
    else:
        def local(queries, kk, out_ids, out_scores):
            i, s = orc.search_csc(indptr, indices, data, n_docs, queries.numpy(), kk)
            out_ids.copy_(torch.from_numpy(i))
            out_scores.copy_(torch.from_numpy(s))

        s = sharded.QuerySplitSearcher(local, k)
        ids, sc = s.search(qt)
    ret[rank] = (ids.numpy().copy(), sc.numpy().copy())
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("mode", ["doc", "doc2", "query"])
def test_sharded_search_world2(mode):
    world = 2
    port = _free_port()
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, port, mode, ret), nprocs=world, join=True)
    indptr, indices, data, n_docs, q = _make_case()
    assert len(ret) == world
    for rank in range(world):
        ids, sc = ret[rank]
        assert ids.shape == (len(q), 20)
        for i in range(len(q)):
            dense = orc.scores_dense(indptr, indices, data, n_docs, q[i])
            orc.check_topk_against_dense(ids[i], sc[i], dense, 20, exact=True)
    assert np.array_equal(ret[0][0], ret[1][0]) or mode == "query"


def test_doc_ranges_cover_corpus():
    from mojo_bm25_b200 import sharded

    for n, w in [(10, 3), (100_000_000, 8), (7, 8), (0, 2)]:
        spans = [sharded.doc_range_of_rank(n, w, r) for r in range(w)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))


def test_short_and_empty_tail_shards_single_process():
    """A tail shard with fewer than k documents (or none) is searched with k_local = its size and
    padded with (id -1, score -inf) entries that the merge ignores (host logic, oracle injected)."""
    from mojo_bm25_b200 import sharded

    indptr, indices, data, n_docs, q = _make_case()
    keep = indices < 13  # a 13-document corpus cut into ceil(13/4) = 4-document ranges: 4, 4, 4, 1
    col_of = np.repeat(np.arange(len(indptr) - 1), np.diff(indptr))
    ptr = np.zeros(len(indptr), np.int32)
    np.cumsum(np.bincount(col_of[keep], minlength=len(indptr) - 1), out=ptr[1:])
    ind, dat = indices[keep], data[keep]
    for n_shards, k in [(4, 4), (4, 13), (7, 5)]:
        parts = orc.partition_csc_by_doc_range(ptr, ind, dat, 13, n_shards)

        def make(part):
            p, i_, d, nd, base = part

            def local(queries, kk, out_ids, out_scores):
                assert 0 < kk <= nd
                i, s = orc.search_csc(p, i_, d, nd, queries.numpy(), kk)
                out_ids.copy_(torch.from_numpy(i + base))
                out_scores.copy_(torch.from_numpy(s))

            return local

        s = sharded.DocShardedSearcher([make(p) for p in parts], k, merge=_oracle_merge, shard_docs=[p[3] for p in parts])
        ids, sc = s.search(torch.from_numpy(q))
        for i in range(len(q)):
            dense = orc.scores_dense(ptr, ind, dat, 13, q[i])
            orc.check_topk_against_dense(ids[i].numpy(), sc[i].numpy(), dense, k, exact=True)
