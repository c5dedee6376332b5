"""GPU tests of the document-sharded query path (the reference is single
device, main.py:205): real DeviceIndex shard handles -> packed [shards][2][Q][k] buffer ->
bm25_merge_topk with list_stride != 0, checked bitwise against the oracle on the UNSHARDED index.
World size 1 here (the packed buffer and the strided merge are identical at any world size; the
collective itself is covered on CPU by tests/test_sharded_gloo.py and on GPUs by bench.py's
`doc_shard.parity_checked`)."""
import numpy as np
import pytest

from oracle import bm25_oracle as orc
from oracle import c_oracle

pytestmark = pytest.mark.gpu


def _searcher(parts, k):
    from mojo_bm25_b200 import engine, sharded

    shards = [engine.DeviceIndex(ptr, ind, dat, n_docs=nd, doc_id_base=base) for ptr, ind, dat, nd, base in parts]
    return sharded.DocShardedSearcher.from_index(shards, k), shards


@pytest.mark.parametrize("workload,scale,n_shards,k", [("B", 0.04, 4, 50), ("D", 0.004, 8, 100), ("E", 0.02, 3, 1000)])
def test_packed_shard_buffer_and_strided_merge_match_the_oracle(workload, scale, n_shards, k):
    import torch
    from mojo_bm25_b200 import synth

    idx, q, _ = synth.make_workload(workload, scale=scale)
    indptr, indices, data = idx.numpy()
    qn = q.numpy()[:24]
    k = min(k, idx.n_docs)
    parts = orc.partition_csc_by_doc_range(indptr, indices, data, idx.n_docs, n_shards)
    s, shards = _searcher(parts, k)
    assert s.world == 1 and len(s.local_searches) == n_shards
    ids, sc = s.search(torch.from_numpy(qn).cuda())
    ids2, sc2 = s.search(torch.from_numpy(qn).cuda())  # the buffers are reused
    torch.cuda.synchronize()
    assert torch.equal(ids, ids2) and torch.equal(sc, sc2)
    ids, sc = ids.cpu().numpy(), sc.cpu().numpy()
    for i in range(len(qn)):
        dense = c_oracle.scores_dense(indptr, indices, data, idx.n_docs, qn[i])
        orc.check_topk_against_dense(ids[i], sc[i], dense, k, exact=True)
    # identical to one unsharded handle (same deterministic tie rule: score desc, doc id asc)
    from mojo_bm25_b200 import engine

    whole = engine.DeviceIndex(indptr, indices, data, n_docs=idx.n_docs)
    wi, ws = whole.search(qn, k)
    assert np.array_equal(wi, ids) and np.array_equal(ws.view(np.uint32), sc.view(np.uint32))


def test_tail_shard_smaller_than_k_and_empty_shard():
    """ceil(N / shards) ranges leave a short (or empty) tail shard; the global k <= N stays valid:
    short shards are searched with k_local = their size and padded with (id -1, -inf) entries."""
    import scipy.sparse as sp
    import torch

    rng = np.random.default_rng(3)
    n_docs = 13
    m = sp.random(n_docs, 9, density=0.5, format="csc", dtype=np.float32, random_state=np.random.RandomState(1),
                  data_rvs=lambda n: (0.1 + rng.random(n)).astype(np.float32))
    m.sort_indices()
    indptr, indices, data = m.indptr.astype(np.int32), m.indices.astype(np.int32), m.data
    q = rng.integers(-1, 9, size=(6, 3)).astype(np.int32)
    for n_shards, k in [(4, 4), (4, 13), (7, 5)]:  # per = 4 -> sizes 4,4,4,1 ; per = 2 -> the 7th shard holds 1 doc
        parts = orc.partition_csc_by_doc_range(indptr, indices, data, n_docs, n_shards)
        assert min(p[3] for p in parts) < k
        s, shards = _searcher(parts, k)
        ids, sc = s.search(torch.from_numpy(q).cuda())
        torch.cuda.synchronize()
        ids, sc = ids.cpu().numpy(), sc.cpu().numpy()
        for i in range(len(q)):
            dense = c_oracle.scores_dense(indptr, indices, data, n_docs, q[i])
            orc.check_topk_against_dense(ids[i], sc[i], dense, k, exact=True)
    # N = 5 over 4 shards: the last range is empty
    parts = orc.partition_csc_by_doc_range(indptr[:4], indices[: indptr[3]] % 5, data[: indptr[3]], 5, 4)
    assert parts[-1][3] == 0


def test_two_searches_on_different_streams_share_one_workspace_safely():
    """The per-handle workspace is ordered across streams by an event (bm25_b200.h threading
    contract): interleaved searches on two streams return what serial searches return."""
    import torch
    from mojo_bm25_b200 import engine, synth

    idx, q, k = synth.make_workload("B", device="cuda", scale=0.1)
    index = engine.DeviceIndex.from_torch(idx.indptr, idx.indices, idx.data, idx.n_docs)
    qa, qb = q[:300].contiguous(), q[300:700].contiguous()
    ra = index.search_device(qa, k)
    rb = index.search_device(qb, k)
    torch.cuda.synchronize()
    ra = (ra[0].clone(), ra[1].clone())
    rb = (rb[0].clone(), rb[1].clone())
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    for _ in range(20):
        with torch.cuda.stream(s1):
            xa = index.search_device(qa, k, stream=s1.cuda_stream)
        with torch.cuda.stream(s2):
            xb = index.search_device(qb, k, stream=s2.cuda_stream)
        hid, hsc = index.search(qa.cpu().numpy(), k)  # host entry point runs on the handle's own stream
        torch.cuda.synchronize()
        assert torch.equal(xa[0], ra[0]) and torch.equal(xa[1], ra[1])
        assert torch.equal(xb[0], rb[0]) and torch.equal(xb[1], rb[1])
        assert np.array_equal(hid, ra[0].cpu().numpy())


def test_concurrent_host_threads_on_one_handle():
    """Several Python threads calling the host entry point of ONE handle (ctypes releases the GIL):
    the mutex + stream ordering serialise them; every thread gets the serial answer."""
    import threading
    from mojo_bm25_b200 import engine, synth

    idx, q, k = synth.make_workload("B", scale=0.05)
    indptr, indices, data = idx.numpy()
    qn = q.numpy()
    index = engine.DeviceIndex(indptr, indices, data, n_docs=idx.n_docs)
    parts = [qn[i::4] for i in range(4)]
    want = [index.search(p, k) for p in parts]
    errors = []

    def work(i):
        try:
            for _ in range(15):
                ids, sc = index.search(parts[i], k)
                assert np.array_equal(ids, want[i][0]) and np.array_equal(sc.view(np.uint32), want[i][1].view(np.uint32))
        except Exception as e:  # noqa: BLE001
            errors.append((i, repr(e)))

    threads = [threading.Thread(target=work, args=(i,)) for i in range(4)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errors, errors


def test_split_device_index_equals_one_handle():
    """An index too large for one handle's int32 posting offsets (int64 indptr) is held as several
    document-range handles on one device (engine.SplitDeviceIndex); forced here with a tiny limit."""
    from mojo_bm25_b200 import engine, synth

    idx, q, k = synth.make_workload("B", scale=0.03)
    indptr, indices, data = idx.numpy()
    qn = q.numpy()[:40]
    whole = engine.DeviceIndex(indptr, indices, data, n_docs=idx.n_docs)
    want = whole.search(qn, 25)
    split = engine.SplitDeviceIndex(indptr.astype(np.int64), indices.astype(np.int64), data, idx.n_docs,
                                    max_postings=idx.nnz // 5 + 7)
    assert len(split.parts) >= 5 and sum(p.n_docs for p in split.parts) == idx.n_docs
    got = split.search(qn, 25)
    assert np.array_equal(got[0], want[0]) and np.array_equal(got[1].view(np.uint32), want[1].view(np.uint32))
    got = split.search(qn, 3)  # another k: its own searcher
    assert np.array_equal(got[0], want[0][:, :3])
    with pytest.raises(ValueError):
        split.search(np.array([[idx.n_terms]], np.int32), 3)
    assert isinstance(engine.open_index(indptr, indices, data, idx.n_docs), engine.DeviceIndex)
