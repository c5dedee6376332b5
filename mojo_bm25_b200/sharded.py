"""Multi-GPU query path (new design; the reference is single-device).

One process per GPU (torch.distributed, NCCL over NVLink 5 / NVSwitch).

* ``DocShardedSearcher``: the index is partitioned by document range, rank g owning documents
  [base_g, base_g + n_g).  Every rank scores the SAME query batch against its shard, writes its
  local top-k straight into the send half of one packed all-gather buffer ([2][Q][k]: ids, then
  score bits), one ``all_gather_into_tensor`` moves Q*k*8 B per rank, and bm25_merge_topk selects
  the global top-k from the [W][2][Q][k] buffer on every rank.  Scores are comparable across
  shards because the weights are precomputed with global statistics at index time.
* ``QuerySplitSearcher``: the (small) index is replicated and the query batch is split across
  ranks -- no data-path collective except the optional gather of results.

The local-search and merge callables are injectable so that the host logic (partitioning,
packing, the collective) is exercised on CPU with the gloo backend in tests/.
"""
from __future__ import annotations

from typing import Callable, Optional, Tuple

import torch
import torch.distributed as dist


def doc_range_of_rank(n_docs: int, world: int, rank: int) -> Tuple[int, int]:
    """[lo, hi) of the documents rank owns: contiguous ranges of ceil(N / world)."""
    per = -(-int(n_docs) // int(world))
    lo = min(int(n_docs), rank * per)
    return lo, min(int(n_docs), lo + per)


class DocShardedSearcher:
    def __init__(self, local_search, k: int, merge: Optional[Callable] = None,
                 group: Optional[dist.ProcessGroup] = None, shard_docs=None):
        """``local_search(queries, k, out_ids, out_scores)`` fills the two [Q,k] outputs with the
        shard-local top-k (GLOBAL doc ids); pass a LIST of such callables when this rank owns
        several document shards (a corpus of more than 2^31 postings on few GPUs: every shard is
        its own int32-indexed handle).  Every rank must own the same number of shards.
        ``merge(ids_view, scores_view, k, list_stride, n_lists, n_queries, k_in)`` defaults to the
        CUDA merge kernel.  ``shard_docs[s]`` = number of documents of local shard ``s``: a shard
        with fewer than k documents (the tail shard of a small corpus, or an empty one) is searched
        with k_local = its size and its list is padded with (id -1, score -inf) entries, which the
        merge ignores -- the global k <= N stays valid although the reference's k <= n_docs rule
        (bm25_native.py:204-214) would reject the shard-local call."""
        self.local_searches = list(local_search) if isinstance(local_search, (list, tuple)) else [local_search]
        self.shard_docs = None if shard_docs is None else [int(n) for n in shard_docs]
        if self.shard_docs is not None and len(self.shard_docs) != len(self.local_searches):
            raise ValueError("shard_docs must name one size per local shard")
        self.k = int(k)
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        if merge is None:
            from .engine import merge_topk_device

            merge = merge_topk_device
        self.merge = merge
        self._send = None
        self._recv = None
        self.timing = False  # record CUDA events around local searches / all-gather / merge
        self._ev = None

    @classmethod
    def from_index(cls, index, k: int, group=None):
        """``index``: one DeviceIndex, or the list of DeviceIndex shards this rank owns."""
        idxs = list(index) if isinstance(index, (list, tuple)) else [index]

        def make(ix):
            return lambda q, kk, oi, os_: ix.search_device(q, kk, out_ids=oi, out_scores=os_)

        return cls([make(ix) for ix in idxs], k, group=group, shard_docs=[ix.n_docs for ix in idxs])

    def _buffers(self, n_queries: int, device):
        n_local = len(self.local_searches)
        shape = (n_local, 2, n_queries, self.k)
        if self._send is None or self._send.shape != shape or self._send.device != device:
            self._send = torch.empty(shape, dtype=torch.int32, device=device)
            self._recv = torch.empty((self.world,) + shape, dtype=torch.int32, device=device)
        return self._send, self._recv

    def search(self, queries: torch.Tensor):
        """queries int32 [Q,T] (identical on all ranks) -> global (ids [Q,k], scores [Q,k])."""
        n_queries = queries.shape[0]
        n_local = len(self.local_searches)
        send, recv = self._buffers(n_queries, queries.device)
        ev = None
        if self.timing and queries.is_cuda:
            if self._ev is None:
                self._ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
            ev = self._ev
            ev[0].record()
        for s, local in enumerate(self.local_searches):
            k_local = self.k if self.shard_docs is None else min(self.k, self.shard_docs[s])
            if k_local == self.k:
                local(queries, self.k, send[s, 0], send[s, 1].view(torch.float32))
                continue
            send[s, 0].fill_(-1)
            send[s, 1].view(torch.float32).fill_(float("-inf"))
            if k_local > 0:
                ids = torch.empty((n_queries, k_local), dtype=torch.int32, device=queries.device)
                sc = torch.empty((n_queries, k_local), dtype=torch.float32, device=queries.device)
                local(queries, k_local, ids, sc)
                send[s, 0, :, :k_local].copy_(ids)
                send[s, 1, :, :k_local].view(torch.float32).copy_(sc)
        if ev:
            ev[1].record()
        if self.world > 1:
            dist.all_gather_into_tensor(recv.view(self.world * n_local * 2, n_queries, self.k),
                                        send.view(n_local * 2, n_queries, self.k), group=self.group)
        else:
            recv[0].copy_(send)
        if ev:
            ev[2].record()
        out = self.merge(recv[0, 0, 0], recv[0, 0, 1].view(torch.float32), self.k,
                         list_stride=2 * n_queries * self.k, n_lists=self.world * n_local,
                         n_queries=n_queries, k_in=self.k)
        if ev:
            ev[3].record()
        return out

    def last_timing_ms(self):
        """(local searches, all-gather [device copy at world 1], merge) device ms of the last search
        (needs ``timing = True``); waits for that search."""
        ev = self._ev
        ev[3].synchronize()
        return ev[0].elapsed_time(ev[1]), ev[1].elapsed_time(ev[2]), ev[2].elapsed_time(ev[3])


class QuerySplitSearcher:
    def __init__(self, local_search: Callable, k: int, group: Optional[dist.ProcessGroup] = None):
        self.local_search = local_search
        self.k = int(k)
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0

    @classmethod
    def from_index(cls, index, k: int, group=None):
        return cls(lambda q, kk, oi, os_: index.search_device(q, kk, out_ids=oi, out_scores=os_), k, group=group)

    def my_slice(self, n_queries: int) -> Tuple[int, int]:
        per = -(-n_queries // self.world)
        lo = min(n_queries, self.rank * per)
        return lo, min(n_queries, lo + per)

    def search(self, queries: torch.Tensor, gather: bool = True):
        """queries int32 [Q,T] (identical on all ranks).  Each rank scores its contiguous slice;
        with ``gather`` every rank receives the full [Q,k] result."""
        n_queries = queries.shape[0]
        per = -(-n_queries // self.world)
        lo, hi = self.my_slice(n_queries)
        out = torch.zeros((2, per, self.k), dtype=torch.int32, device=queries.device)
        if hi > lo:
            self.local_search(queries[lo:hi].contiguous(), self.k, out[0, : hi - lo], out[1, : hi - lo].view(torch.float32))
        if not gather or self.world == 1:
            return out[0, : hi - lo], out[1, : hi - lo].view(torch.float32)
        recv = torch.empty((self.world, 2, per, self.k), dtype=torch.int32, device=queries.device)
        dist.all_gather_into_tensor(recv.view(self.world * 2, per, self.k), out, group=self.group)
        ids = recv[:, 0].reshape(self.world * per, self.k)[:n_queries]
        scores = recv[:, 1].reshape(self.world * per, self.k)[:n_queries].view(torch.float32)
        return ids, scores
