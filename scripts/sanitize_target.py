"""Tiny hot-path run for compute-sanitizer (one tool per gpurun call):
   compute-sanitizer --tool racecheck python scripts/sanitize_target.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

from mojo_bm25_b200 import engine, synth

idx, q, k = synth.make_workload("tiny")
indptr, indices, data = idx.numpy()
index = engine.DeviceIndex(indptr, indices, data, idx.n_docs)
qn = q.numpy()[:16]
ref = None
for opts in [dict(), dict(cap=k + 64, consumer_warps=4, tile_docs=512), dict(no_hot=1, no_priming=1, splits=3)]:
    for n in ("cap", "consumer_warps", "tile_docs", "no_hot", "no_priming", "splits"):
        index.set_option(n, opts.get(n, 0))
    ids, sc = index.search(qn, k)
    ids100, _ = index.search(qn, 300)
    if ref is None:
        ref = (ids, sc)
    assert np.array_equal(ids, ref[0]) and np.array_equal(sc.view(np.uint32), ref[1].view(np.uint32))
print("sanitize target ok")
