"""CPU baseline runner -- TEST/BENCH INFRASTRUCTURE, NOT PRODUCT CODE.

Times the oracle port of the reference's CPU path (``BM25v.search``, reference
bm25_native.py:76-158, restated in oracle/bm25_oracle.py) on this host's cores.  The reference
itself is single-threaded Python/scipy; to use every host core the query sample is split into
contiguous chunks that run in forked worker processes over the same (read-only) CSC arrays.

Runs as a separate process (no CUDA context) when called from bench.py:

    python -m oracle.cpu_baseline --dir /dev/shm/xyz --k 10 --procs 8 --steps 3 --warmup 1

where --dir holds indptr.npy / indices.npy / data.npy / queries.npy / meta.json.
Prints one JSON line: {"qps": ..., "cores": ..., "n_queries": ..., "step_seconds": [...]}.
"""
import argparse
import json
import multiprocessing as mp
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import bm25_oracle as orc  # noqa: E402

_G = {}


def _work(args):
    lo, hi, k = args
    return orc.search_csc(_G["indptr"], _G["indices"], _G["data"], _G["n_docs"], _G["queries"][lo:hi], k)


def run(indptr, indices, data, n_docs, queries, k, procs, steps, warmup):
    """Returns (list of per-step seconds, (ids, scores) of the last step)."""
    _G.update(indptr=indptr, indices=indices, data=data, n_docs=n_docs, queries=queries)
    n = len(queries)
    procs = max(1, min(procs, n))
    bounds = np.linspace(0, n, procs + 1).astype(int)
    tasks = [(int(bounds[i]), int(bounds[i + 1]), k) for i in range(procs) if bounds[i + 1] > bounds[i]]
    times, last = [], None
    if procs == 1:
        for s in range(warmup + steps):
            t0 = time.perf_counter()
            last = [_work(tasks[0])]
            if s >= warmup:
                times.append(time.perf_counter() - t0)
    else:
        with mp.get_context("fork").Pool(procs) as pool:
            for s in range(warmup + steps):
                t0 = time.perf_counter()
                last = pool.map(_work, tasks)
                if s >= warmup:
                    times.append(time.perf_counter() - t0)
    ids = np.concatenate([r[0] for r in last])
    scores = np.concatenate([r[1] for r in last])
    return times, (ids, scores)


def host_cores() -> int:
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--dir", required=True)
    ap.add_argument("--k", type=int, required=True)
    ap.add_argument("--procs", type=int, default=0)
    ap.add_argument("--steps", type=int, default=1)
    ap.add_argument("--warmup", type=int, default=0)
    a = ap.parse_args()
    meta = json.load(open(os.path.join(a.dir, "meta.json")))
    arrs = {n: np.load(os.path.join(a.dir, n + ".npy"), mmap_mode="r") for n in ("indptr", "indices", "data", "queries")}
    procs = a.procs or host_cores()
    queries = np.ascontiguousarray(arrs["queries"])
    times, _ = run(arrs["indptr"], arrs["indices"], arrs["data"], meta["n_docs"], queries, a.k, procs, a.steps, a.warmup)
    mean = float(np.mean(times))
    print(json.dumps(dict(qps=len(queries) / mean, cores=min(procs, len(queries)), n_queries=len(queries),
                          step_seconds=times)))


if __name__ == "__main__":
    main()
