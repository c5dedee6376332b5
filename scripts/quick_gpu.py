"""Developer timing probe (not the bench): per-kernel device times for named workloads."""
import argparse
import json
import sys
import os

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from mojo_bm25_b200 import engine, synth

ap = argparse.ArgumentParser()
ap.add_argument("--workloads", default="B")
ap.add_argument("--configs", default="default", help="comma list of option sets, each 'name=value:name=value' "
                "(bm25_index_set_option names: tile_docs, splits, consumer_warps, cap, waves, no_theta_share)")
ap.add_argument("--iters", type=int, default=5)
ap.add_argument("--scale", type=float, default=1.0)
ap.add_argument("--k", type=int, default=0, help="override top-k")
ap.add_argument("--compress", action="store_true", help="bf16 / 4-byte packed postings")
ap.add_argument("--sort-queries", action="store_true", help="order the batch by heaviest (lowest-id) term")
args = ap.parse_args()

flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for wl in args.workloads.split(","):
    idx, q, k = synth.make_workload(wl, device="cuda", scale=args.scale)
    index = engine.DeviceIndex.from_torch(idx.indptr, idx.indices, idx.data, idx.n_docs)
    index.set_option("timing", 1)
    if args.k:
        k = args.k
    if args.compress:
        index.compress()
    if args.sort_queries:
        key = torch.where(q >= 0, q, torch.full_like(q, 1 << 30)).min(dim=1).values
        q = q[torch.argsort(key, stable=True)].contiguous()
    qn = q.cpu().numpy()
    pbytes = index.posting_bytes(qn, 0)
    print(f"# {wl}: docs={idx.n_docs} terms={idx.n_terms} nnz={idx.nnz} Q={q.shape[0]} T={q.shape[1]} k={k} "
          f"posting_bytes={pbytes/1e9:.3f} GB", flush=True)
    for cfg in args.configs.split(","):
        opts = dict(tile_docs=0, splits=0, consumer_warps=0, cap=0, waves=0, no_theta_share=0, no_priming=0, no_hot=0, heavy_min=0, cand_smem=0, no_bulk_clear=0, no_query_sort=0, generic_kernel=0, q_major=0, no_epoch=0, no_packed=0)
        if cfg != "default":
            for kv in cfg.split(":"):
                name, val = kv.split("=")
                opts[name] = int(val)
        for name, val in opts.items():
            try:
                index.set_option(name, val)
            except ValueError:
                if val != 0:  # an older library build (A/B runs) may not know a knob left at its default
                    raise
        times = []
        for it in range(args.iters + 2):
            flush.zero_()
            index.search_device(q, k)
            times.append(index.last_timing_ms())
        t = np.array(times[2:])
        seg, score, merge = np.median(t, axis=0)
        tot = seg + score + merge
        print(json.dumps(dict(workload=wl, cfg=cfg, tile_docs=index.info.tile_docs, seg_ms=round(float(seg), 4),
                              score_ms=round(float(score), 4), merge_ms=round(float(merge), 4),
                              qps=round(q.shape[0] / tot * 1e3, 1), score_GBps=round(pbytes / score / 1e6, 1))), flush=True)
    index.close()
    del idx, index
    torch.cuda.empty_cache()
