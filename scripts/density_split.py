"""Developer probe: score-kernel time of the sparse and the dense half of a workload's batch."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from mojo_bm25_b200 import engine, synth
wl = sys.argv[1] if len(sys.argv) > 1 else "10M"
idx, q, k = synth.make_workload(wl, device="cuda")
index = engine.DeviceIndex.from_torch(idx.indptr, idx.indices, idx.data, idx.n_docs)
index.set_option("timing", 1)
df = (idx.indptr[1:] - idx.indptr[:-1]).float()
dens = torch.where(q >= 0, df[q.clamp(min=0).long()], torch.zeros_like(q, dtype=torch.float32)).sum(1) / idx.n_docs
order = torch.argsort(dens)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for name, sel in [("all", order), ("sparse-half", order[: len(order) // 2]), ("dense-half", order[len(order) // 2:]),
                  ("sparsest-quarter", order[: len(order) // 4])]:
    qq = q[sel].contiguous()
    ts = []
    for it in range(6):
        flush.zero_(); index.search_device(qq, k); ts.append(index.last_timing_ms()[1])
    pb = index.posting_bytes(qq.cpu().numpy(), 0)
    t = float(np.median(ts[2:]))
    print(json.dumps(dict(workload=wl, subset=name, n=len(qq), score_ms=round(t, 4), GBps=round(pb / t / 1e6, 1),
                          mean_density=round(float(dens[sel].mean()), 4), max_density=round(float(dens[sel].max()), 4))))
