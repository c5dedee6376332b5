"""Developer probe: run the sparse (or dense) half of a workload's batch alone (for ncu)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mojo_bm25_b200 import engine, synth
wl, which = sys.argv[1], sys.argv[2]
idx, q, k = synth.make_workload(wl, device="cuda")
index = engine.DeviceIndex.from_torch(idx.indptr, idx.indices, idx.data, idx.n_docs)
df = (idx.indptr[1:] - idx.indptr[:-1]).float()
dens = torch.where(q >= 0, df[q.clamp(min=0).long()], torch.zeros_like(q, dtype=torch.float32)).sum(1) / idx.n_docs
order = torch.argsort(dens)
n = len(order)
sel = {"sparse": order[: n // 2], "dense": order[n // 2:], "sparsest": order[: n // 4]}[which]
qq = q[sel].contiguous()
for _ in range(4):
    index.search_device(qq, k)
torch.cuda.synchronize()
print("done", wl, which, len(qq))
