import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mojo_bm25_b200 import engine, synth
idx, q, k = synth.make_workload("B", device="cuda")
index = engine.DeviceIndex.from_torch(idx.indptr, idx.indices, idx.data, idx.n_docs)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for kk in (10, 100, 10, 100):
    index.set_option("timing", 1)
    ts = []
    for i in range(12):
        flush.zero_(); index.search_device(q, kk); ts.append(round(index.last_timing_ms()[1], 4))
    print("k", kk, "sync each step:", ts)
    index.set_option("timing", 0)
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(12)]
    for a, b in ev:
        flush.zero_(); a.record(); index.search_device(q, kk); b.record()
    torch.cuda.synchronize()
    print("k", kk, "queued steps  :", [round(a.elapsed_time(b), 4) for a, b in ev])
