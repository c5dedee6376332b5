"""Index-build timing: tokens/s of the GPU builder next to its numpy
restatement on the host.  python scripts/bench_index_build.py [--docs 1000000] [--terms 100000] [--len 40]"""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from mojo_bm25_b200 import index_build
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from _index_build_ref import build_csc_reference_numpy

ap = argparse.ArgumentParser()
ap.add_argument("--docs", type=int, default=1_000_000)
ap.add_argument("--terms", type=int, default=100_000)
ap.add_argument("--len", type=int, default=40)
ap.add_argument("--cpu-docs", type=int, default=100_000, help="documents of the host sample")
a = ap.parse_args()
g = torch.Generator(device="cuda").manual_seed(0)
lens = torch.poisson(torch.full((a.docs,), float(a.len), device="cuda"), generator=g).long()
ptr = torch.zeros(a.docs + 1, dtype=torch.int64, device="cuda")
torch.cumsum(lens, 0, out=ptr[1:])
n_tok = int(ptr[-1])
u = torch.rand(n_tok, device="cuda", generator=g, dtype=torch.float64)
tok = (torch.floor((a.terms + 1.0) ** u) - 1).clamp_(0, a.terms - 1).to(torch.int32)  # Zipf(1) term ids
for _ in range(2):
    out = index_build.build_csc(tok, ptr, a.terms, device="cuda")
torch.cuda.synchronize()
t0 = time.perf_counter()
out = index_build.build_csc(tok, ptr, a.terms, device="cuda")
torch.cuda.synchronize()
gpu_s = time.perf_counter() - t0
nc = int(ptr[a.cpu_docs])
tok_h, ptr_h = tok[:nc].cpu().numpy(), ptr[: a.cpu_docs + 1].cpu().numpy()
t0 = time.perf_counter()
ref = build_csc_reference_numpy(tok_h, ptr_h, a.terms)
cpu_s = time.perf_counter() - t0
print(json.dumps({"metric": "index build tokens/s", "gpu_tokens_per_s": n_tok / gpu_s, "gpu_seconds": gpu_s,
                  "tokens": n_tok, "docs": a.docs, "postings": int(out[1].numel()),
                  "cpu_numpy_tokens_per_s": nc / cpu_s, "cpu_sample_docs": a.cpu_docs, "cpu_seconds": cpu_s}))
