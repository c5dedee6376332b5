"""Seeded synthetic corpora and query batches of the shapes named in BASELINE.json.

"Direct CSC synthesis" (SURVEY.md section 8d): the index is generated column by column without
ever materialising documents.  Term rank r (0-based; term id == rank, id 0 = heaviest list) has
document frequency  df_r = round(N * (1 - exp(-c / (r+1)^s)))  with c solved so that
sum_r df_r = N * U (a saturating Zipf law: stop words approach df = N).  The doc ids of a column
are a sorted STRATIFIED-uniform sample of size df_r from [0, N): posting i is drawn uniformly from
the i-th of df_r equal strata, which yields strictly increasing ids in O(nnz) fully vectorised
work (on the GPU for the large configs).  NOTE: this is more regular than the uniform sample
without replacement SURVEY.md 8d names -- every document tile receives the same number of postings
of a term (+-1), the friendliest case for a document-range-tiled kernel.  ``clustered=True`` is
the hard counterpart: every term gets its own bursty density over the document range (64 segments
with Gamma(0.3)-distributed weights, so a few segments hold most of the term's postings and many
hold almost none); DESIGN.md reports both.  Weights follow the bm25s "lucene" formula of the
bundled index (reference animal_index_bm25/data.csc.index.npy):
    w = idf_r * tf / (tf + k1 * (1 - b + b * dl_d / avgdl)),  idf_r = ln((N - df_r + .5)/(df_r + .5) + 1)
with dl_d ~ clip(round(lognormal(ln 1.5U, 0.5)), 8, 512) and tf ~ geometric(p = 0.7).

Everything is torch so the same code runs on CPU tensors (tests) and CUDA tensors (bench).
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Optional

import numpy as np
import torch


@dataclass
class SynthIndex:
    indptr: torch.Tensor   # int32 [V+1]
    indices: torch.Tensor  # int32 [nnz]
    data: torch.Tensor     # float32 [nnz]
    n_docs: int
    n_terms: int

    @property
    def nnz(self) -> int:
        return int(self.indices.numel())

    def numpy(self):
        return (self.indptr.cpu().numpy(), self.indices.cpu().numpy(), self.data.cpu().numpy())


def zipf_doc_freqs(n_docs: int, n_terms: int, mean_unique: float, s: float = 1.0) -> np.ndarray:
    """df_r of the saturating Zipf model, int64 [V]."""
    w = np.arange(1, n_terms + 1, dtype=np.float64) ** (-s)
    target = float(mean_unique)
    lo, hi = 0.0, 1.0
    while np.sum(1.0 - np.exp(-hi * w)) < target:
        hi *= 2.0
    for _ in range(200):
        mid = 0.5 * (lo + hi)
        if np.sum(1.0 - np.exp(-mid * w)) < target:
            lo = mid
        else:
            hi = mid
    c = 0.5 * (lo + hi)
    df = np.rint(n_docs * (1.0 - np.exp(-c * w))).astype(np.int64)
    return np.clip(df, 0, n_docs)


def _clustered_quantiles(q: torch.Tensor, col: torch.Tensor, seed: int, n_seg: int = 64) -> torch.Tensor:
    """Map the stratified quantiles q in [0,1) of the postings (term ``col``) through a per-term
    piecewise-linear inverse CDF over ``n_seg`` equal document segments whose weights are
    Gamma(0.3) draws, hashed from (seed, term, segment) so that no per-term table is stored."""
    dev = q.device
    seg = torch.arange(n_seg, device=dev, dtype=torch.int64)

    def weights(c):  # [n, n_seg] float64, deterministic in (seed, c, seg)
        h = (c[:, None] * 1_000_003 + seg[None, :] * 7_919 + seed * 104_729 + 12_345) % 2_147_483_647
        u = ((h * 48_271) % 2_147_483_647).to(torch.float64) / 2_147_483_647.0
        return torch.clamp(u, 1e-9, 1.0) ** (1.0 / 0.3) + 1e-4  # ~ Gamma(0.3)-like: heavy mass near 0

    uniq, inv = torch.unique(col, return_inverse=True)
    w = weights(uniq)
    cdf = torch.cumsum(w / w.sum(dim=1, keepdim=True), dim=1)  # [n_uniq, n_seg]
    cdf[:, -1] = 1.0
    rows = cdf[inv]  # [n, n_seg]
    k = torch.clamp((rows < q[:, None]).sum(dim=1), max=n_seg - 1)  # segment whose CDF range holds q
    hi = rows.gather(1, k[:, None])[:, 0]
    lo = torch.where(k > 0, rows.gather(1, torch.clamp(k - 1, min=0)[:, None])[:, 0], torch.zeros_like(hi))
    frac = torch.clamp((q - lo) / torch.clamp(hi - lo, min=1e-18), 0.0, 1.0)
    return (k.to(torch.float64) + frac) / n_seg


def synth_index(n_docs: int, n_terms: int, mean_unique: float, s: float = 1.0, seed: int = 0,
                device: str = "cpu", k1: float = 1.5, b: float = 0.75,
                chunk: int = 1 << 26, clustered: bool = False) -> SynthIndex:
    dev = torch.device(device)
    gen = torch.Generator(device=dev)
    gen.manual_seed(seed)
    df_np = zipf_doc_freqs(n_docs, n_terms, mean_unique, s)
    nnz = int(df_np.sum())
    if nnz >= 2 ** 31:
        raise ValueError(f"nnz={nnz} overflows int32 indptr; shard the corpus by document range")
    df = torch.from_numpy(df_np).to(dev)
    indptr64 = torch.zeros(n_terms + 1, dtype=torch.int64, device=dev)
    torch.cumsum(df, 0, out=indptr64[1:])
    mu = math.log(1.5 * mean_unique)
    dl = torch.exp(mu + 0.5 * torch.randn(n_docs, generator=gen, device=dev, dtype=torch.float32))
    dl = torch.clamp(torch.round(dl), 8.0, 512.0)
    avgdl = float(dl.double().mean().item()) if n_docs > 0 else 1.0
    norm = (k1 * (1.0 - b + b * dl / avgdl)).to(torch.float32)  # [N]
    dfd = df.to(torch.float64)
    idf = torch.log((n_docs - dfd + 0.5) / (dfd + 0.5) + 1.0).to(torch.float32)  # [V]
    indices = torch.empty(nnz, dtype=torch.int32, device=dev)
    data = torch.empty(nnz, dtype=torch.float32, device=dev)
    log1mp = math.log(1.0 - 0.7)
    for start in range(0, nnz, chunk):
        end = min(nnz, start + chunk)
        pos = torch.arange(start, end, dtype=torch.int64, device=dev)
        col = torch.searchsorted(indptr64, pos, right=True) - 1
        i = pos - indptr64[col]
        dfc = df[col]
        lo = (i * n_docs) // dfc
        hi = ((i + 1) * n_docs) // dfc
        u = torch.rand(end - start, generator=gen, device=dev, dtype=torch.float64)
        doc = lo + torch.clamp((u * (hi - lo).to(torch.float64)).to(torch.int64), max=(hi - lo - 1))
        if clustered:
            # bursty terms (df <= N/4; denser terms stay stratified): quantile -> per-term warped position,
            # then made strictly increasing inside every term (ids are unique per term)
            sub = chunk_sub = 1 << 20
            for s0 in range(0, end - start, sub):
                sl = slice(s0, min(end - start, s0 + sub))
                sparse = dfc[sl] * 4 <= n_docs
                if bool(sparse.any()):
                    qq = (i[sl].to(torch.float64) + u[sl]) / dfc[sl].to(torch.float64)
                    pos = _clustered_quantiles(qq[sparse], col[sl][sparse], seed)
                    dsl = doc[sl]
                    dsl[sparse] = torch.clamp((pos * n_docs).to(torch.int64), max=n_docs - 1)
                    doc[sl] = dsl
            big = int(n_docs) + int(nnz) + 1
            e = doc - i + col * big                       # strictly increasing ids: d_i = max_{j<=i}(d_j - j) + i per term
            doc = torch.cummax(e, dim=0).values - col * big + i
            over = doc - (n_docs - dfc + i)               # ... and never beyond N - (df - i): shift the tail back
            doc = doc - torch.clamp(over, min=0)
        u2 = torch.rand(end - start, generator=gen, device=dev, dtype=torch.float32).clamp_(min=1e-12)
        tf = 1.0 + torch.floor(torch.log(u2) / log1mp)
        wgt = idf[col] * tf / (tf + norm[doc])
        indices[start:end] = doc.to(torch.int32)
        data[start:end] = wgt
        del pos, col, i, dfc, lo, hi, u, doc, u2, tf, wgt
    return SynthIndex(indptr64.to(torch.int32), indices, data, int(n_docs), int(n_terms))


def _zipf_ranks(n: int, lo: int, hi: int, gen, dev) -> torch.Tensor:
    """n iid ranks in [lo, hi) with P(r) ~ 1/(r+1) (log-uniform inverse CDF)."""
    u = torch.rand(n, generator=gen, device=dev, dtype=torch.float64)
    a, bb = float(lo + 1), float(hi + 1)
    r = torch.floor(a * (bb / a) ** u).to(torch.int64) - 1
    return torch.clamp(r, lo, hi - 1)


def synth_queries(n_terms: int, n_queries: int, n_query_terms: int, r0: int = 8, seed: int = 1,
                  device: str = "cpu", poisson_mean: Optional[float] = None, max_terms: int = 16,
                  heavy_terms: int = 0, heavy_range: int = 100) -> torch.Tensor:
    """int32 [Q,T] query matrix, -1 padded.

    * fixed length: ``n_query_terms`` Zipf(s=1) terms from ranks [r0, V)             (configs B, D)
    * ``poisson_mean``: T_q ~ clip(Poisson(mean), 1, max_terms), width = max_terms   (config C)
    * ``heavy_terms`` > 0: that many uniform terms from ranks [0, heavy_range) followed by
      ``n_query_terms - heavy_terms`` Zipf terms from [heavy_range, V)               (config E)
    """
    dev = torch.device(device)
    gen = torch.Generator(device=dev)
    gen.manual_seed(seed)
    r0 = min(r0, max(n_terms - 1, 0))
    if poisson_mean is not None:
        width = max_terms
        lens = torch.poisson(torch.full((n_queries,), float(poisson_mean), device=dev), generator=gen)
        lens = torch.clamp(lens, 1, max_terms).to(torch.int64)
        q = _zipf_ranks(n_queries * width, r0, n_terms, gen, dev).view(n_queries, width)
        mask = torch.arange(width, device=dev)[None, :] >= lens[:, None]
        q[mask] = -1
        return q.to(torch.int32)
    if heavy_terms > 0:
        hr = min(heavy_range, n_terms)
        heavy = torch.randint(0, hr, (n_queries, heavy_terms), generator=gen, device=dev)
        light = _zipf_ranks(n_queries * (n_query_terms - heavy_terms), min(hr, n_terms - 1), n_terms, gen, dev)
        return torch.cat([heavy, light.view(n_queries, -1)], dim=1).to(torch.int32)
    return _zipf_ranks(n_queries * n_query_terms, r0, n_terms, gen, dev).view(n_queries, n_query_terms).to(torch.int32)


# Named workloads (BASELINE.json configs / BASELINE.md section 5)
WORKLOADS = {
    "B": dict(n_docs=1_000_000, n_terms=100_000, mean_unique=40, n_queries=1000, n_query_terms=4, k=10, r0=8),
    "C": dict(n_docs=8_800_000, n_terms=1_000_000, mean_unique=30, n_queries=10000, poisson_mean=6.0,
              max_terms=16, n_query_terms=16, k=100, r0=8),
    "10M": dict(n_docs=10_000_000, n_terms=1_000_000, mean_unique=30, n_queries=1000, n_query_terms=6, k=100, r0=8),
    "D": dict(n_docs=12_500_000, n_terms=1_000_000, mean_unique=30, n_queries=10000, poisson_mean=6.0,
              max_terms=16, n_query_terms=16, k=100, r0=8),  # per shard; 8 shards = 100M docs
    "E": dict(n_docs=1_000_000, n_terms=100_000, mean_unique=40, n_queries=1000, n_query_terms=64, k=1000,
              heavy_terms=8, heavy_range=100, r0=100),
    "tiny": dict(n_docs=20_000, n_terms=2_000, mean_unique=20, n_queries=64, n_query_terms=4, k=10, r0=8),
}


def make_workload(name: str, device: str = "cpu", index_seed: int = 0, query_seed: int = 1, scale: float = 1.0):
    """Returns (SynthIndex, queries int32 [Q,T], k) of a named workload (optionally scaled down).
    A trailing "c" (``"Bc"``, ``"10Mc"``) selects the clustered (bursty) variant of the index."""
    clustered = name.endswith("c") and name[:-1] in WORKLOADS
    if clustered:
        name = name[:-1]
    cfg = dict(WORKLOADS[name])
    n_docs = max(64, int(cfg["n_docs"] * scale))
    n_terms = max(16, int(cfg["n_terms"] * scale))
    idx = synth_index(n_docs, n_terms, cfg["mean_unique"], seed=index_seed, device=device, clustered=clustered)
    q = synth_queries(n_terms, cfg["n_queries"], cfg["n_query_terms"], r0=cfg.get("r0", 8), seed=query_seed,
                      device=device, poisson_mean=cfg.get("poisson_mean"), max_terms=cfg.get("max_terms", 16),
                      heavy_terms=cfg.get("heavy_terms", 0), heavy_range=cfg.get("heavy_range", 100))
    return idx, q, min(cfg["k"], n_docs)
