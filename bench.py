#!/usr/bin/env python
"""bench.py -- BM25 query hot-path benchmark (contract: see DESIGN.md "Measurement").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload B|C|10M|D|E|tiny]
                    [--mode auto|query-split|doc-shard] [--impl ours|reference] [--no-subrecords]

A *step* is one pass of the hot path (threshold priming + cursor starts -> query ordering -> score
accumulation + per-range top-k -> merge) over one batch of synthetic queries.  Default workload =
BASELINE.json configs[1] ("B": 1M docs, 100k-term Zipf vocabulary, 1000 queries x 4 terms, top-10).
Rank 0 prints ONE JSON line.  `value` = whole-job queries/s with the index and the queries
resident in HBM (CUDA-event timed, L2 flushed between steps); `e2e` = the same metric through the
host-buffer C-ABI call (bm25_search_host: pinned H2D of the queries, kernels, D2H of ids+scores,
every step).  The same line carries sub-records measured in the same run so that the whole
BASELINE metric is on it:

    k100      workload B at k = 100
    wE        config E (1M docs, 1000 x 64 terms incl. stop-word-length lists, k = 1000)
    w10M      the 10M-document target index of north_star (1000 x 6 terms, k = 100), own roofline
    wC        config C (8.8M passages, 10k queries x ~6 terms, k = 100)           [N = 1 only]
    doc_shard config D: 100M documents as 8 document-range shards spread over the N GPUs (all 8 on
              one GPU at N = 1), local top-k + NCCL all-gather + merge kernel; strong scaling;
              local / all-gather / merge device ms and an oracle parity check of a query sample
"""
from __future__ import annotations

import argparse
import json
import os
import shutil
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "queries/sec"
UNIT = "queries/s"
DTYPE = "f32"


# ----------------------------------------------------------------------------------------------
def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="B")
    ap.add_argument("--mode", default="auto", choices=["auto", "query-split", "doc-shard"])
    ap.add_argument("--scale", type=float, default=1.0, help="shrink the workload (debug only)")
    ap.add_argument("--k", type=int, default=0, help="override top-k (debug only)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-subrecords", action="store_true", help="only the headline workload")
    ap.add_argument("--cpu-sample", type=int, default=0, help="queries in the CPU baseline sample")
    ap.add_argument("--doc-shard-scale", type=float, default=1.0, help="shrink config D (debug only)")
    a = ap.parse_args()
    a.warmup = max(a.warmup, 3) if a.impl == "ours" else a.warmup
    return a


def dist_env():
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    return rank, world, local


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            p = json.load(open(path))
            return float(p["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, burst copy)"
        except Exception:  # noqa: BLE001
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


# ----------------------------------------------------------------------------------------------
class ClockSampler:
    """Samples SM clock + throttle reasons of one GPU during the timed regions (NVML thread)."""

    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap",
               0x80: "hw_power_brake_slowdown", 0x2: "applications_clocks_setting", 0x100: "display_clock_setting"}

    def __init__(self, torch_device_index: int, period_s: float = 0.005):
        self.period = period_s
        self.samples, self.reason_bits, self.max_mhz = [], 0, None
        self._stop = threading.Event()
        self._thr = None
        self._h = None
        try:
            import pynvml
            import torch

            pynvml.nvmlInit()
            uuid = str(torch.cuda.get_device_properties(torch_device_index).uuid)
            if not uuid.startswith("GPU-"):
                uuid = "GPU-" + uuid
            try:
                self._h = pynvml.nvmlDeviceGetHandleByUUID(uuid.encode())
            except Exception:  # noqa: BLE001
                self._h = pynvml.nvmlDeviceGetHandleByIndex(torch_device_index)
            self._nv = pynvml
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self._h, pynvml.NVML_CLOCK_SM))
        except Exception:  # noqa: BLE001
            self._h = None

    def _loop(self):
        nv = self._nv
        while not self._stop.is_set():
            try:
                self.samples.append(float(nv.nvmlDeviceGetClockInfo(self._h, nv.NVML_CLOCK_SM)))
                self.reason_bits |= int(nv.nvmlDeviceGetCurrentClocksEventReasons(self._h))
            except Exception:  # noqa: BLE001
                pass
            self._stop.wait(self.period)

    def start(self):
        if self._h is not None:
            self._thr = threading.Thread(target=self._loop, daemon=True)
            self._thr.start()

    def stop(self):
        self._stop.set()
        if self._thr:
            self._thr.join()
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["unsampled"]}
        s = sorted(self.samples)
        reasons = [n for b, n in self.REASONS.items() if self.reason_bits & b]
        return {"sm_mhz": s[len(s) // 2], "sm_max_mhz": self.max_mhz, "reasons": reasons, "samples": len(s)}


# ----------------------------------------------------------------------------------------------
def dump_for_cpu(dirpath, indptr, indices, data, n_docs, queries):
    import numpy as np

    os.makedirs(dirpath, exist_ok=True)
    np.save(os.path.join(dirpath, "indptr.npy"), indptr)
    np.save(os.path.join(dirpath, "indices.npy"), indices)
    np.save(os.path.join(dirpath, "data.npy"), data)
    np.save(os.path.join(dirpath, "queries.npy"), queries)
    json.dump({"n_docs": int(n_docs)}, open(os.path.join(dirpath, "meta.json"), "w"))


def run_cpu_port(indptr, indices, data, n_docs, queries, k, steps, warmup):
    """Oracle port of BM25v.search on all host cores, in a separate CUDA-free process."""
    base = "/dev/shm" if os.path.isdir("/dev/shm") else None
    d = tempfile.mkdtemp(prefix="bm25_cpu_", dir=base)
    try:
        dump_for_cpu(d, indptr, indices, data, n_docs, queries)
        out = subprocess.run([sys.executable, "-m", "oracle.cpu_baseline", "--dir", d, "--k", str(k), "--steps",
                              str(steps), "--warmup", str(warmup)], cwd=ROOT, capture_output=True, text=True,
                             env={**os.environ, "CUDA_VISIBLE_DEVICES": "", "OMP_NUM_THREADS": "1"})
        if out.returncode != 0:
            raise RuntimeError("cpu baseline failed: " + out.stderr[-2000:])
        return json.loads(out.stdout.strip().splitlines()[-1])
    finally:
        shutil.rmtree(d, ignore_errors=True)


def cpu_sample_size(n_queries, n_docs, cores, override=0):
    if override:
        return min(n_queries, override)
    # ~25 ms per query per million documents per core for the numpy port; aim at ~10-20 s total
    per_q = 0.04 * max(n_docs, 1) / 1e6
    n = int(15.0 * cores / max(per_q, 1e-4))
    return max(cores, min(n_queries, n, 64 * cores))


def parallelism(mode, world, n_q, n_docs=0, total_shards=1, k=0):
    if mode == "single":
        return "1 GPU"
    if mode == "query-split":
        return (f"query-split x{world}: index replicated, each rank answers its own {n_q}-query batch, "
                "no data-path collective")
    per_rank = total_shards // world
    return (f"doc-shard x{world}: {total_shards} shards of {n_docs} docs ({n_docs * total_shards} total), "
            f"{per_rank} per rank, same batch on every rank, local top-k + NCCL all-gather of "
            f"{n_q * k * 8 * per_rank} B per rank + merge kernel")


def headline_mode(args, world):
    mode = args.mode
    if mode == "auto":
        mode = "query-split" if world > 1 and args.workload != "D" else ("doc-shard" if world > 1 else "single")
    if args.workload == "D" and mode != "query-split":
        return "doc-shard"  # the 100M-doc corpus only exists as 8 document shards, also on one GPU
    if world == 1:
        return "single"
    return mode


def workload_config(args, wl_name, n_docs, n_terms, nnz, n_q, width, k, r0, mode, world, total_shards):
    cfg = {
        "workload": f"{wl_name}: synthetic saturating-Zipf CSC index, {n_docs} docs, {n_terms} terms, "
                    f"nnz={nnz}; {n_q} queries x {width} term slots (r0={r0}), top-{k}",
        "n_docs": n_docs, "n_terms": n_terms, "nnz": nnz, "n_queries": n_q,
        "query_width": width, "k": int(k), "index_seed": 0, "query_seed": 1,
        "cache": "L2 flushed (256 MiB memset) between timed steps; index is larger than L2",
        "parallelism": parallelism(mode, world, n_q, n_docs, total_shards, k), "shards": total_shards,
    }
    if args.scale != 1.0:
        cfg["scale"] = args.scale
    return cfg


# ----------------------------------------------------------------------------------------------
def run_reference(args):
    """--impl reference: the reference's CPU implementation of the path (oracle port of
    BM25v.search, bm25_native.py:76-158; the Python reference itself cannot travel to the GPU box)
    on all host cores, each step a bounded query sample of the same workload."""
    rank, world, local = dist_env()
    if rank != 0:
        return
    import numpy as np
    import torch

    from mojo_bm25_b200 import synth
    from oracle import cpu_baseline

    wl_cfg = synth.WORKLOADS[args.workload]
    dev = "cuda" if torch.cuda.is_available() else "cpu"
    idx, q, k = synth.make_workload(args.workload, device=dev, scale=args.scale)
    if args.k:
        k = args.k
    indptr, indices, data = idx.numpy()
    qn = q.cpu().numpy()
    cores = cpu_baseline.host_cores()
    n_sample = cpu_sample_size(len(qn), idx.n_docs, cores, args.cpu_sample)
    # keep the whole run within minutes: shrink the per-step sample with the step count
    n_sample = max(cores, min(n_sample, int(n_sample * 8 / max(args.steps + args.warmup, 1)) or cores))
    res = run_cpu_port(indptr, indices, data, idx.n_docs, qn[:n_sample], k, args.steps, args.warmup)
    qps = res["qps"]
    sample = f"first {n_sample} of {len(qn)} queries per step, {res['cores']} forked workers"
    mode = headline_mode(args, args.gpus)
    total_shards = (8 if args.workload == "D" else args.gpus) if mode == "doc-shard" else 1
    line = {
        "impl": "reference", "metric": METRIC, "value": qps, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * float(np.mean(res["step_seconds"])), "higher_is_better": True,
        "scaling": "strong" if (mode == "doc-shard" and args.workload == "D") else "weak", "vs_baseline": None,
        "dtype": DTYPE, "data": "synthetic",
        "config": workload_config(args, args.workload, idx.n_docs, idx.n_terms, idx.nnz, int(q.shape[0]),
                                  int(q.shape[1]), k, wl_cfg.get("r0", 8), mode, args.gpus, total_shards),
        "cpu_baseline": {"value": qps, "unit": UNIT, "cores": res["cores"], "kind": "port", "sample": sample},
        "e2e": {"value": qps, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------------
class Ctx:
    """Per-process benchmark context: device, distributed handles, L2 flush buffer, peaks."""

    def __init__(self, args):
        import torch

        self.args = args
        self.rank, self.world, self.local = dist_env()
        torch.cuda.set_device(self.local)
        self.dev = torch.device("cuda", self.local)
        self.dist = None
        if self.world > 1:
            import torch.distributed as dist

            dist.init_process_group("nccl", device_id=self.dev)
            self.dist = dist
        self.flush = torch.empty(256 << 20, dtype=torch.uint8, device=self.dev)
        self.peak, self.peak_src = load_peaks()

    def barrier(self):
        if self.dist is not None:
            self.dist.barrier()

    def max_over_ranks(self, x: float) -> float:
        import torch

        if self.dist is None:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=self.dev)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())


def timed_steps(ctx, step, steps, warmup):
    """W untimed steps, then EXACTLY `steps` steps, each bracketed by CUDA events on the launching
    stream with the L2 flushed in between (the flush is outside the events); barrier + synchronize
    on both sides; returns (per-step ms list of this rank, MAX-over-ranks mean ms per step, wall s)."""
    import torch

    for _ in range(warmup):
        ctx.flush.zero_()
        step()
    torch.cuda.synchronize()
    starts = [torch.cuda.Event(enable_timing=True) for _ in range(steps)]
    ends = [torch.cuda.Event(enable_timing=True) for _ in range(steps)]
    ctx.barrier()
    torch.cuda.synchronize()
    wall0 = time.perf_counter()
    for i in range(steps):
        ctx.flush.zero_()
        starts[i].record()
        step()
        ends[i].record()
    torch.cuda.synchronize()
    ctx.barrier()
    wall = time.perf_counter() - wall0
    step_ms = [s.elapsed_time(e) for s, e in zip(starts, ends)]
    return step_ms, ctx.max_over_ranks(float(sum(step_ms))) / steps, wall


def kernel_split(ctx, indexes, step, steps):
    """Second, separate pass (its waits are NOT inside the timed loop of `value`): per-kernel device
    time from the library's own events on the launching stream -> median (segments + query order,
    score + top-k, merge) ms per step, summed over this rank's shards."""
    import numpy as np

    for ix in indexes:
        ix.set_option("timing", 1)
    rows = []
    for _ in range(max(3, min(steps, 10))):
        ctx.flush.zero_()
        step()
        rows.append(tuple(map(sum, zip(*[ix.last_timing_ms() for ix in indexes]))))
    for ix in indexes:
        ix.set_option("timing", 0)
    return np.median(np.array(rows), axis=0)  # median: one slow launch (e.g. a clock ramp after the sync) must not move it


def measured_traffic(workload):
    """DRAM bytes per launch of the score kernel from the committed ncu --set full capture of this
    workload (profiles/r2_traffic.json, written by scripts/ncu_traffic.py) -- reported only while the
    capture's source stamp matches the kernel sources this run was built from; otherwise None."""
    try:
        sys.path.insert(0, os.path.join(ROOT, "scripts"))
        import ncu_traffic

        t = json.load(open(os.path.join(ROOT, "profiles", "r2_traffic.json")))
        if t.get("kernel_source_sha256") != ncu_traffic.source_stamp() or workload not in t:
            return None, "no ncu capture of these kernel sources for this workload"
        return t[workload]["dram_read"] + t[workload]["dram_write"], "profiles/r2_traffic.json (ncu --set full capture of the same kernel sources)"
    except Exception as e:  # noqa: BLE001
        return None, f"unavailable ({type(e).__name__})"


def roofline_record(ctx, posting_bytes, km, batch_bytes, ms_per_step, traffic_key=None):
    score_ms = float(km[1])
    traffic, traffic_src = measured_traffic(traffic_key) if traffic_key else (None, "no capture for this record")
    achieved = posting_bytes / (score_ms * 1e-3) / 1e9
    return {
        "bound": "hbm", "kernel": "k_score_topk (score accumulation + per-range top-k)",
        "achieved": achieved, "peak": ctx.peak, "unit": "GB/s", "frac": achieved / ctx.peak,
        "frac_of_nominal_8000": achieved / 8000.0, "peak_source": ctx.peak_src,
        "algorithmic_bytes_per_launch": int(posting_bytes), "kernel_ms": score_ms,
        "kernel_share_of_step": score_ms / float(km.sum()),
        "other_kernels_ms": {"k_segments+k_query_order": float(km[0]), "k_merge": float(km[2])},
        "whole_batch_GBps": batch_bytes / (ms_per_step * 1e-3) / 1e9,
        # DRAM bytes per launch need a profiler: taken from the committed ncu capture of the same sources
        "traffic": traffic, "traffic_source": traffic_src,
    }


def e2e_host(ctx, index, qn, k, steps, units):
    """The call a user makes: bm25_search_host (pinned H2D of the queries + kernels + D2H of
    ids and scores + synchronise) every step."""
    import torch

    for _ in range(2):
        index.search(qn, k)
    ctx.barrier()
    total = 0.0
    for _ in range(steps):
        # the L2 flush is measurement apparatus: it has finished before the step's clock starts; the
        # call itself returns only after the results are in host memory (it synchronises its stream)
        ctx.flush.zero_()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        hid, hsc = index.search(qn, k)
        total += time.perf_counter() - t0
    e2e_s = ctx.max_over_ranks(total)
    return {"value": units / (e2e_s / steps), "unit": UNIT, "h2d_bytes_per_step": int(qn.nbytes),
            "d2h_bytes_per_step": int(hid.nbytes + hsc.nbytes), "api": "bm25_search_host via DeviceIndex.search"}


def measure_single(ctx, index, q, k, steps, warmup, want_e2e=True, traffic_key=None):
    """One index on this rank, this rank's batch `q`: value / roofline / e2e of one workload."""
    import torch

    qn = q.cpu().numpy()
    n_q = int(q.shape[0])
    out_ids = torch.empty((n_q, k), dtype=torch.int32, device=ctx.dev)
    out_sc = torch.empty((n_q, k), dtype=torch.float32, device=ctx.dev)

    def step():
        return index.search_device(q, k, out_ids=out_ids, out_scores=out_sc)

    step_ms, ms_per_step, wall = timed_steps(ctx, step, steps, warmup)
    km = kernel_split(ctx, [index], step, steps)
    posting_bytes = index.posting_bytes(qn, 0)  # 8 * sum df (score kernel)
    units = n_q * ctx.world  # query-split: every rank answers its own batch
    rec = {
        "value": units / (ms_per_step / 1e3), "ms_per_step": ms_per_step, "steps": steps,
        "roofline": roofline_record(ctx, posting_bytes, km, posting_bytes + 8 * k * n_q, ms_per_step, traffic_key),
        "step_ms_min": float(min(step_ms)), "step_ms_median": float(sorted(step_ms)[len(step_ms) // 2]),
        "wall_s_timed_region": wall,
    }
    if want_e2e:
        rec["e2e"] = e2e_host(ctx, index, qn, k, steps, units)
    return rec


def sub_workload(ctx, name, k=0, steps=10, warmup=3, index=None, idx=None, compress=False):
    """A sub-record: another named workload (or another k on an existing index), same measurement.
    compress: the handle is converted to the 4-byte posting format (bf16 weights) first; its roofline
    is then counted at 4 bytes per posting -- reported beside, never instead of, the int32/fp32 record."""
    from mojo_bm25_b200 import engine, synth

    args = ctx.args
    own = index is None
    if own:
        idx, q, k0 = synth.make_workload(name, device=str(ctx.dev), index_seed=0, query_seed=1 + ctx.rank, scale=args.scale)
        index = engine.DeviceIndex.from_torch(idx.indptr, idx.indices, idx.data, idx.n_docs)
        if compress:
            index.compress("bf16")
    else:
        cfg = synth.WORKLOADS[name]
        q = synth.synth_queries(idx.n_terms, cfg["n_queries"], cfg["n_query_terms"], r0=cfg.get("r0", 8),
                                seed=1 + ctx.rank, device=str(ctx.dev), poisson_mean=cfg.get("poisson_mean"),
                                max_terms=cfg.get("max_terms", 16), heavy_terms=cfg.get("heavy_terms", 0),
                                heavy_range=cfg.get("heavy_range", 100))
        k0 = min(cfg["k"], idx.n_docs)
    k = k or k0
    # the ncu captures are of the workloads' own k on one GPU (scripts/profile_all.sh)
    tkey = (name + ("_bf16" if compress else "")) if (k == k0 and ctx.args.scale == 1.0) else None
    rec = measure_single(ctx, index, q, k, steps, warmup, traffic_key=tkey)
    mode = "single" if ctx.world == 1 else "query-split"
    rec["config"] = workload_config(args, name, idx.n_docs, idx.n_terms, idx.nnz, int(q.shape[0]), int(q.shape[1]), k,
                                    synth.WORKLOADS[name].get("r0", 8), mode, ctx.world, 1)
    rec["metric"], rec["unit"] = METRIC, UNIT
    if compress:
        rec["posting_format"] = ("compressed: uint16 tile-local doc id + bf16 weight = 4 B per posting (weights rounded to "
                                 "bf16; results bit-identical to the reference on the rounded matrix, tests/test_gpu_compressed.py)")
        rec["roofline"]["bytes_per_posting"] = 4
        rec["roofline"]["frac_if_counted_at_8_bytes_per_posting"] = 2 * rec["roofline"]["frac"]
    if own:
        index.close()
    return rec


# ----------------------------------------------------------------------------------------------
def run_doc_shard(ctx, workload, steps, warmup, scale=1.0, parity_queries=4, want_cpu=False):
    """Config D (or --mode doc-shard on another workload): document-range shards, each its own
    int32-indexed handle synthesised with seed = shard number; rank r owns shards r, r+W, ...
    (strong scaling for D: a FIXED corpus of 8 shards x 12.5M docs and a fixed batch).  Every step:
    local searches into the packed send buffer, one NCCL all-gather, the merge kernel."""
    import numpy as np
    import torch

    from mojo_bm25_b200 import engine, sharded, synth

    world, rank, dev = ctx.world, ctx.rank, ctx.dev
    total_shards = 8 if workload == "D" else world
    if total_shards % world:
        raise SystemExit(f"workload D has 8 document shards; --gpus must divide 8 (got {world})")
    my_shards = list(range(rank, total_shards, world))
    indexes, synths, q, k = [], [], None, None
    for sh in my_shards:
        idx, q, k = synth.make_workload(workload, device=str(dev), index_seed=sh, query_seed=1, scale=scale)
        indexes.append(engine.DeviceIndex.from_torch(idx.indptr, idx.indices, idx.data, idx.n_docs,
                                                     doc_id_base=sh * idx.n_docs))
        synths.append(idx)
    if ctx.args.k:
        k = ctx.args.k
    n_q, n_docs = int(q.shape[0]), synths[0].n_docs
    qn = q.cpu().numpy()
    searcher = sharded.DocShardedSearcher.from_index(indexes, k)

    def step():
        return searcher.search(q)

    step_ms, ms_per_step, wall = timed_steps(ctx, step, steps, warmup)
    # second pass: where the step goes (device events; not inside the timed loop above)
    searcher.timing = True
    parts = []
    for _ in range(max(3, min(steps, 5))):
        ctx.flush.zero_()
        step()
        parts.append(searcher.last_timing_ms())
    searcher.timing = False
    local_ms, gather_ms, merge_ms = (ctx.max_over_ranks(float(x)) for x in np.median(np.array(parts), axis=0))
    km = kernel_split(ctx, indexes, step, steps)
    posting_bytes = sum(ix.posting_bytes(qn, 0) for ix in indexes)
    # e2e: pinned queries -> H2D -> local search(es) -> all-gather -> merge -> D2H of the result
    q_pin = torch.from_numpy(qn).pin_memory()
    q_dev = torch.empty_like(q)
    for w in range(2 + steps):
        if w == 2:
            torch.cuda.synchronize()
            ctx.barrier()
            t0 = time.perf_counter()
        ctx.flush.zero_()
        q_dev.copy_(q_pin, non_blocking=True)
        gi, gs = searcher.search(q_dev)
        hid, hsc = gi.cpu(), gs.cpu()
    e2e_s = ctx.max_over_ranks(time.perf_counter() - t0)
    # parity: a query sample against the C oracle on the same shard arrays (outside any timed region)
    from oracle import bm25_oracle as orc
    from oracle import c_oracle

    sample = list(range(0, n_q, max(1, n_q // parity_queries)))[:parity_queries]
    lists_i = np.zeros((len(my_shards), len(sample), k), np.int32)
    lists_s = np.zeros((len(my_shards), len(sample), k), np.float32)
    for s, (sh, idx) in enumerate(zip(my_shards, synths)):
        indptr, indices, data = idx.numpy()
        for j, qi in enumerate(sample):
            dense = c_oracle.scores_dense(indptr, indices, data, idx.n_docs, qn[qi])
            order = np.lexsort((np.arange(idx.n_docs), -dense.astype(np.float64)))[:k]
            lists_i[s, j], lists_s[s, j] = order + sh * idx.n_docs, dense[order]
        del indptr, indices, data
    if ctx.dist is not None:
        gathered = [None] * world
        ctx.dist.all_gather_object(gathered, (lists_i, lists_s))
        lists_i = np.concatenate([g[0] for g in gathered])
        lists_s = np.concatenate([g[1] for g in gathered])
    want_i, want_s = orc.merge_topk_lists(lists_i, lists_s, k)
    got_i, got_s = hid.numpy()[sample], hsc.numpy()[sample]
    parity_ok = bool(np.array_equal(got_i, want_i) and np.array_equal(got_s.view(np.uint32), want_s.view(np.uint32)))
    if not parity_ok:
        raise SystemExit("doc-shard parity check against the oracle FAILED")
    cpu = None
    if want_cpu and rank == 0:
        from oracle import cpu_baseline

        cores = cpu_baseline.host_cores()
        n_sample = cpu_sample_size(n_q, n_docs, cores, ctx.args.cpu_sample)
        indptr, indices, data = synths[0].numpy()
        res = run_cpu_port(indptr, indices, data, n_docs, qn[:n_sample], k, 1, 0)
        cpu = {"value": res["qps"] / total_shards, "unit": UNIT, "cores": res["cores"], "kind": "port",
               "sample": f"first {n_sample} of {n_q} queries, oracle port of BM25v.search, {res['cores']} forked workers, "
                         f"{res['step_seconds'][0]:.2f} s; timed on shard 0 of {total_shards} and divided by "
                         f"{total_shards} (a query visits every shard)"}
    rec = {
        "metric": METRIC, "unit": UNIT, "value": n_q / (ms_per_step / 1e3), "ms_per_step": ms_per_step, "steps": steps,
        "scaling": "strong" if workload == "D" else "weak", "n_gpus": world,
        "local_ms": local_ms, "all_gather_ms": gather_ms, "merge_ms": merge_ms,
        "all_gather_bytes_per_rank": int(n_q * k * 8 * len(my_shards)),
        "parity_checked": len(sample), "parity": "bit-exact vs C oracle (ids and score bits)",
        "e2e": {"value": n_q / (e2e_s / steps), "unit": UNIT, "h2d_bytes_per_step": int(qn.nbytes),
                "d2h_bytes_per_step": int(hid.numel() * 4 + hsc.numel() * 4),
                "api": "DocShardedSearcher.search (bm25_search per shard + all_gather + bm25_merge_topk)"},
        "roofline": roofline_record(ctx, posting_bytes, km, posting_bytes + 8 * k * n_q, ms_per_step),
        "config": workload_config(ctx.args, workload, n_docs, synths[0].n_terms, synths[0].nnz, n_q, int(q.shape[1]), k,
                                  synth.WORKLOADS[workload].get("r0", 8), "doc-shard", world, total_shards),
        "step_ms_min": float(min(step_ms)), "wall_s_timed_region": wall,
    }
    if cpu:
        rec["cpu_baseline"] = cpu
    for ix in indexes:
        ix.close()
    del indexes, synths, searcher
    torch.cuda.empty_cache()
    return rec


# ----------------------------------------------------------------------------------------------
def run_ours(args):
    import torch

    from mojo_bm25_b200 import engine, synth

    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl ours needs a CUDA device: the BM25 path has no CPU fallback")
    ctx = Ctx(args)
    rank, world = ctx.rank, ctx.world
    mode = headline_mode(args, world)
    sampler = ClockSampler(ctx.local)
    sampler.start()
    launches0 = engine.kernel_launches()
    wl_cfg = synth.WORKLOADS[args.workload]
    subs = {}

    if mode == "doc-shard":
        rec = run_doc_shard(ctx, args.workload, args.steps, args.warmup, scale=args.scale,
                            want_cpu=not args.no_cpu_baseline and world == 1)
        launches = engine.kernel_launches() - launches0
        line = {"metric": METRIC, "value": rec["value"], "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": rec["ms_per_step"], "higher_is_better": True,
                "scaling": rec["scaling"], "vs_baseline": None, "dtype": DTYPE, "data": "synthetic",
                "config": rec["config"], "e2e": rec["e2e"], "gpu_launches": int(launches), "roofline": rec["roofline"],
                "cpu_baseline": rec.get("cpu_baseline"), "doc_shard": rec}
    else:
        # headline: the same index (seed 0) on every rank, rank-specific queries (seed 1 + rank)
        idx, q, k = synth.make_workload(args.workload, device=str(ctx.dev), index_seed=0, query_seed=1 + rank,
                                        scale=args.scale)
        if args.k:
            k = args.k
        index = engine.DeviceIndex.from_torch(idx.indptr, idx.indices, idx.data, idx.n_docs)
        rec = measure_single(ctx, index, q, k, args.steps, args.warmup,
                             traffic_key=args.workload if (args.scale == 1.0 and not args.k) else None)
        launches = engine.kernel_launches() - launches0  # headline only: timed loop + kernel-split pass + e2e
        n_q = int(q.shape[0])
        shape = (idx.n_docs, idx.n_terms, idx.nnz)
        cpu = None
        if not args.no_cpu_baseline and world == 1:
            from oracle import cpu_baseline

            cores = cpu_baseline.host_cores()
            n_sample = cpu_sample_size(n_q, idx.n_docs, cores, args.cpu_sample)
            indptr, indices, data = idx.numpy()
            res = run_cpu_port(indptr, indices, data, idx.n_docs, q.cpu().numpy()[:n_sample], k, 1, 0)
            cpu = {"value": res["qps"], "unit": UNIT, "cores": res["cores"], "kind": "port",
                   "sample": f"first {n_sample} of {n_q} queries of the same batch, oracle port of BM25v.search, "
                             f"{res['cores']} forked workers, {res['step_seconds'][0]:.2f} s"}
        if not args.no_subrecords and args.scale == 1.0:
            sub_steps = max(5, min(args.steps, 10))
            if args.workload == "B":
                subs["k100"] = sub_workload(ctx, "B", k=100, steps=sub_steps, index=index, idx=idx)
                subs["wE"] = sub_workload(ctx, "E", steps=sub_steps, index=index, idx=idx)  # E queries, same 1M-doc index
        index.close()
        del index, idx
        torch.cuda.empty_cache()
        if not args.no_subrecords and args.scale == 1.0:
            sub_steps = max(5, min(args.steps, 10))
            if args.workload != "10M":
                subs["w10M"] = sub_workload(ctx, "10M", steps=sub_steps)
                torch.cuda.empty_cache()
                subs["w10M_bf16"] = sub_workload(ctx, "10M", steps=sub_steps, compress=True)
            torch.cuda.empty_cache()
            if world == 1 and args.workload != "C":
                subs["wC"] = sub_workload(ctx, "C", steps=5)
            torch.cuda.empty_cache()
            if 8 % world == 0:
                subs["doc_shard"] = run_doc_shard(ctx, "D", max(3, min(args.steps, 5)), 3, scale=args.doc_shard_scale)
        line = {"metric": METRIC, "value": rec["value"], "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": rec["ms_per_step"], "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": DTYPE, "data": "synthetic",
                "config": workload_config(args, args.workload, shape[0], shape[1], shape[2], n_q, int(q.shape[1]), k,
                                          wl_cfg.get("r0", 8), mode, world, 1),
                "e2e": rec["e2e"], "gpu_launches": int(launches), "roofline": rec["roofline"], "cpu_baseline": cpu,
                "wall_s_timed_region": rec["wall_s_timed_region"], "step_ms_min": rec["step_ms_min"],
                "step_ms_median": rec["step_ms_median"]}
        line.update(subs)
    line["clocks"] = sampler.stop()
    if rank == 0:
        print(json.dumps(line), flush=True)
    if ctx.dist is not None:
        ctx.dist.barrier()
        ctx.dist.destroy_process_group()


def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
