#!/usr/bin/env python
"""bench.py -- BM25 query hot-path benchmark (contract: see DESIGN.md "Measurement").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload B|C|10M|D|E|tiny]
                    [--mode auto|query-split|doc-shard] [--impl ours|reference]

A *step* is one pass of the hot path (segment table -> score accumulation + per-range top-k ->
merge) over one batch of synthetic queries.  Default workload = BASELINE.json configs[1] ("B":
1M docs, 100k-term Zipf vocabulary, 1000 queries x 4 terms, top-10).  One JSON line is printed by
rank 0.  `value` = whole-job queries/s with the index and the queries resident in HBM (CUDA-event
timed, L2 flushed between steps); `e2e` = the same metric through the host-buffer C-ABI call
(bm25_search_host: pinned H2D of the queries, kernels, D2H of ids+scores, every step).
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
    ap.add_argument("--cpu-sample", type=int, default=0, help="queries in the CPU baseline sample")
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
    """Samples SM clock + throttle reasons of one GPU during the timed region (NVML thread)."""

    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap",
               0x80: "hw_power_brake_slowdown", 0x2: "applications_clocks_setting", 0x100: "display_clock_setting"}

    def __init__(self, torch_device_index: int, period_s: float = 0.02):
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


def workload_config(args, wl_cfg, idx, queries, k, extra=None):
    cfg = {
        "workload": f"{args.workload}: synthetic saturating-Zipf CSC index, {idx.n_docs} docs, {idx.n_terms} terms, "
                    f"nnz={idx.nnz}; {queries.shape[0]} queries x {queries.shape[1]} term slots (r0={wl_cfg.get('r0', 8)}), "
                    f"top-{k}",
        "n_docs": idx.n_docs, "n_terms": idx.n_terms, "nnz": idx.nnz, "n_queries": int(queries.shape[0]),
        "query_width": int(queries.shape[1]), "k": int(k), "index_seed": 0, "query_seed": 1,
        "cache": "L2 flushed (256 MiB memset) between timed steps; index is larger than L2",
    }
    if args.scale != 1.0:
        cfg["scale"] = args.scale
    if extra:
        cfg.update(extra)
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
    line = {
        "impl": "reference", "metric": METRIC, "value": qps, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * float(np.mean(res["step_seconds"])), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": DTYPE, "data": "synthetic",
        "config": workload_config(args, wl_cfg, idx, q, k),
        "cpu_baseline": {"value": qps, "unit": UNIT, "cores": res["cores"], "kind": "port", "sample": sample},
        "e2e": {"value": qps, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------------
def run_ours(args):
    import numpy as np
    import torch

    from mojo_bm25_b200 import engine, sharded, synth

    rank, world, local = dist_env()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl ours needs a CUDA device: the BM25 path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist = None
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=dev)
    mode = args.mode
    if mode == "auto":
        mode = "query-split" if world > 1 and args.workload != "D" else ("doc-shard" if world > 1 else "single")
    if args.workload == "D" and mode != "query-split":
        mode = "doc-shard"  # the 100M-doc corpus only exists as 8 document shards, also on one GPU
    elif world == 1:
        mode = "single"

    wl_cfg = synth.WORKLOADS[args.workload]
    # doc-shard: the corpus is a set of document-range shards, each synthesised with seed = shard
    #   number and searched through its own int32-indexed handle.  Workload D is a FIXED corpus of 8
    #   shards x 12.5M docs = 100M docs (3e9 postings): rank r owns shards r, r+W, ... (strong scaling;
    #   one GPU holds all 8).  Other workloads under --mode doc-shard get one shard per rank (weak).
    # query-split / single: the same index (seed 0) on every rank, rank-specific queries (seed 1+rank).
    if mode == "doc-shard":
        total_shards = 8 if args.workload == "D" else world
        if total_shards % world:
            raise SystemExit(f"workload D has 8 document shards; --gpus must divide 8 (got {world})")
        my_shards = list(range(rank, total_shards, world))
    else:
        total_shards, my_shards = 1, [0]
    query_seed = 1 if mode == "doc-shard" else 1 + rank
    indexes, synths, q, k = [], [], None, None
    for sh in my_shards:
        idx, q, k = synth.make_workload(args.workload, device=str(dev), index_seed=sh, query_seed=query_seed,
                                        scale=args.scale)
        base = sh * idx.n_docs if mode == "doc-shard" else 0
        ix = engine.DeviceIndex.from_torch(idx.indptr, idx.indices, idx.data, idx.n_docs, doc_id_base=base)
        ix.set_option("timing", 1)
        indexes.append(ix)
        synths.append(idx)
    if args.k:
        k = args.k
    idx, index = synths[0], indexes[0]
    qn = q.cpu().numpy()
    n_q = q.shape[0]
    posting_bytes = sum(ix.posting_bytes(qn, 0) for ix in indexes)  # 8 * sum df (score kernel)
    batch_bytes = posting_bytes + 8 * k * n_q
    out_ids = torch.empty((n_q, k), dtype=torch.int32, device=dev)
    out_sc = torch.empty((n_q, k), dtype=torch.float32, device=dev)
    searcher = sharded.DocShardedSearcher.from_index(indexes, k) if mode == "doc-shard" else None

    def step():
        if searcher is not None:
            return searcher.search(q)
        return index.search_device(q, k, out_ids=out_ids, out_scores=out_sc)

    def kernel_ms():  # (segments, score, merge) device ms of this rank's last step, summed over its shards
        return tuple(map(sum, zip(*[ix.last_timing_ms() for ix in indexes])))

    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    for _ in range(args.warmup):
        flush.zero_()
        step()
    torch.cuda.synchronize()

    sampler = ClockSampler(local)
    sampler.start()
    launches0 = engine.kernel_launches()
    starts = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    ends = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    kern_ms = []
    if dist is not None:
        dist.barrier()
    torch.cuda.synchronize()
    wall0 = time.perf_counter()
    for i in range(args.steps):
        flush.zero_()
        starts[i].record()
        step()
        ends[i].record()
        kern_ms.append(kernel_ms())  # waits for this step's kernels (device events)
    torch.cuda.synchronize()
    if dist is not None:
        dist.barrier()
    wall = time.perf_counter() - wall0
    launches = engine.kernel_launches() - launches0
    step_ms = [s.elapsed_time(e) for s, e in zip(starts, ends)]
    total_ms = float(sum(step_ms))
    if dist is not None:
        t = torch.tensor([total_ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms = float(t.item())
    ms_per_step = total_ms / args.steps
    units = n_q * (world if mode == "query-split" else 1)  # queries the whole job answered per step
    value = units / (ms_per_step / 1e3)

    # ---- end-to-end through the host-buffer C-ABI call (pinned H2D + kernels + D2H each step) --
    e2e = None
    if mode != "doc-shard":
        for _ in range(2):
            index.search(qn, k)
        if dist is not None:
            dist.barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            flush.zero_()
            hid, hsc = index.search(qn, k)
        torch.cuda.synchronize()
        e2e_s = (time.perf_counter() - t0)
        if dist is not None:
            t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            e2e_s = float(t.item())
        e2e = {"value": units / (e2e_s / args.steps), "unit": UNIT, "h2d_bytes_per_step": int(qn.nbytes),
               "d2h_bytes_per_step": int(hid.nbytes + hsc.nbytes), "api": "bm25_search_host via DeviceIndex.search"}
    else:
        # doc-shard e2e: pinned queries -> H2D -> local search(es) -> all-gather -> merge -> D2H of the result
        q_pin = torch.from_numpy(qn).pin_memory()
        q_dev = torch.empty_like(q)
        for w in range(2 + args.steps):
            if w == 2:
                torch.cuda.synchronize()
                if dist is not None:
                    dist.barrier()
                t0 = time.perf_counter()
            flush.zero_()
            q_dev.copy_(q_pin, non_blocking=True)
            gi, gs = searcher.search(q_dev)
            hid, hsc = gi.cpu(), gs.cpu()
        e2e_s = time.perf_counter() - t0
        if dist is not None:
            t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            e2e_s = float(t.item())
        e2e = {"value": units / (e2e_s / args.steps), "unit": UNIT, "h2d_bytes_per_step": int(qn.nbytes),
               "d2h_bytes_per_step": int(hid.numel() * 4 + hsc.numel() * 4),
               "api": "DocShardedSearcher.search (bm25_search per shard + all_gather + bm25_merge_topk)"}
    clocks = sampler.stop()

    if rank != 0:
        if dist is not None:
            dist.barrier()
            dist.destroy_process_group()
        return

    peak, peak_src = load_peaks()
    km = np.array(kern_ms)
    score_ms = float(km[:, 1].mean())
    achieved = posting_bytes / (score_ms * 1e-3) / 1e9
    roofline = {
        "bound": "hbm", "kernel": "k_score_topk (score accumulation + per-range top-k)",
        "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
        "frac_of_nominal_8000": achieved / 8000.0, "peak_source": peak_src,
        "algorithmic_bytes_per_launch": int(posting_bytes), "kernel_ms": score_ms,
        "kernel_share_of_step": score_ms / float(km.sum(axis=1).mean()),
        "other_kernels_ms": {"k_segments": float(km[:, 0].mean()), "k_merge": float(km[:, 2].mean())},
        "whole_batch_GBps": batch_bytes / (ms_per_step * 1e-3) / 1e9,
        "traffic": None,
    }
    prof = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(prof):
        try:
            tr = json.load(open(prof)).get(args.workload)
            if tr:
                roofline["traffic"] = tr["dram_bytes_per_launch"]
                roofline["traffic_source"] = tr.get("source")
        except Exception:  # noqa: BLE001
            pass

    cpu = None
    if not args.no_cpu_baseline and world == 1:
        from oracle import cpu_baseline

        cores = cpu_baseline.host_cores()
        n_sample = cpu_sample_size(n_q, idx.n_docs, cores, args.cpu_sample)
        indptr, indices, data = idx.numpy()
        res = run_cpu_port(indptr, indices, data, idx.n_docs, qn[:n_sample], k, 1, 0)
        cpu = {"value": res["qps"] / total_shards, "unit": UNIT, "cores": res["cores"], "kind": "port",
               "sample": f"first {n_sample} of {n_q} queries of the same batch, oracle port of BM25v.search, "
                         f"{res['cores']} forked workers, {res['step_seconds'][0]:.2f} s"
                         + (f"; timed on shard 0 of {total_shards} and divided by {total_shards} (a query visits every shard)"
                            if total_shards > 1 else "")}

    par = {"single": "1 GPU", "query-split": f"query-split x{world}: index replicated, each rank answers its own "
           f"{n_q}-query batch, no data-path collective",
           "doc-shard": f"doc-shard x{world}: {total_shards} shards of {idx.n_docs} docs ({idx.n_docs * total_shards} "
           f"total), {len(my_shards)} per rank, same batch on every rank, local top-k + NCCL all-gather of "
           f"{n_q * k * 8 * len(my_shards)} B per rank + merge kernel"}[mode]
    strong = mode == "doc-shard" and args.workload == "D"
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong" if strong else "weak",
        "vs_baseline": None, "dtype": DTYPE,
        "data": "synthetic", "config": workload_config(args, wl_cfg, idx, q, k, {"parallelism": par, "shards": total_shards}),
        "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": cpu, "clocks": clocks,
        "wall_s_timed_region": wall, "step_ms_min": float(min(step_ms)), "step_ms_median": float(np.median(step_ms)),
    }
    print(json.dumps(line), flush=True)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
