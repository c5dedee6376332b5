"""Print the headline and the sub-records of a bench.py JSON line: python scripts/show_bench.py file.json"""
import json
import sys

d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(f"N={d['n_gpus']} value={d['value']:.0f} q/s ms/step={d['ms_per_step']:.4f} e2e={d['e2e']['value']:.0f} "
      f"frac={d['roofline']['frac']:.3f} launches={d['gpu_launches']} | {d['config']['parallelism'][:60]}")
print("clocks", d.get("clocks"), "cpu", (d.get("cpu_baseline") or {}).get("value"))
for s in ["k100", "wE", "w10M", "wC", "doc_shard"]:
    x = d.get(s)
    if x:
        extra = {k: round(v, 3) for k, v in x.items() if k.endswith("_ms")}
        print(f"  {s:9s} value={x['value']:.0f} ms/step={x['ms_per_step']:.3f} frac={x['roofline']['frac']:.3f} "
              f"e2e={x['e2e']['value']:.0f} {extra} parity_checked={x.get('parity_checked')}")
