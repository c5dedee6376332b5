"""Aggregate an ncu source page by line ranges: python scripts/ncu_regions.py rep 'name:lo-hi,name:lo-hi'"""
import csv, subprocess, sys
rep = sys.argv[1]
regions = [(r.split(":")[0], int(r.split(":")[1].split("-")[0]), int(r.split(":")[1].split("-")[1])) for r in sys.argv[2].split(",")]
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Line No")
hdr = rows[hi]; col = {}
for i, h in enumerate(hdr): col.setdefault(h, i)
samp, inst = col["# Samples"], col["Instructions Executed"]
agg = {n: [0.0, 0.0] for n, _, _ in regions}; agg["other"] = [0.0, 0.0]
seen_first = False
for r in rows[hi + 1:]:
    if len(r) <= inst or r[0] == "": continue
    try: ln = int(r[0]); s = float(r[samp] or 0); n = float(r[inst] or 0)
    except ValueError: continue
    for name, lo, hi_ in regions:
        if lo <= ln <= hi_: agg[name][0] += s; agg[name][1] += n; break
    else: agg["other"][0] += s; agg["other"][1] += n
ts = sum(v[0] for v in agg.values()) or 1; ti = sum(v[1] for v in agg.values()) or 1
for k, v in agg.items(): print(f"{k:>14}: {100*v[0]/ts:5.1f}% samples  {100*v[1]/ti:5.1f}% instr  ({v[1]:.3g} warp instr)")
