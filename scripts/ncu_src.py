"""Summarise `ncu --page source --csv --print-source cuda,sass` by CUDA source line."""
import csv
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == "Line No")
hdr = rows[hdr_i]
col = {}
for i, h in enumerate(hdr):
    col.setdefault(h, i)
samp, inst = col["# Samples"], col["Instructions Executed"]
stall_cols = [h for h in hdr if h.startswith("stall_")]
data = []
for r in rows[hdr_i + 1:]:
    if len(r) <= inst or r[0] == "":
        continue  # SASS rows have an empty line number
    try:
        data.append((float(r[samp] or 0), float(r[inst] or 0), r))
    except ValueError:
        pass
tot_s = sum(d[0] for d in data) or 1
tot_i = sum(d[1] for d in data) or 1
print(f"total samples {tot_s:.0f}, total warp instructions {tot_i:.0f}")
for s, i, r in sorted(data, key=lambda d: -d[0])[:top]:
    stalls = sorted(((float(r[col[h]] or 0), h) for h in stall_cols), reverse=True)[:3]
    st = " ".join(f"{h[6:]}={v:.0f}" for v, h in stalls if v > 0)
    print(f"{100*s/tot_s:5.1f}% smp {100*i/tot_i:5.1f}% ins  L{r[0]:>4} {r[1].strip()[:88]:<88} {st}")
