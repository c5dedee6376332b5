"""Per-source-line instruction counts of an ncu report (sorted by executed warp instructions)."""
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
data = []
for r in rows[hdr_i + 1:]:
    if len(r) <= inst or r[0] == "":
        continue
    try:
        data.append((float(r[inst] or 0), float(r[samp] or 0), r))
    except ValueError:
        pass
tot_i = sum(d[0] for d in data) or 1
print(f"total warp instructions {tot_i:.0f}")
for i, s, r in sorted(data, key=lambda d: -d[0])[:top]:
    print(f"{100*i/tot_i:5.1f}% ins {i:14.0f}  L{r[0]:>4} {r[1].strip()[:100]}")
