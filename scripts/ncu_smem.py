"""Shared-memory wavefronts per SASS instruction of an ncu report (top contributors), with the
CUDA source line each instruction belongs to."""
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
wf, ideal, inst = col["L1 Wavefronts Shared"], col["L1 Wavefronts Shared Ideal"], col["Instructions Executed"]
cur = ""
data = []
for r in rows[hdr_i + 1:]:
    if len(r) <= wf:
        continue
    if r[0] != "":
        cur = f"L{r[0]} {r[1].strip()[:70]}"
        continue
    try:
        w = float(r[wf] or 0)
    except ValueError:
        continue
    if w > 0:
        data.append((w, float(r[ideal] or 0), float(r[inst] or 0), r[1].strip()[:40], cur))
tot = sum(d[0] for d in data)
print(f"total shared wavefronts {tot:.0f}")
for w, i, n, sass, src in sorted(data, key=lambda d: -d[0])[:top]:
    print(f"{100*w/tot:5.1f}%  wf={w/1e6:7.1f}M ideal={i/1e6:7.1f}M inst={n/1e6:6.1f}M  {sass:<40} {src}")
