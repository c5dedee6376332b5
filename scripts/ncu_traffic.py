"""DRAM traffic per launch of the captured score kernels -> profiles/<tag>_traffic.json, stamped with
the sha256 of the kernel sources the captures were made from (bench.py reports `roofline.traffic`
only while that stamp matches the sources it runs)."""
import csv
import hashlib
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def source_stamp():
    h = hashlib.sha256()
    for f in ("mojo_bm25_b200/csrc/bm25_kernels.cuh", "mojo_bm25_b200/csrc/bm25_capi.cu"):
        h.update(open(os.path.join(ROOT, f), "rb").read())
    return h.hexdigest()


def main():
    tag = sys.argv[1] if len(sys.argv) > 1 else "r2"
    out = {"kernel_source_sha256": source_stamp(), "unit": "bytes per launch (dram__bytes_read.sum + dram__bytes_write.sum, ncu --set full)"}
    for wl in ("B", "10M", "E", "C", "10M_bf16"):
        rep = os.path.join(ROOT, "gpurun_out", f"{tag}_score_{wl}.ncu-rep")
        if not os.path.exists(rep):
            continue
        raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(raw.splitlines()))
        hdr, units, r = rows[0], rows[1], rows[2]

        def val(name):
            i = hdr.index(name)
            scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[units[i]]
            return float(r[i]) * scale

        out[wl] = {"dram_read": val("dram__bytes_read.sum"), "dram_write": val("dram__bytes_write.sum"),
                   "kernel": r[hdr.index("Kernel Name")]}
    json.dump(out, open(os.path.join(ROOT, "profiles", f"{tag}_traffic.json"), "w"), indent=1)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
