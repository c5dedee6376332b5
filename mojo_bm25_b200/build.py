"""In-tree build of libbm25_b200.so (sm_100a only).  nvcc cross-compiles without a GPU."""
import os
import shutil
import subprocess

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
CSRC = os.path.join(PKG, "csrc")
SO = os.path.join(PKG, "libbm25_b200.so")
SOURCES = [os.path.join(CSRC, "bm25_capi.cu")]
DEPS = SOURCES + [os.path.join(CSRC, "bm25_kernels.cuh"), os.path.join(ROOT, "include", "bm25_b200.h")]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-shared", "-Xcompiler", "-fPIC", "-diag-suppress", "128",
]


def nvcc_path() -> str:
    cand = os.environ.get("NVCC") or shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(cand):
        raise RuntimeError("nvcc not found; libbm25_b200.so cannot be built (there is no CPU fallback)")
    return cand


def is_stale() -> bool:
    if not os.path.exists(SO):
        return True
    t = os.path.getmtime(SO)
    return any(os.path.getmtime(d) > t for d in DEPS)


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile the CUDA extension for sm_100a if missing or stale; returns the .so path."""
    if not force and not is_stale():
        return SO
    cmd = [nvcc_path()] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", SO] + SOURCES
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + " ".join(cmd) + "\n" + res.stdout + res.stderr)
    if verbose:
        print(res.stderr)
    return SO


if __name__ == "__main__":
    print(build(force=True, verbose=True))
