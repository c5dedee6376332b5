"""CPU-side checks of the C-ABI library: it builds, loads, exports every symbol that
include/bm25_b200.h declares, and fails loudly (no CPU fallback) when there is no GPU."""
import ctypes
import os
import re

import numpy as np
import pytest

from mojo_bm25_b200 import _lib, build

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "bm25_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(bm25_[a-z_0-9]+)\s*\(", src)))


def test_library_builds_and_exports_all_declared_symbols():
    so = build.build()
    assert os.path.exists(so)
    lib = ctypes.CDLL(so)
    declared = _declared_symbols()
    assert len(declared) >= 14
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/bm25_b200.h but not exported"
    assert sorted(_lib.SYMBOLS) == declared


def test_version_and_launch_counter():
    lib = _lib.load()
    assert b"sm_100a" in lib.bm25_version()
    assert lib.bm25_kernel_launches() >= 0


def test_sass_contains_only_sm100a():
    import shutil
    import subprocess

    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        pytest.skip("cuobjdump not available")
    out = subprocess.run([cuobjdump, "-lelf", build.build()], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_\d+a?", out))
    assert archs == {"sm_100a"}, archs


def test_no_cpu_fallback_without_gpu():
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from mojo_bm25_b200.engine import DeviceIndex

    indptr = np.array([0, 1, 2], np.int32)
    with pytest.raises(Exception) as ei:
        DeviceIndex(indptr, np.array([0, 1], np.int32), np.array([1.0, 2.0], np.float32), n_docs=2)
    assert "no CPU fallback" in str(ei.value) or "CUDA" in str(ei.value)


def test_argument_validation_happens_before_device_use():
    lib = _lib.load()
    h = ctypes.c_void_p()
    rc = lib.bm25_index_create(None, None, None, 0, 0, 0, 0, 0, ctypes.byref(h))
    assert rc == _lib.ERR_INVALID
    assert b"NULL" in lib.bm25_last_error()
    rc = lib.bm25_search(None, None, 1, 1, 1, None, None, None)
    assert rc == _lib.ERR_INVALID


def test_int64_index_arrays_are_range_checked_before_the_int32_cast():
    """int64 indptr / indices must not wrap silently (no GPU needed: the check runs
    before the library is called)."""
    from mojo_bm25_b200 import engine

    big = np.array([0, 2 ** 31 + 5], np.int64)
    with pytest.raises(ValueError, match="int32"):
        engine.DeviceIndex(big, np.zeros(1, np.int64), np.ones(1, np.float32), n_docs=10)
    with pytest.raises(ValueError, match="int32"):
        engine.DeviceIndex(np.array([0, 1], np.int64), np.array([2 ** 40], np.int64), np.ones(1, np.float32), n_docs=10)
    with pytest.raises(ValueError, match="integer"):
        engine.DeviceIndex(np.array([0.0, 1.0]), np.array([0], np.int64), np.ones(1, np.float32), n_docs=10)


def test_round_to_bf16_is_round_to_nearest_even():
    """Host-side twin of k_quantize (compressed handles): round to nearest even on the upper 16 bits,
    the same values torch's float32 -> bfloat16 conversion produces."""
    import numpy as np
    import torch

    from mojo_bm25_b200 import engine

    x = np.array([1.0, 1.00390625, 1.005859375, 1.01171875, 3.1415927, 1e-30, 6.5e4], np.float32)
    r = engine.round_to_bf16(x)
    assert np.all((r.view(np.uint32) & 0xFFFF) == 0)
    # ties go to the even mantissa: 1 + 2^-8 -> 1.0, 1 + 3*2^-8 -> 1 + 2^-6; 1 + 3*2^-9 rounds up to 1 + 2^-7
    assert r[1] == np.float32(1.0) and r[2] == np.float32(1.0078125) and r[3] == np.float32(1.015625)
    assert np.all(np.abs(r - x) <= np.abs(x) * 2.0 ** -8)
    rng = np.random.default_rng(0)
    y = (rng.standard_normal(100000) * 10.0 ** rng.uniform(-20, 20, 100000)).astype(np.float32)
    t = torch.from_numpy(y).to(torch.bfloat16).to(torch.float32).numpy()
    assert np.array_equal(t.view(np.uint32), engine.round_to_bf16(y).view(np.uint32))
