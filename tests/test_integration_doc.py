"""INTEGRATION.md promises a ~40-line ctypes stub a maintainer can drop into the reference.  Run that
exact code block (with the library path made absolute) so the document cannot rot."""
import os

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _stub_source():
    text = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    code = text[text.index("```python") + len("```python"):]
    code = code[: code.index("```")]
    so = os.path.join(ROOT, "mojo_bm25_b200", "libbm25_b200.so")
    assert 'ctypes.CDLL("libbm25_b200.so")' in code
    return code.replace('ctypes.CDLL("libbm25_b200.so")', f'ctypes.CDLL("{so}")')


def test_stub_binds_every_symbol_it_uses():
    from mojo_bm25_b200 import build

    build.build()
    ns = {}
    exec(compile(_stub_source(), "INTEGRATION.md", "exec"), ns)  # loads the library, declares argtypes
    assert "DeviceIndex" in ns and callable(ns["_check"])


@pytest.mark.gpu
def test_stub_searches_like_the_shipped_binding():
    ns = {}
    exec(compile(_stub_source(), "INTEGRATION.md", "exec"), ns)
    indptr = np.array([0, 2, 3], np.int32)
    indices = np.array([7, 900, 3], np.int32)
    data = np.array([1.0, 2.0, 5.0], np.float32)
    ix = ns["DeviceIndex"](indptr, indices, data, 1000)
    ids, sc = ix.search(np.array([[0, -1], [1, 0]], np.int32), 3)
    assert ids.tolist() == [[900, 7, 0], [3, 900, 7]] and sc.tolist() == [[2.0, 1.0, 0.0], [5.0, 2.0, 1.0]]
    with pytest.raises(ValueError):  # token id >= vocabulary, as bm25_native.py:116-121
        ix.search(np.array([[5]], np.int32), 3)
