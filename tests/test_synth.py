"""Synthetic corpora: canonical CSC form, determinism, and the difference
between the stratified-uniform and the clustered (bursty) doc-id generators.  CPU only."""
import numpy as np
import pytest

from mojo_bm25_b200 import synth


def _columns(idx):
    ip, ind, dat = idx.numpy()
    return ip, ind, dat


@pytest.mark.parametrize("name", ["tiny", "tinyc"])
def test_canonical_form_and_determinism(name):
    idx, q, k = synth.make_workload(name)
    ip, ind, dat = _columns(idx)
    assert ip[0] == 0 and ip[-1] == len(ind) == idx.nnz and np.all(np.diff(ip) >= 0)
    assert ind.min() >= 0 and ind.max() < idx.n_docs and np.all(dat > 0) and np.all(np.isfinite(dat))
    starts = np.zeros(len(ind), bool)
    starts[ip[:-1][np.diff(ip) > 0]] = True
    assert np.all((np.diff(ind) > 0) | starts[1:])  # strictly increasing doc ids inside every column
    idx2, q2, _ = synth.make_workload(name)
    assert np.array_equal(ind, idx2.indices.numpy()) and np.array_equal(dat, idx2.data.numpy())
    assert np.array_equal(q.numpy(), q2.numpy()) and q.dtype.is_floating_point is False


def test_clustered_generator_is_bursty_and_keeps_the_document_frequencies():
    a, _, _ = synth.make_workload("tiny")
    b, _, _ = synth.make_workload("tinyc")
    assert np.array_equal(a.indptr.numpy(), b.indptr.numpy())  # same df per term, only the placement differs
    ipa, inda, _ = _columns(a)
    ipb, indb, _ = _columns(b)
    t = 20
    ha = np.bincount(inda[ipa[t]:ipa[t + 1]] // 2048, minlength=10)[:9]
    hb = np.bincount(indb[ipb[t]:ipb[t + 1]] // 2048, minlength=10)[:9]
    assert ha.max() - ha.min() <= 2          # stratified: every tile gets the same share
    assert hb.max() > 1.25 * hb.min()        # clustered: tiles differ
