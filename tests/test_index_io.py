"""On-disk bm25s index format: loader and byte-exact writer.  CPU only."""
import filecmp
import json
import os

import numpy as np
import pytest

from mojo_bm25_b200 import index_io


def test_load_bundled_index(golden_dir):
    d = index_io.load_index(os.path.join(golden_dir, "animal_index_bm25"), load_corpus=True)
    assert d.num_docs == 4 and d.num_terms == 20 and d.indices.shape == (20,)
    assert d.indptr.dtype == np.int32 and d.indices.dtype == np.int32 and d.data.dtype == np.float32
    assert d.vocab["fish"] == 17 and d.vocab[""] == 20  # vocab id 20 has no CSC column
    assert d.params["method"] == "lucene" and d.params["k1"] == 1.5 and d.params["b"] == 0.75
    assert d.corpus_offsets == [0, 57, 129, 192] and d.corpus[3]["text"].startswith("a fish")
    g = json.load(open(os.path.join(golden_dir, "golden_bundled.json")))
    assert d.indptr.tolist() == g["indptr"] and d.indices.tolist() == g["indices"]
    assert d.data.view(np.uint32).tolist() == g["data_bits"]


def test_documents_are_fetched_through_the_offset_index(golden_dir, tmp_path):
    """corpus.mmindex.json is used to seek: a document is readable even when every OTHER line of
    corpus.jsonl is unparsable, and only the requested bytes are touched."""
    import shutil

    src = os.path.join(golden_dir, "animal_index_bm25")
    dst = str(tmp_path / "idx")
    shutil.copytree(src, dst)
    raw = open(os.path.join(dst, "corpus.jsonl"), "rb").read()
    offs = json.load(open(os.path.join(dst, "corpus.mmindex.json")))
    assert offs == [0, 57, 129, 192]
    broken = bytearray(raw)
    for i in range(offs[1], offs[3] - 1):  # wreck documents 1 and 2, keep their length (and newlines)
        if broken[i] != 0x0A:
            broken[i] = ord("#")
    open(os.path.join(dst, "corpus.jsonl"), "wb").write(bytes(broken))
    d = index_io.load_index(dst, load_corpus=True)
    assert isinstance(d.corpus, index_io.JsonlCorpus) and len(d.corpus) == 4
    assert d.corpus[3]["text"].startswith("a fish") and d.corpus[0]["id"] == 0 and d.corpus[-1]["id"] == 3
    with pytest.raises(ValueError):
        d.corpus[1]
    whole = index_io.load_index(src, load_corpus=True)
    assert [doc["id"] for doc in whole.corpus] == [0, 1, 2, 3] and whole.corpus[1:3][1]["id"] == 2


def test_writer_is_byte_exact(golden_dir, tmp_path):
    src = os.path.join(golden_dir, "animal_index_bm25")
    d = index_io.load_index(src, load_corpus=True)
    index_io.save_index(str(tmp_path), d.indptr, d.indices, d.data, d.vocab, d.num_docs, corpus=d.corpus)
    for f in sorted(os.listdir(src)):
        assert filecmp.cmp(os.path.join(src, f), os.path.join(tmp_path, f), shallow=False), f


@pytest.mark.parametrize("defect", ["indptr_end", "indptr_monotone", "doc_range", "length"])
def test_malformed_index_is_rejected(golden_dir, tmp_path, defect):
    d = index_io.load_index(os.path.join(golden_dir, "animal_index_bm25"))
    indptr, indices, data = d.indptr.copy(), d.indices.copy(), d.data.copy()
    if defect == "indptr_end":
        indptr[-1] = 19
    elif defect == "indptr_monotone":
        indptr[3], indptr[4] = indptr[4], indptr[3] - 1
    elif defect == "doc_range":
        indices[5] = 4
    else:
        data = data[:-1]
    with pytest.raises(ValueError):
        index_io.validate(indptr, indices, data, 4)
