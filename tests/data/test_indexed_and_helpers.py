"""Indexed datasets (3 on-disk formats) and the C++ index helpers vs Python oracles."""
import os
import tempfile

import numpy as np
import pytest
import torch

from libai_b200.data.data_utils import get_indexed_dataset, helpers, indexed_dataset


@pytest.mark.parametrize("impl", ["mmap", "lazy", "cached"])
def test_indexed_dataset_roundtrip(impl):
    rng = np.random.RandomState(0)
    docs = [[rng.randint(0, 1000, size=rng.randint(1, 9)) for _ in range(rng.randint(1, 4))] for _ in range(12)]
    with tempfile.TemporaryDirectory() as d:
        prefix = os.path.join(d, "corpus")
        builder = indexed_dataset.make_builder(prefix + ".bin", impl=impl, vocab_size=1000)
        for doc in docs:
            for sent in doc:
                builder.add_item(torch.IntTensor(sent))
            builder.end_document()
        builder.finalize(prefix + ".idx")
        assert indexed_dataset.infer_dataset_impl(prefix) == ("mmap" if impl == "mmap" else "cached")
        ds = get_indexed_dataset(prefix, impl, True)
        flat = [s for doc in docs for s in doc]
        if impl == "cached":
            assert ds.supports_prefetch
            ds.prefetch(list(range(len(flat))))
        assert len(ds) == len(flat)
        for i, s in enumerate(flat):
            assert np.array_equal(np.asarray(ds[i]), s)
        assert list(ds.doc_idx) == list(np.cumsum([0] + [len(doc) for doc in docs]))
        assert list(ds.sizes) == [len(s) for s in flat]


def _sample_idx_oracle(sizes, doc_idx, seq_length, num_epochs, tokens_per_epoch):
    num_samples = (num_epochs * tokens_per_epoch - 1) // seq_length
    out = np.zeros((num_samples + 1, 2), dtype=np.int32)
    di, off = 0, 0
    for s in range(1, num_samples + 1):
        remaining = seq_length + 1
        while remaining != 0:
            length = sizes[doc_idx[di]] - off
            remaining -= length
            if remaining <= 0:
                off += remaining + length - 1
                remaining = 0
            else:
                di += 1
                off = 0
        out[s] = (di, off)
    return out


def test_build_sample_idx_matches_oracle():
    rng = np.random.RandomState(1)
    sizes = rng.randint(5, 60, size=40).astype(np.int32)
    epochs = 3
    doc_idx = np.concatenate([rng.permutation(40) for _ in range(epochs)]).astype(np.int32)
    tokens_per_epoch = int(sizes.sum())
    got = helpers.build_sample_idx(sizes, doc_idx, 32, epochs, tokens_per_epoch)
    assert np.array_equal(got, _sample_idx_oracle(sizes, doc_idx, 32, epochs, tokens_per_epoch))


def test_build_mapping_properties():
    rng = np.random.RandomState(2)
    n_docs = 30
    sent_per_doc = rng.randint(2, 8, size=n_docs)
    docs = np.concatenate([[0], np.cumsum(sent_per_doc)]).astype(np.int64)
    sizes = rng.randint(4, 40, size=int(docs[-1])).astype(np.int32)
    a = helpers.build_mapping(docs, sizes, 2, 10_000, 64, 0.1, 1234, False, 2)
    b = helpers.build_mapping(docs, sizes, 2, 10_000, 64, 0.1, 1234, False, 2)
    assert np.array_equal(a, b) and a.shape[1] == 3 and len(a) > 0      # deterministic in the seed
    c = helpers.build_mapping(docs, sizes, 2, 10_000, 64, 0.1, 4321, False, 2)
    assert not np.array_equal(a, c)
    for start, end, target in a:
        d = np.searchsorted(docs, start, side="right") - 1
        assert docs[d] <= start < end <= docs[d + 1]                     # a sample never crosses a document
        assert end - start >= 2 and 2 <= target <= 64


def test_build_blending_indices():
    weights = np.array([0.7, 0.3])
    di = np.zeros(1000, dtype=np.uint8)
    dsi = np.zeros(1000, dtype=np.int64)
    helpers.build_blending_indices(di, dsi, weights, 2, 1000, False)
    frac = (di == 0).mean()
    assert abs(frac - 0.7) < 0.01
    for k in (0, 1):
        assert np.array_equal(dsi[di == k], np.arange((di == k).sum()))
