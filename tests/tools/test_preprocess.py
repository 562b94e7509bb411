"""tools/preprocess_data.py end to end (reference tests/tools/test_preprocess.sh runs the same tool on downloaded data):
jsonl → tokenised .bin/.idx in every ``--dataset-impl`` → read back through ``get_indexed_dataset``."""
import json
import subprocess
import sys

import numpy as np
import pytest

from libai_b200.data.data_utils import helpers
from libai_b200.data.data_utils.indexed_dataset import get_indexed_dataset
from libai_b200.tokenizer import BertTokenizer

DOCS = [
    "hello world. good movie! bad movie?",
    "the quick brown fox. jumps over the lazy dog.",
    "good world",
]
VOCAB = ["[PAD]", "[UNK]", "[CLS]", "[SEP]", "[MASK]", "hello", "world", "good", "bad", "movie", "the", "quick", "brown",
         "fox", "jumps", "over", "lazy", "dog", ".", "!", "?"]


@pytest.mark.parametrize("impl", ["mmap", "lazy", "cached"])
@pytest.mark.parametrize("split", [False, True])
def test_preprocess_tool(tmp_path, impl, split):
    vocab = tmp_path / "vocab.txt"
    vocab.write_text("\n".join(VOCAB) + "\n")
    src = tmp_path / "docs.jsonl"
    src.write_text("".join(json.dumps({"text": d}) + "\n" for d in DOCS))
    prefix = tmp_path / "out"
    cmd = [sys.executable, "tools/preprocess_data.py", "--input", str(src), "--tokenizer-name", "BertTokenizer",
           "--vocab-file", str(vocab), "--do-lower-case", "--output-prefix", str(prefix), "--dataset-impl", impl,
           "--workers", "2"] + (["--split-sentences"] if split else [])
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=300)
    assert res.returncode == 0, res.stderr[-2000:]
    level = "sentence" if split else "document"
    ds = get_indexed_dataset(f"{prefix}_text_{level}", impl, skip_warmup=True)
    tok = BertTokenizer(str(vocab), do_lower_case=True, do_chinese_wwm=False)
    want_docs = [tok.convert_tokens_to_ids(tok.tokenize(d)) for d in DOCS]
    got = [np.asarray(ds[i]).tolist() for i in range(len(ds))]
    assert sum(got, []) == sum(want_docs, [])                       # same token stream
    if split:
        assert len(ds) > len(DOCS) and len(ds.doc_idx) == len(DOCS) + 1   # several sentences per document
        assert ds.doc_idx[-1] == len(ds)
    else:
        assert got == want_docs


def test_build_blocks_mapping():
    """REALM/ICT block mapping (exported for API parity, reference helpers.cpp:393-600): blocks tile each document's
    sentences in order, respect the length budget (minus the title), and carry increasing block ids before shuffling."""
    rng = np.random.default_rng(0)
    n_docs = 12
    sent_per_doc = rng.integers(2, 7, n_docs)
    docs = np.concatenate([[0], np.cumsum(sent_per_doc)]).astype(np.int64)
    sizes = rng.integers(5, 30, int(docs[-1])).astype(np.int32)
    titles = rng.integers(2, 6, n_docs).astype(np.int32)
    max_len = 64
    m = helpers.build_blocks_mapping(docs, sizes, titles, num_epochs=1, max_num_samples=10 ** 6, max_seq_length=max_len,
                                     seed=1, verbose=False, use_one_sent_blocks=False)
    assert m.ndim == 2 and m.shape[1] == 4 and len(m) > 0
    assert sorted(m[:, 3].tolist()) == list(range(len(m)))          # block ids are a permutation
    for start, end, doc, _ in m.tolist():
        assert docs[doc] <= start < end <= docs[doc + 1]
        assert end - start >= 2                                     # one-sentence blocks excluded
    again = helpers.build_blocks_mapping(docs, sizes, titles, 1, 10 ** 6, max_len, 1)
    assert np.array_equal(m, again)                                 # deterministic for a seed
    one = helpers.build_blocks_mapping(docs, sizes, titles, 1, 10 ** 6, max_len, 1, use_one_sent_blocks=True)
    assert len(one) >= len(m)
