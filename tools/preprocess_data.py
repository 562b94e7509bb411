#!/usr/bin/env python
"""jsonl corpus → tokenised Megatron-format indexed dataset (``<prefix>_<key>_{document,sentence}.bin/.idx``).

CLI parity with the reference tools/preprocess_data.py:63-267 (``--input --json-keys --split-sentences
--keep-newlines --tokenizer-name --vocab-file --merges-file --do-lower-case --extra-ids --append-eod
--do-chinese-wwm --output-prefix --dataset-impl --workers --log-interval``).  Documents are tokenised by a process
pool; the parent writes the ids in input order, one ``end_document`` per json line.  Sentence splitting uses nltk's
punkt model when installed and a punctuation regex (Latin + CJK terminators) otherwise.
"""
import argparse
import json
import multiprocessing
import os
import re
import sys
import time

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), os.path.pardir)))

from libai_b200 import tokenizer as tokenizers  # noqa: E402
from libai_b200.data.data_utils import indexed_dataset  # noqa: E402

_SENT_END = re.compile(r"(?<=[.!?。！？；;])[\"'”’)\]]*\s*")


class _RegexSplitter:
    def __init__(self, keep_newlines):
        self.keep_newlines = keep_newlines

    def tokenize(self, text):
        chunks = text.split("\n") if not self.keep_newlines else [text]
        out = []
        for chunk in chunks:
            start = 0
            for m in _SENT_END.finditer(chunk):
                if m.end() > start and chunk[start : m.end()].strip():
                    out.append(chunk[start : m.end()] if self.keep_newlines else chunk[start : m.end()].strip())
                start = m.end()
            if chunk[start:].strip():
                out.append(chunk[start:] if self.keep_newlines else chunk[start:].strip())
        return out


class _Identity:
    def tokenize(self, *text):
        return text


def _make_splitter(args):
    if not args.split_sentences:
        return _Identity()
    try:
        import nltk

        splitter = nltk.load("tokenizers/punkt/english.pickle")
        return splitter
    except Exception:
        return _RegexSplitter(args.keep_newlines)


def build_tokenizer_from_args(args):
    cls = getattr(tokenizers, args.tokenizer_name)
    kwargs = {}
    if args.tokenizer_name == "BertTokenizer":
        kwargs.update(vocab_file=args.vocab_file, do_lower_case=args.do_lower_case, do_chinese_wwm=args.do_chinese_wwm)
        if args.extra_ids > 0:
            kwargs["additional_special_tokens"] = [f"<extra_id_{i}>" for i in range(args.extra_ids)]
    elif args.tokenizer_name in ("GPT2Tokenizer", "RobertaTokenizer"):
        kwargs.update(vocab_file=args.vocab_file, merges_file=args.merges_file)
    elif args.tokenizer_name == "T5Tokenizer":
        kwargs.update(vocab_file=args.vocab_file, extra_ids=args.extra_ids or 100)
    tok = cls(**kwargs)
    if args.append_eod and tok.eod_token is None:
        tok.eod_token = tok.eos_token if tok.eos_token is not None else tok.pad_token
    return tok


class Encoder:
    """Per-process state (tokenizer + sentence splitter) and the per-line work function."""

    tokenizer = None
    splitter = None

    def __init__(self, args):
        self.args = args

    def initializer(self):
        Encoder.tokenizer = build_tokenizer_from_args(self.args)
        Encoder.splitter = _make_splitter(self.args)

    def encode(self, json_line):
        data = json.loads(json_line)
        ids = {}
        for key in self.args.json_keys:
            doc_ids = []
            for sentence in Encoder.splitter.tokenize(data[key]):
                sentence_ids = Encoder.tokenizer.convert_tokens_to_ids(Encoder.tokenizer.tokenize(sentence))
                if len(sentence_ids) > 0:
                    doc_ids.append(sentence_ids)
            if len(doc_ids) > 0 and self.args.append_eod:
                doc_ids[-1].append(Encoder.tokenizer.eod_token_id)
            ids[key] = doc_ids
        return ids, len(json_line)


def get_args(argv=None):
    parser = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    g = parser.add_argument_group(title="input data")
    g.add_argument("--input", type=str, required=True, help="Path to input JSON (one document per line)")
    g.add_argument("--json-keys", nargs="+", default=["text"], help="space separate listed of keys to extract from json")
    g.add_argument("--split-sentences", action="store_true", help="Split documents into sentences.")
    g.add_argument("--keep-newlines", action="store_true", help="Keep newlines between sentences when splitting.")
    g = parser.add_argument_group(title="tokenizer")
    g.add_argument("--tokenizer-name", type=str, required=True,
                   choices=["BertTokenizer", "GPT2Tokenizer", "T5Tokenizer", "RobertaTokenizer"])
    g.add_argument("--vocab-file", type=str, default=None, help="Path to the vocab file")
    g.add_argument("--merges-file", type=str, default=None, help="Path to the BPE merge file (if necessary).")
    g.add_argument("--do-lower-case", action="store_true", help="Whether to do lower case.")
    g.add_argument("--extra-ids", type=int, default=0, help="Number of extra ids.")
    g.add_argument("--append-eod", action="store_true", help="Append an <eod> token to the end of a document.")
    g.add_argument("--do-chinese-wwm", action="store_true", help="Whether to do whole word mask for Chinese.")
    g = parser.add_argument_group(title="output data")
    g.add_argument("--output-prefix", type=str, required=True, help="Path to binary output file without suffix")
    g.add_argument("--dataset-impl", type=str, default="mmap", choices=["lazy", "cached", "mmap"])
    g = parser.add_argument_group(title="runtime")
    g.add_argument("--workers", type=int, default=1, help="Number of worker processes to launch")
    g.add_argument("--log-interval", type=int, default=100, help="Interval between progress updates")
    args = parser.parse_args(argv)
    args.keep_empty = False
    if args.tokenizer_name.startswith("Bert") and not args.split_sentences:
        print("Bert tokenizer detected, are you sure you don't want to split sentences?")
    return args


def main(argv=None):
    args = get_args(argv)
    start = time.time()
    encoder = Encoder(args)
    tokenizer = build_tokenizer_from_args(args)
    level = "sentence" if args.split_sentences else "document"
    print(f"Vocab size: {tokenizer.vocab_size}")
    print(f"Output prefix: {args.output_prefix}")

    builders, bin_files, idx_files = {}, {}, {}
    for key in args.json_keys:
        bin_files[key] = f"{args.output_prefix}_{key}_{level}.bin"
        idx_files[key] = f"{args.output_prefix}_{key}_{level}.idx"
        builders[key] = indexed_dataset.make_builder(bin_files[key], impl=args.dataset_impl, vocab_size=len(tokenizer))

    fin = open(args.input, "r", encoding="utf-8")
    if args.workers > 1:
        pool = multiprocessing.Pool(args.workers, initializer=encoder.initializer)
        encoded = pool.imap(encoder.encode, fin, 25)
    else:
        encoder.initializer()
        encoded = map(encoder.encode, fin)

    proc_start, total_bytes = time.time(), 0
    print("Time to startup:", proc_start - start)
    i = 0
    for i, (doc, nbytes) in enumerate(encoded, start=1):
        total_bytes += nbytes
        for key, sentences in doc.items():
            if len(sentences) == 0:
                continue
            for sentence in sentences:
                import torch

                builders[key].add_item(torch.IntTensor(sentence))
            builders[key].end_document()
        if i % args.log_interval == 0:
            elapsed = time.time() - proc_start
            print(f"Processed {i} documents ({i / elapsed:.1f} docs/s, {total_bytes / elapsed / 1024 / 1024:.2f} MB/s).",
                  file=sys.stderr)
    print(f"Done! {i} documents")
    for key in args.json_keys:
        builders[key].finalize(idx_files[key])
    fin.close()


if __name__ == "__main__":
    main()
