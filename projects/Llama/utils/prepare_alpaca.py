"""Tokenise the Alpaca instruction data for SFT (reference projects/Llama/utils/prepare_alpaca.py).

    python projects/Llama/utils/prepare_alpaca.py --data alpaca_data_cleaned.json --tokenizer <dir>/tokenizer.model \\
        --out alpaca_data --max-seq-length 512
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..", "..", "..")))

from projects.common.sft import prepare_sft_corpus  # noqa: E402
from projects.Llama.tokenizer import LlamaTokenizer  # noqa: E402


def main(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("--data", required=True, help="alpaca-format json (instruction / input / output)")
    ap.add_argument("--tokenizer", required=True, help="sentencepiece tokenizer.model")
    ap.add_argument("--out", default="alpaca_data")
    ap.add_argument("--max-seq-length", type=int, default=512)
    ap.add_argument("--test-split-size", type=int, default=2000)
    ap.add_argument("--no-mask-inputs", action="store_true")
    args = ap.parse_args(argv)
    n_train, n_test = prepare_sft_corpus(args.data, args.out, LlamaTokenizer(args.tokenizer), args.max_seq_length,
                                         args.test_split_size, not args.no_mask_inputs)
    print(f"train has {n_train:,} samples, test has {n_test:,} samples → {args.out}")


if __name__ == "__main__":
    main()
