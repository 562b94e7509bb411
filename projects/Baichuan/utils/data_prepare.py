"""Tokenise an Alpaca-format instruction file for Baichuan SFT (reference projects/Baichuan/utils/).

    python projects/Baichuan/utils/data_prepare.py --data alpaca.json --out data_baichuan [tokenizer files…]
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..", "..", "..")))

from libai_b200.config import LazyConfig, instantiate  # noqa: E402
from projects.common.sft import prepare_sft_corpus  # noqa: E402


def main(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("--data", required=True)
    ap.add_argument("--out", default="data_baichuan")
    ap.add_argument("--config", default="projects/Baichuan/configs/baichuan_config.py")
    ap.add_argument("--max-seq-length", type=int, default=512)
    ap.add_argument("opts", nargs="*", help="tokenization.tokenizer.<key>=<value> overrides")
    args = ap.parse_args(argv)
    cfg = LazyConfig.apply_overrides(LazyConfig.load(args.config), args.opts)
    tokenizer = instantiate(cfg.tokenization.tokenizer)
    print(prepare_sft_corpus(args.data, args.out, tokenizer, args.max_seq_length))


if __name__ == "__main__":
    main()
