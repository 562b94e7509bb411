"""Fetch the GPT-2 demo corpus used by the PaLM recipe (reference projects/PaLM/tools/download_demo_dataset.py)."""
import argparse
import os
import sys

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..", "..", "..")))

from libai_b200.utils.file_utils import get_data_from_cache  # noqa: E402

BASE = "https://oneflow-static.oss-cn-beijing.aliyuncs.com/ci-files/dataset/libai/gpt_dataset/"
FILES = ["gpt2-vocab.json", "gpt2-merges.txt", "loss_compara_content_sentence.bin", "loss_compara_content_sentence.idx"]

if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("-o", "--output", default="./projects/PaLM/gpt_dataset")
    args = ap.parse_args()
    os.makedirs(args.output, exist_ok=True)
    for f in FILES:
        print("fetching", f)
        get_data_from_cache(BASE + f, args.output)
