"""Fetch the QQP tsv files + vocabulary (reference projects/QQP/dataset/download_qqp_data.py; needs network)."""
import os
import sys

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..", "..", "..")))

from libai_b200.utils.file_utils import get_data_from_cache  # noqa: E402

BASE = "https://oneflow-static.oss-cn-beijing.aliyuncs.com/ci-files/dataset/libai/QQP/"
FILES = ["train.tsv", "dev.tsv", "bert-base-chinese-vocab.txt"]

if __name__ == "__main__":
    out = "./projects/QQP/QQP_DATA"
    os.makedirs(out, exist_ok=True)
    for f in FILES:
        print(get_data_from_cache(BASE + f, out))
