"""Download helper for the CLUE benchmark files (reference projects/text_classification/dataset/download_clue_data.py).
Needs network access; files are fetched through ``libai_b200.utils.file_utils.get_data_from_cache``."""
import argparse
import os
import sys

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..", "..", "..")))

from libai_b200.utils.file_utils import get_data_from_cache  # noqa: E402

if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--url", required=True, help="archive or file URL of the task")
    ap.add_argument("--data_dir", default="./projects/text_classification/dataset/clue_data")
    args = ap.parse_args()
    os.makedirs(args.data_dir, exist_ok=True)
    print(get_data_from_cache(args.url, args.data_dir))
