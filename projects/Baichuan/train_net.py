"""Baichuan training entry point: ``bash tools/train.sh projects/Baichuan/train_net.py projects/Baichuan/configs/baichuan_sft.py 8``."""
import os
import sys

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), "..", ".."))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))

from libai_b200.config import default_argument_parser  # noqa: E402
from train_net import main  # noqa: E402

if __name__ == "__main__":
    main(default_argument_parser().parse_args())
