"""HF T5 checkpoint → this project's parameter names (reference projects/T5/utils/weight_convert.py).

    python projects/T5/utils/weight_convert.py --hf <dir> --out <dir>
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..", "..", "..")))

import torch  # noqa: E402

from libai_b200.config import LazyConfig  # noqa: E402
from projects.MT5.utils.mt5_loader import T5LoaderHuggerFace  # noqa: E402
from projects.T5.models.t5_model import T5ForPreTraining  # noqa: E402

if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--hf", required=True)
    ap.add_argument("--out", required=True)
    ap.add_argument("--config", default="projects/T5/configs/t5_model_config.py")
    args = ap.parse_args()
    cfg = LazyConfig.load(args.config).cfg
    model = T5LoaderHuggerFace(T5ForPreTraining, cfg, args.hf).load()
    os.makedirs(args.out, exist_ok=True)
    torch.save(model.state_dict(), os.path.join(args.out, "model"))
    print("saved", os.path.join(args.out, "model"))
