"""Export a GPT-2 LM (reference libai/onnx_export/gpt2_to_onnx.py).

    python -m libai_b200.onnx_export.gpt2_to_onnx --config configs/gpt2_pretrain.py --checkpoint output/gpt2_output \\
        --output output/gpt2.onnx
"""
import argparse

import torch

from libai_b200.config import LazyConfig
from libai_b200.models import GPTForPreTraining
from libai_b200.models.utils.model_loader import GPT2LoaderLiBai
from libai_b200.onnx_export.export import ExportWrapper, export_model
from libai_b200.utils import distributed as dist


class gpt2Graph(ExportWrapper):
    """The traced callable: ``input_ids -> logits`` (the reference wraps the eager model in an ``nn.Graph`` of the same
    name, libai/onnx_export/gpt2_to_onnx.py:41-53; here the static graph is whatever the exporter traces)."""

    def __init__(self, eager_model):
        super().__init__(eager_model, ["input_ids"], "prediction_scores")


def get_model(config_file, checkpoint=None):
    cfg = LazyConfig.load(config_file)
    dist.setup_dist_util(cfg.train.dist)
    if checkpoint is None:
        return GPTForPreTraining(cfg.model.cfg)
    return GPT2LoaderLiBai(GPTForPreTraining, cfg.model.cfg, checkpoint).load()


def main(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="configs/gpt2_pretrain.py")
    ap.add_argument("--checkpoint", default=None)
    ap.add_argument("--output", default="output/gpt2.onnx")
    ap.add_argument("--seq-len", type=int, default=5)
    args = ap.parse_args(argv)
    model = get_model(args.config, args.checkpoint)
    ids = torch.ones(1, args.seq_len, dtype=torch.long)
    fmt, file = export_model(model, {"input_ids": ids}, "prediction_scores", args.output,
                             dynamic_axes={"input_ids": {0: "batch", 1: "seq"}, "prediction_scores": {0: "batch", 1: "seq"}})
    print(f"exported {fmt}: {file}")


if __name__ == "__main__":
    main()
