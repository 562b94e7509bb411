"""Export the T5 encoder-decoder (reference libai/onnx_export/t5_to_onnx.py): inputs are the five tensors of the
pre-training forward (ids + the three attention masks), output the LM logits."""
import argparse

import torch

from libai_b200.config import LazyConfig
from libai_b200.models import T5ForPreTraining
from libai_b200.models.utils.model_loader.base_loader import ModelLoaderLiBai
from libai_b200.onnx_export.export import ExportWrapper, export_model
from libai_b200.utils import distributed as dist


class t5Graph(ExportWrapper):
    """The traced callable of the T5 export: the five encoder/decoder inputs -> logits (reference
    libai/onnx_export/t5_to_onnx.py nn.Graph of the same name)."""

    def __init__(self, eager_model):
        super().__init__(eager_model, ["encoder_input_ids", "decoder_input_ids", "encoder_attn_mask",
                                       "decoder_attn_mask", "encoder_decoder_attn_mask"], "prediction_scores")


def get_model(config_file, checkpoint=None):
    cfg = LazyConfig.load(config_file)
    dist.setup_dist_util(cfg.train.dist)
    if checkpoint is None:
        return T5ForPreTraining(cfg.model.cfg)
    return ModelLoaderLiBai(T5ForPreTraining, cfg.model.cfg, checkpoint).load()


def example_inputs(batch=1, enc_len=5, dec_len=3):
    return {
        "encoder_input_ids": torch.ones(batch, enc_len, dtype=torch.long),
        "decoder_input_ids": torch.ones(batch, dec_len, dtype=torch.long),
        "encoder_attn_mask": torch.ones(batch, enc_len, enc_len, dtype=torch.bool),
        "decoder_attn_mask": torch.ones(batch, dec_len, dec_len, dtype=torch.bool).tril(),
        "encoder_decoder_attn_mask": torch.ones(batch, dec_len, enc_len, dtype=torch.bool),
    }


def main(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="configs/t5_large_pretrain.py")
    ap.add_argument("--checkpoint", default=None)
    ap.add_argument("--output", default="output/t5.onnx")
    args = ap.parse_args(argv)
    model = get_model(args.config, args.checkpoint)
    fmt, file = export_model(model, example_inputs(), "prediction_scores", args.output)
    print(f"exported {fmt}: {file}")


if __name__ == "__main__":
    main()
