"""Adapter fine-tuning entry point (reference projects/Llama/adapter/train_net.py): load the HF weights, freeze the
backbone, train the prompts + gates with the default trainer."""
import os
import sys

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), "..", "..", ".."))
sys.path.insert(0, ROOT)

from libai_b200.config import LazyConfig, default_argument_parser, try_get_key  # noqa: E402
from libai_b200.engine import DefaultTrainer, default_setup  # noqa: E402


class AdapterTrainer(DefaultTrainer):
    @classmethod
    def build_model(cls, cfg):
        path = try_get_key(cfg, "model.cfg.pretrained_model_path")
        if path and os.path.isdir(path):
            from projects.Llama.utils.llama_loader import LlamaLoaderHuggerFace

            model = LlamaLoaderHuggerFace(cfg.model, cfg.model.cfg, path).load()
        else:
            model = super().build_model(cfg)
        return model.freeze_backbone()


LlamaTrainer = AdapterTrainer      # name used by the reference's adapter entry point


def main(args):
    cfg = LazyConfig.load(args.config_file)
    cfg = LazyConfig.apply_overrides(cfg, args.opts)
    default_setup(cfg, args)
    return AdapterTrainer(cfg).train()


if __name__ == "__main__":
    main(default_argument_parser().parse_args())
