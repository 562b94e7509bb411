"""MAE entry point (reference projects/MAE/train_net.py): optional fine-tuning from a pre-trained MAE checkpoint
(this framework's or the official PyTorch one)."""
import logging
import os
import sys

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), "..", ".."))
sys.path.insert(0, ROOT)

from libai_b200.config import LazyConfig, default_argument_parser, try_get_key  # noqa: E402
from libai_b200.engine import DefaultTrainer, default_setup  # noqa: E402

logger = logging.getLogger("libai_b200.mae")


class Trainer(DefaultTrainer):
    @classmethod
    def build_model(cls, cfg):
        model = super().build_model(cfg)
        ft = try_get_key(cfg, "finetune")
        if ft is not None and ft.enable:
            logger.info(f"Loading pretrained weight ({ft.weight_style}) for finetuning: {ft.path}")
            if ft.weight_style == "pytorch":
                from projects.MAE.utils.weight_convert import load_torch_checkpoint

                load_torch_checkpoint(model, ft.path, num_heads=cfg.model.num_heads)
            else:
                from libai_b200.utils.checkpoint import Checkpointer

                Checkpointer(model).load(ft.path, checkpointables=[])
        return model


def main(args):
    cfg = LazyConfig.apply_overrides(LazyConfig.load(args.config_file), args.opts)
    default_setup(cfg, args)
    if args.eval_only:
        model = Trainer.build_model(cfg)
        return Trainer.test(cfg, Trainer.build_test_loader(cfg, None), model)
    return Trainer(cfg).train()


if __name__ == "__main__":
    main(default_argument_parser().parse_args())
