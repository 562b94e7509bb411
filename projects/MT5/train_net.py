"""mT5 training entry point (reference projects/MT5/train_net.py): optionally start from HF weights
(``model.cfg.pretrained_model_path``)."""
import os
import sys

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), "..", ".."))
sys.path.insert(0, ROOT)

from libai_b200.config import LazyConfig, default_argument_parser, try_get_key  # noqa: E402
from libai_b200.engine import DefaultTrainer, default_setup  # noqa: E402


class Mt5Trainer(DefaultTrainer):
    @classmethod
    def build_model(cls, cfg):
        path = try_get_key(cfg, "model.cfg.pretrained_model_path")
        if path:
            from projects.MT5.utils.mt5_loader import T5LoaderHuggerFace

            return T5LoaderHuggerFace(cfg.model, cfg.model.cfg, path).load()
        return super().build_model(cfg)


def main(args):
    cfg = LazyConfig.apply_overrides(LazyConfig.load(args.config_file), args.opts)
    default_setup(cfg, args)
    return Mt5Trainer(cfg).train()


if __name__ == "__main__":
    main(default_argument_parser().parse_args())
