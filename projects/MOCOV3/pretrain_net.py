"""MoCo v3 entry point (reference projects/MOCOV3/pretrain_net.py)."""
import os
import sys

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), "..", ".."))
sys.path.insert(0, ROOT)

from libai_b200.config import LazyConfig, default_argument_parser, try_get_key  # noqa: E402
from libai_b200.engine import DefaultTrainer, default_setup  # noqa: E402


class MoCoPretrainingTrainer(DefaultTrainer):
    @classmethod
    def build_model(cls, cfg):
        # the cosine momentum schedule spans the whole run
        if try_get_key(cfg, "model.max_iter") is not None:
            cfg.model.max_iter = cfg.train.train_iter
        model = super().build_model(cfg)
        probe = try_get_key(cfg, "model.linear_prob")
        if probe:
            from projects.MOCOV3.utils.load_checkpoint import load_checkpoint

            load_checkpoint(model, probe["path"], probe.get("weight_style", "oneflow"), cfg.model.num_heads)
            for n, p in model.named_parameters():
                p.requires_grad = n.startswith("head")
        return model


def main(args):
    cfg = LazyConfig.apply_overrides(LazyConfig.load(args.config_file), args.opts)
    default_setup(cfg, args)
    return MoCoPretrainingTrainer(cfg).train()


if __name__ == "__main__":
    main(default_argument_parser().parse_args())
