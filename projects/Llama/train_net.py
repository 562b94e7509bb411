"""Llama training entry point (reference projects/Llama/train_net.py): the generic ``tools/train_net.py`` flow with a
trainer whose ``build_model`` starts from the pretrained HuggingFace checkpoint when
``cfg.model.cfg.pretrained_model_path`` points at one (supervised fine-tuning), else from random initialisation.

    bash tools/train.sh projects/Llama/train_net.py projects/Llama/configs/llama_sft.py 8
"""
import logging
import os
import sys

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), "..", ".."))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))

from libai_b200.config import default_argument_parser, try_get_key  # noqa: E402
from libai_b200.engine import DefaultTrainer  # noqa: E402
from train_net import main  # noqa: E402

logger = logging.getLogger("libai_b200." + __name__)


class LlamaTrainer(DefaultTrainer):
    @classmethod
    def construct_model(cls, cfg):
        path = try_get_key(cfg, "model.cfg.pretrained_model_path")
        if path and os.path.isdir(path) and os.path.exists(os.path.join(path, "config.json")):
            from projects.Llama.utils.llama_loader import LlamaLoaderHuggerFace

            logger.info("loading pretrained weights from %s", path)
            return LlamaLoaderHuggerFace(cfg.model, cfg.model.cfg, path).load()
        return super().construct_model(cfg)


if __name__ == "__main__":
    main(default_argument_parser().parse_args(), trainer_cls=LlamaTrainer)
