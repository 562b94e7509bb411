"""Fine-tune Stable Diffusion (full, DreamBooth, or LoRA) and export an inference checkpoint.

Spec: reference projects/Stable_Diffusion/train_net.py:35-147 — the default trainer plus an ``after_train`` hook that
writes ``<output_dir>/model_sd_for_inference`` (whole pipeline in the diffusers folder layout, or only the LoRA
attention-processor weights), per-rank seeding, and no periodic full checkpoints in LoRA mode.

    bash tools/train.sh projects/Stable_Diffusion/train_net.py projects/Stable_Diffusion/configs/lora_config.py 8
"""
import logging
import os
import random
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..", "..")))

from libai_b200.config import LazyConfig, default_argument_parser, try_get_key  # noqa: E402
from libai_b200.engine import DefaultTrainer, default_setup, hooks  # noqa: E402
from libai_b200.engine.trainer import HookBase  # noqa: E402
from libai_b200.utils import distributed as dist  # noqa: E402
from libai_b200.utils.checkpoint import Checkpointer  # noqa: E402

logger = logging.getLogger("libai_b200." + __name__)


class SdCheckpointer(HookBase):
    def __init__(self, model, save_path):
        self._model, self._save_path = model, save_path

    def after_train(self):
        from projects.Stable_Diffusion.modules.lora import save_attn_procs
        from projects.Stable_Diffusion.pipeline import StableDiffusionPipeline

        model = self._model.module if hasattr(self._model, "module") else self._model
        save_path = os.path.join(self._save_path, "model_sd_for_inference")
        logger.info(f"saving stable diffusion model to {save_path}")
        if not dist.is_main_process():
            return
        if hasattr(model, "lora_layers"):
            save_attn_procs(model.lora_layers, save_path)
        else:
            StableDiffusionPipeline(model.tokenizer, model.text_encoder, model.vae, model.unet,
                                    model.noise_scheduler).save_pretrained(save_path)


class Trainer(DefaultTrainer):
    def build_hooks(self):
        ret = [hooks.IterationTimer(), hooks.LRScheduler(), SdCheckpointer(self.model, self.cfg.train.output_dir)]
        if not try_get_key(self.cfg, "model.train_with_lora", default=False):
            ret.append(hooks.PeriodicCheckpointer(self.checkpointer, self.cfg.train.checkpointer.period))
        if dist.is_main_process():
            ret.append(hooks.PeriodicWriter(self.build_writers(), self.cfg.train.log_period))
        return ret


def main(args):
    cfg = LazyConfig.load(args.config_file)
    cfg = LazyConfig.apply_overrides(cfg, args.opts)
    default_setup(cfg, args)

    seed_for_rank = cfg.train.seed + dist.get_rank()       # every rank draws different noise / timesteps
    torch.manual_seed(seed_for_rank)
    np.random.seed(seed_for_rank)
    random.seed(seed_for_rank)

    if args.fast_dev_run:
        cfg.train.train_epoch = 0
        cfg.train.train_iter = 20
        cfg.train.evaluation.eval_period = 10
        cfg.train.log_period = 1

    if args.eval_only:
        model = Trainer.build_model(cfg)
        Checkpointer(model, save_dir=cfg.train.output_dir).resume_or_load(cfg.train.load_weight, resume=args.resume)
        test_loader = Trainer.build_test_loader(cfg, None)
        if len(test_loader) == 0:
            logger.info("No dataset in dataloader.test, please set dataset for dataloader.test")
        Trainer.test(cfg, test_loader, model)
        return

    trainer = Trainer(cfg)
    return trainer.train()


if __name__ == "__main__":
    main(default_argument_parser().parse_args())
