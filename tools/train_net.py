#!/usr/bin/env python
"""Training / evaluation entry point.

Spec: reference tools/train_net.py:32-72 — load the LazyConfig, apply CLI overrides,
``default_setup``, per-rank seeding ``train.seed + rank``, ``--fast-dev-run`` (20 iterations, eval
every 10, log every iteration), ``--eval-only`` (load weights, run ``DefaultTrainer.test``),
otherwise ``DefaultTrainer(cfg).train()``.

    python tools/train_net.py --config-file configs/gpt2_synthetic.py train.train_iter=50
    bash tools/train.sh tools/train_net.py configs/gpt2_synthetic.py 8 train.dist.tensor_parallel_size=2
"""
import logging
import os
import random
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))

from libai_b200.config import LazyConfig, default_argument_parser, try_get_key  # noqa: E402
from libai_b200.engine import DefaultTrainer, default_setup  # noqa: E402
from libai_b200.utils import distributed as dutil  # noqa: E402
from libai_b200.utils.checkpoint import Checkpointer  # noqa: E402

logger = logging.getLogger("libai_b200." + __name__)


def main(args, trainer_cls=DefaultTrainer):
    """``trainer_cls``: a ``DefaultTrainer`` subclass (projects override ``build_model`` & co. and reuse this entry point)."""
    cfg = LazyConfig.load(args.config_file)
    cfg = LazyConfig.apply_overrides(cfg, args.opts)
    default_setup(cfg, args)

    seed = cfg.train.seed + dutil.get_rank()
    np.random.seed(seed)
    random.seed(seed)
    # device RNG (dropout masks): same stream on all tensor-parallel ranks of a replica, see dutil.model_parallel_seed
    torch.manual_seed(dutil.model_parallel_seed(cfg.train.seed))

    if args.fast_dev_run:
        cfg.train.train_epoch = 0
        cfg.train.train_iter = 20
        cfg.train.evaluation.eval_period = 10
        cfg.train.log_period = 1

    if args.eval_only:
        tokenizer = None
        if try_get_key(cfg, "tokenization") is not None:
            tokenizer = trainer_cls.build_tokenizer(cfg)
        model = trainer_cls.build_model(cfg)
        Checkpointer(model, save_dir=cfg.train.output_dir).resume_or_load(cfg.train.load_weight, resume=args.resume)
        test_loader = trainer_cls.build_test_loader(cfg, tokenizer)
        if len(test_loader) == 0:
            logger.info("No dataset in dataloader.test, please set dataset for dataloader.test")
        _ = trainer_cls.test(cfg, test_loader, model)
        return

    trainer = trainer_cls(cfg)
    return trainer.train()


if __name__ == "__main__":
    args = default_argument_parser().parse_args()
    main(args)
