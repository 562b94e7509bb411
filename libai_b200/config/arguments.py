"""Command line of the training / evaluation entry points.

Spec: reference libai/config/arguments.py:21-80 — ``--config-file --resume --eval-only
--fast-dev-run`` followed by free ``key=value`` overrides.
"""
import argparse
import sys


def default_argument_parser(epilog=None):
    prog = sys.argv[0] if sys.argv else "train_net.py"
    parser = argparse.ArgumentParser(
        epilog=epilog
        or f"""
Examples:

Run on a single GPU:
    $ python {prog} --config-file configs/gpt2_pretrain.py

Run on 8 GPUs of one node (one process per GPU over NCCL):
    $ bash tools/train.sh {prog} configs/gpt2_pretrain.py 8 train.dist.tensor_parallel_size=2

Change some config options:
    $ python {prog} --config-file cfg.py train.load_weight=/path/to/weight optim.lr=0.001
""",
        formatter_class=argparse.RawDescriptionHelpFormatter,
    )
    parser.add_argument("--config-file", default="", metavar="FILE", help="path to config file")
    parser.add_argument(
        "--resume",
        action="store_true",
        help="resume from train.output_dir/last_checkpoint (model, optimizer, scheduler, sampler)",
    )
    parser.add_argument("--eval-only", action="store_true", help="perform evaluation only")
    parser.add_argument(
        "--fast-dev-run",
        action="store_true",
        help="run 20 training iterations with evaluation every 10 and logging every iteration",
    )
    parser.add_argument(
        "opts",
        help="config overrides in the form key=value (see LazyConfig.apply_overrides)",
        default=None,
        nargs=argparse.REMAINDER,
    )
    return parser
