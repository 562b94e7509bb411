"""Evaluate a (fine-tuned / adapter-tuned) ChatGLM checkpoint with the lm-evaluation-harness request types.

Spec: reference projects/ChatGLM/utils/eval_adapter.py:20-177 — an ``lm_eval`` ``BaseLM`` subclass plus a ``__main__``
that builds the topology, loads a HuggingFace- or LiBai-format checkpoint and runs a task list.  Here the request
types live once in :mod:`projects.Eval_LLM.eval_harness` (no hard ``lm_eval`` dependency, local JSONL tasks for offline
use); this module only binds them to the ChatGLM model / tokenizer / loaders and adds the command line.

    python -m torch.distributed.run --nproc-per-node 8 projects/ChatGLM/utils/eval_adapter.py \\
        --model-path /path/to/checkpoint --format libai --tensor-parallel 8 --tasks hellaswag --limit 200
"""
from __future__ import annotations

import argparse
from pathlib import Path
from typing import List, Optional

from libai_b200.config import DictConfig, instantiate
from libai_b200.utils import distributed as dist
from projects.Eval_LLM.eval_harness import EvalHarnessBase, run_eval_harness  # noqa: F401  (re-exported)
from projects.ChatGLM.configs.chatglm_config import cfg, tokenization
from projects.ChatGLM.chatglm import ChatGLMForConditionalGeneration
from projects.ChatGLM.utils.chatglm_loader import ChatGLMLoaderHuggerFace, ChatGLMLoaderLiBai


def load_model(model_path: str, fmt: str = "libai", **overrides):
    """``fmt``: "huggingface" (HF ``config.json`` + weights) or "libai" (a ``Checkpointer`` directory / model file)."""
    loader_cls = ChatGLMLoaderHuggerFace if fmt == "huggingface" else ChatGLMLoaderLiBai
    return loader_cls(model=ChatGLMForConditionalGeneration, libai_cfg=cfg, pretrained_model_path=model_path, **overrides).load()


def evaluate(model_path: str, fmt: str = "libai", eval_tasks: List[str] = ("hellaswag",), tokenizer_path: Optional[str] = None,
             data_parallel: int = 1, tensor_parallel: int = 1, pipeline_parallel: int = 1, batch_size_per_gpu: int = 1,
             limit: Optional[int] = None, save_filepath: Optional[Path] = None, device_type: str = "cuda"):
    dist.setup_dist_util(DictConfig(dict(
        data_parallel_size=data_parallel, tensor_parallel_size=tensor_parallel, pipeline_parallel_size=pipeline_parallel,
        pipeline_num_layers=cfg.hidden_layers, device_type=device_type)))
    if tokenizer_path is not None:
        tokenization.tokenizer.vocab_file = tokenizer_path
    tokenizer = instantiate(tokenization.tokenizer)
    model = load_model(model_path, fmt)
    return run_eval_harness(model, tokenizer, "chatglm", eval_tasks=list(eval_tasks), batch_size_per_gpu=batch_size_per_gpu,
                            save_filepath=save_filepath, limit=limit, cfg=cfg)


def main(argv=None):
    ap = argparse.ArgumentParser(description=__doc__.split("\n")[0])
    ap.add_argument("--model-path", required=True)
    ap.add_argument("--format", default="libai", choices=["libai", "huggingface"])
    ap.add_argument("--tokenizer-path", default=None)
    ap.add_argument("--tasks", nargs="+", default=["hellaswag"], help="lm_eval task names / globs, or local *.jsonl task files")
    ap.add_argument("--data-parallel", type=int, default=1)
    ap.add_argument("--tensor-parallel", type=int, default=1)
    ap.add_argument("--pipeline-parallel", type=int, default=1)
    ap.add_argument("--batch-size", type=int, default=1, help="per GPU")
    ap.add_argument("--limit", type=int, default=None)
    ap.add_argument("--save", type=Path, default=None)
    ap.add_argument("--device", default="cuda")
    a = ap.parse_args(argv)
    return evaluate(a.model_path, a.format, a.tasks, a.tokenizer_path, a.data_parallel, a.tensor_parallel,
                    a.pipeline_parallel, a.batch_size, a.limit, a.save, a.device)


if __name__ == "__main__":
    main()
