"""Evaluate an LLM of one of the projects on lm-evaluation-harness tasks or local JSONL tasks.

Spec: reference projects/Eval_LLM/main.py:11-88.  Edit ``config.py`` (or pass overrides) and run

    bash tools/infer.sh projects/Eval_LLM/main.py 1 eval_config.model_type=llama \
        eval_config.pretrained_model_path=/data/Llama-2-7b-hf eval_config.hf_tokenizer_path=/data/Llama-2-7b-hf \
        eval_config.model_weight_type=huggingface

``model_type`` selects an entry of ``special_arguments.json`` (model class, config file and loader per family).
"""
import importlib
import json
import os
import sys

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), "..", ".."))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from libai_b200.config import LazyCall, LazyConfig  # noqa: E402
from libai_b200.models.utils.model_loader.base_loader import ModelLoaderLiBai  # noqa: E402
from libai_b200.utils import distributed as dist  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


class LLMLoaderLibai(ModelLoaderLiBai):
    def __init__(self, model, libai_cfg, pretrained_model_path, base_model_prefix, **kwargs):
        super().__init__(model, libai_cfg, pretrained_model_path, **kwargs)
        self.base_model_prefix_2 = base_model_prefix


def get_special_arguments(model_type):
    with open(os.path.join(HERE, "special_arguments.json"), "r") as f:
        arguments = json.load(f)
    assert model_type in arguments, f"unknown model_type {model_type!r}; known: {sorted(arguments)}"
    return arguments[model_type]


def build_model(eval_config, model_cfg, special_arguments):
    model_class = getattr(importlib.import_module(special_arguments["model_class_prefix"]),
                          special_arguments["model_class"])
    weight_type = eval_config.model_weight_type
    assert weight_type in ("huggingface", "libai", "random"), "model_weight_type must be huggingface, libai or random"
    if weight_type == "random":
        return model_class(model_cfg.cfg)
    if weight_type == "huggingface":
        loader = getattr(importlib.import_module(special_arguments["huggingface_loader_prefix"]),
                         special_arguments["huggingface_loader"])
        return loader(model_class, model_cfg.cfg, eval_config.pretrained_model_path).load()
    return LLMLoaderLibai(model_class, model_cfg.cfg, eval_config.pretrained_model_path,
                          special_arguments["base_model_prefix_2"]).load()


def build_tokenizer(eval_config, model_cfg):
    """HF tokenizer from ``hf_tokenizer_path`` (like the reference); falls back to the project's own tokenizer from
    its config when no HF path is given."""
    path = eval_config.hf_tokenizer_path
    if path:
        from transformers import AutoTokenizer

        tokenizer = AutoTokenizer.from_pretrained(path, trust_remote_code=True)
        cfg_json = os.path.join(path, "config.json")
        generation_config = json.load(open(cfg_json)) if os.path.exists(cfg_json) else {}
        if tokenizer.pad_token_id is None and generation_config.get("pad_token_id") is not None:
            tokenizer.pad_token_id = generation_config["pad_token_id"]
        if tokenizer.eos_token_id is None and generation_config.get("eos_token_id") is not None:
            tokenizer.eos_token_id = generation_config["eos_token_id"]
        return tokenizer
    from libai_b200.config import instantiate

    return instantiate(model_cfg.tokenization.tokenizer)


def main(argv=None):
    argv = list(sys.argv[1:] if argv is None else argv)
    cfg = LazyConfig.load(os.path.join(HERE, "config.py"))
    cfg = LazyConfig.apply_overrides(cfg, argv)
    if not torch.cuda.is_available():
        cfg.parallel_config.device_type = "cpu"
    dist.setup_dist_util(cfg.parallel_config)
    special_arguments = get_special_arguments(cfg.eval_config.model_type)
    print("Loading Model...")
    model_cfg = LazyConfig.load(os.path.join(ROOT, special_arguments["config_path"]))
    if model_cfg.cfg.get("max_position_embeddings", None) is None:
        model_cfg.cfg.max_position_embeddings = 1024
    model = build_model(cfg.eval_config, model_cfg, special_arguments)
    tokenizer = build_tokenizer(cfg.eval_config, model_cfg)
    if torch.cuda.is_available():
        model = model.to(torch.device("cuda", torch.cuda.current_device()))
    print("Model Loaded!")

    from projects.Eval_LLM.eval_harness import run_eval_harness

    return run_eval_harness(model, tokenizer, cfg.eval_config.model_type, eval_tasks=list(cfg.eval_config.eval_tasks),
                            batch_size_per_gpu=cfg.eval_config.batch_size_per_gpu, limit=cfg.eval_config.limit,
                            save_filepath=cfg.eval_config.save_filepath, cfg=model_cfg.cfg)


if __name__ == "__main__":
    main()
