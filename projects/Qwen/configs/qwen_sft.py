"""Alias of :mod:`projects.Qwen.configs.qwen2_sft` under the reference's file name (projects/Qwen/configs/qwen_sft.py)."""
from projects.Qwen.configs.qwen2_sft import *  # noqa: F401,F403
from projects.Qwen.configs.qwen2_sft import dataloader, graph, model, optim, tokenization, train  # noqa: F401
