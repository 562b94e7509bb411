"""Alias of :mod:`projects.Qwen.configs.qwen2_config` under the file name the reference uses
(projects/Qwen/configs/qwen_config.py) so existing command lines keep working."""
from projects.Qwen.configs.qwen2_config import cfg, model, tokenization  # noqa: F401
