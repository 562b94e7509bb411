"""HF / LiBai loaders for Qwen (reference projects/Qwen/utils/qwen2_loader.py)."""
from libai_b200.models.utils.model_loader.llama_loader import LlamaLoaderHuggerFace, LlamaLoaderLiBai


class Qwen2LoaderHuggerFace(LlamaLoaderHuggerFace):
    pass


class Qwen2LoaderLiBai(LlamaLoaderLiBai):
    pass
