"""HF / LiBai loaders for Aquila (reference projects/Aquila/utils/aquila_loader.py)."""
from libai_b200.models.utils.model_loader.llama_loader import LlamaLoaderHuggerFace, LlamaLoaderLiBai


class AquilaLoaderHuggerFace(LlamaLoaderHuggerFace):
    pass


class AquilaLoaderLiBai(LlamaLoaderLiBai):
    pass
