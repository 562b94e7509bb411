"""HF / LiBai checkpoint loaders (reference projects/Llama/utils/llama_loader.py)."""
from libai_b200.models.utils.model_loader.llama_loader import LlamaLoaderHuggerFace, LlamaLoaderLiBai  # noqa: F401
