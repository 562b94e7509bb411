"""Llama model for the project configs (reference projects/Llama/llama.py); the implementation lives in the core
model zoo (``libai_b200/models/llama_model.py``)."""
from libai_b200.models.llama_model import (  # noqa: F401
    CasualMask,
    CrossEntropyLoss,
    LlamaAttention as MultiheadAttention,
    LlamaDecoderLayer,
    LlamaForCausalLM,
    LlamaMLP as MLP,
    LlamaModel,
    RotaryEmbedding,
    SFTLoss,
    rotary_tables,
)
