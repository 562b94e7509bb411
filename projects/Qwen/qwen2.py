"""Qwen causal LM (reference projects/Qwen/qwen2.py): the Llama architecture with biases on the fused q/k/v projection —
RMSNorm, rotary attention, SwiGLU — on the shared native implementation (``libai_b200/models/llama_model.py``)."""
from libai_b200.config import configurable
from libai_b200.models.llama_model import (  # noqa: F401
    CasualMask,
    CrossEntropyLoss,
    LlamaAttention as MultiheadAttention,
    LlamaDecoderLayer as DecoderLayer,
    LlamaForCausalLM as _LlamaForCausalLM,
    LlamaMLP as MLP,
    LlamaModel as Qwen2Model,
    RotaryEmbedding,
    SFTLoss,
)


class Qwen2ForCausalLM(_LlamaForCausalLM):
    @configurable
    def __init__(self, *args, cfg=None, **kwargs):
        kwargs.setdefault("qkv_bias", True)
        super().__init__(*args, **kwargs)
        self.cfg = cfg

    @classmethod
    def from_config(cls, cfg):
        out = _LlamaForCausalLM.from_config.__func__(cls, cfg)
        out["qkv_bias"] = cfg.get("qkv_bias", True)
        if cfg.get("rope_theta") is not None:
            out["rope_base"] = cfg.rope_theta
        return out
