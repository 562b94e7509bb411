"""Aquila causal LM (reference projects/Aquila/aquila.py): the Llama architecture —
RMSNorm, rotary attention, SwiGLU — on the shared native implementation (``libai_b200/models/llama_model.py``)."""
from libai_b200.config import configurable
from libai_b200.models.llama_model import (  # noqa: F401
    CasualMask,
    CrossEntropyLoss,
    LlamaAttention as MultiheadAttention,
    LlamaDecoderLayer as DecoderLayer,
    LlamaForCausalLM as _LlamaForCausalLM,
    LlamaMLP as MLP,
    LlamaModel as AquilaModel,
    RotaryEmbedding,
    SFTLoss,
)

AquilaCasualMask = CasualMask
AquilaDecoderLayer = DecoderLayer


class AquilaForCausalLM(_LlamaForCausalLM):
    @configurable
    def __init__(self, *args, cfg=None, **kwargs):
        kwargs.setdefault("qkv_bias", False)
        super().__init__(*args, **kwargs)
        self.cfg = cfg

    @classmethod
    def from_config(cls, cfg):
        out = _LlamaForCausalLM.from_config.__func__(cls, cfg)
        out["qkv_bias"] = cfg.get("qkv_bias", False)
        if cfg.get("rope_theta") is not None:
            out["rope_base"] = cfg.rope_theta
        return out
