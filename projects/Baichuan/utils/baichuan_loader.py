"""HF / LiBai loaders for Baichuan (reference projects/Baichuan/utils/baichuan_loader.py)."""
from libai_b200.models.utils.model_loader.llama_loader import LlamaLoaderHuggerFace, LlamaLoaderLiBai


class BaichuanLoaderHuggerFace(LlamaLoaderHuggerFace):
    """HF Baichuan fuses q/k/v into ``W_pack`` ([q; k; v] rows) → per-head interleaved ``query_key_value``."""

    def _convert_state_dict(self, sd, cfg):
        import collections

        sd = collections.OrderedDict(sd)
        heads, hidden = cfg.get("num_attention_heads"), cfg.get("hidden_size")
        for key in [k for k in sd if k.endswith("self_attn.W_pack.weight")]:
            new = key.replace("W_pack", "query_key_value")
            sd[new] = self._fix_qkv_ordering(sd.pop(key), hidden // heads, heads)
        return super()._convert_state_dict(sd, cfg)


class BaichuanLoaderLiBai(LlamaLoaderLiBai):
    pass
