"""GPT-2 loaders (reference libai/models/utils/model_loader/gpt_loader.py:22-174): HF ``Conv1D`` weights are
``[in, out]`` → transposed; ``c_attn`` is ``[q; k; v]`` → per-head interleaved rows."""
import collections
import re

from .base_loader import ModelLoaderHuggerFace, ModelLoaderLiBai


class GPT2LoaderHuggerFace(ModelLoaderHuggerFace):
    def __init__(self, model, libai_cfg, pretrained_model_path, **kwargs):
        super().__init__(model, libai_cfg, pretrained_model_path, **kwargs)
        self.base_model_prefix_1 = "transformer"
        self.base_model_prefix_2 = "GPT_model"

    def _convert_state_dict(self, sd, cfg):
        heads, hidden = cfg.get("num_attention_heads"), cfg.get("hidden_size")
        head_size = hidden // heads
        out = collections.OrderedDict()
        p2 = "GPT_model."
        for key, v in sd.items():
            k = key[len("transformer."):] if key.startswith("transformer.") else key
            if k.endswith(".attn.bias") or k.endswith(".attn.masked_bias") or k == "lm_head.weight":
                continue  # causal-mask buffers / tied head
            if k == "wte.weight":
                out[p2 + "embeddings.token_embeddings.weight"] = v
            elif k == "wpe.weight":
                out[p2 + "embeddings.position_embeddings.weight"] = v
            elif k.startswith("ln_f."):
                out[p2 + "transformer.layernorm_f." + k[5:]] = v
            else:
                m = re.match(r"^h\.(\d+)\.(.+)\.(weight|bias)$", k)
                if not m:
                    out[key] = v
                    continue
                i, name, kind = m.groups()
                base = f"{p2}transformer.layers.{i}."
                if kind == "weight" and v.dim() == 2 and name in ("attn.c_attn", "attn.c_proj", "mlp.c_fc", "mlp.c_proj"):
                    v = v.t()
                if name == "attn.c_attn":
                    v = self._fix_qkv_ordering(v.contiguous(), head_size, heads)
                target = {
                    "ln_1": "input_layernorm", "ln_2": "post_attention_layernorm",
                    "attn.c_attn": "self_attention.query_key_value", "attn.c_proj": "self_attention.dense",
                    "mlp.c_fc": "mlp.dense_h_to_4h", "mlp.c_proj": "mlp.dense_4h_to_h",
                }.get(name)
                out[(base + target + "." + kind) if target else key] = v
        return out

    def _load_config_from_json(self, config_file):
        cfg = self._read_config_json()
        self._map_config(cfg, {
            "n_layer": "hidden_layers", "n_embd": "hidden_size", "n_head": "num_attention_heads",
            "n_positions": "max_seq_length", "embd_pdrop": "embedding_dropout_prob",
            "attn_pdrop": "attention_dropout_prob", "resid_pdrop": "output_dropout_prob",
            "layer_norm_epsilon": "layernorm_epsilon", "vocab_size": "vocab_size",
            "initializer_range": "initializer_range",
        })
        n_inner = cfg.get("n_inner")
        self._update_cfg("ffn_hidden_size", n_inner if n_inner is not None else 4 * cfg["n_embd"])


class GPT2LoaderLiBai(ModelLoaderLiBai):
    def __init__(self, model, libai_cfg, pretrained_model_path, **kwargs):
        super().__init__(model, libai_cfg, pretrained_model_path, **kwargs)
        self.base_model_prefix_2 = "GPT_model"
