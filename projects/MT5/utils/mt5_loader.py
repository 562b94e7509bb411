"""HF T5 / mT5 → ``MT5Model`` loader (reference projects/MT5/utils/mt5_loader.py): q/k/v fused per head
(``[a, 3, d]`` rows), cross-attention k/v fused (``[a, 2, d]``), v1.0 ``wi`` / v1.1 ``wi_0, wi_1`` MLPs, shared or
separate LM head."""
import collections
import re

import torch

from libai_b200.models.utils.model_loader.base_loader import ModelLoaderHuggerFace, ModelLoaderLiBai


class T5LoaderHuggerFace(ModelLoaderHuggerFace):
    def __init__(self, model, libai_cfg, pretrained_model_path, **kwargs):
        super().__init__(model, libai_cfg, pretrained_model_path, **kwargs)
        self.base_model_prefix_1 = "transformer"
        self.base_model_prefix_2 = "mt5_model"

    def _convert_state_dict(self, sd, cfg):
        sd = collections.OrderedDict(sd)
        heads, head_size, layers = cfg.get("num_attention_heads"), cfg.get("head_size"), cfg.get("hidden_layers")
        out = collections.OrderedDict()
        out["embedding.word_embeddings.weight"] = sd.pop("shared.weight")
        for k in ("encoder.embed_tokens.weight", "decoder.embed_tokens.weight"):
            sd.pop(k, None)
        head = sd.pop("lm_head.weight", None)
        if cfg.get("model_type", "mt5") == "mt5":  # separate head module; tied checkpoints reuse the embedding matrix
            out["lm_head.weight"] = head if head is not None else out["embedding.word_embeddings.weight"]
        for stack in ("encoder", "decoder"):
            for i in range(layers):
                src = f"{stack}.block.{i}.layer"
                dst = f"{stack}.layers.{i}"
                sa = f"{src}.0.SelfAttention"
                qkv = torch.cat([sd.pop(f"{sa}.{n}.weight") for n in ("q", "k", "v")], dim=0)
                out[f"{dst}.self_attention.query_key_value.weight"] = self._fix_qkv_ordering(qkv, head_size, heads)
                out[f"{dst}.self_attention.dense.weight"] = sd.pop(f"{sa}.o.weight")
                if f"{sa}.relative_attention_bias.weight" in sd:
                    out[f"{dst}.self_attention.relative_attention_bias"] = sd.pop(f"{sa}.relative_attention_bias.weight")
                out[f"{dst}.input_layernorm.weight"] = sd.pop(f"{src}.0.layer_norm.weight")
                mlp_idx = 1
                if stack == "decoder":
                    ca = f"{src}.1.EncDecAttention"
                    out[f"{dst}.cross_attention.query.weight"] = sd.pop(f"{ca}.q.weight")
                    kv = torch.cat([sd.pop(f"{ca}.k.weight"), sd.pop(f"{ca}.v.weight")], dim=0)
                    out[f"{dst}.cross_attention.key_value.weight"] = self._fix_qkv_ordering(kv, head_size, heads)
                    out[f"{dst}.cross_attention.dense.weight"] = sd.pop(f"{ca}.o.weight")
                    out[f"{dst}.post_attention_layernorm.weight"] = sd.pop(f"{src}.1.layer_norm.weight")
                    out[f"{dst}.post_cross_attention_layernorm.weight"] = sd.pop(f"{src}.2.layer_norm.weight")
                    mlp_idx = 2
                else:
                    out[f"{dst}.post_attention_layernorm.weight"] = sd.pop(f"{src}.1.layer_norm.weight")
                ff = f"{src}.{mlp_idx}.DenseReluDense"
                if f"{ff}.wi.weight" in sd:
                    out[f"{dst}.mlp.dense_h_to_4h.weight"] = sd.pop(f"{ff}.wi.weight")
                    out[f"{dst}.mlp.dense_4h_to_h.weight"] = sd.pop(f"{ff}.wo.weight")
                else:
                    out[f"{dst}.mlp.wi_0.weight"] = sd.pop(f"{ff}.wi_0.weight")
                    out[f"{dst}.mlp.wi_1.weight"] = sd.pop(f"{ff}.wi_1.weight")
                    out[f"{dst}.mlp.wo.weight"] = sd.pop(f"{ff}.wo.weight")
            out[f"{stack}.final_layernorm.weight"] = sd.pop(f"{stack}.final_layer_norm.weight")
        return out

    def _align_prefix(self, model, state_dict):
        own = model.state_dict().keys()
        for prefix in ("mt5_model.", "t5_model."):
            if any(k.startswith(prefix) for k in own):
                return collections.OrderedDict((prefix + k, v) for k, v in state_dict.items())
        return state_dict

    def _load_config_from_json(self, config_file):
        cfg = self._read_config_json()
        self._map_config(cfg, {
            "vocab_size": "vocab_size", "d_model": "hidden_size", "num_layers": "hidden_layers", "num_heads": "num_attention_heads",
            "d_kv": "head_size", "d_ff": "intermediate_size", "relative_attention_num_buckets": "relative_attention_num_buckets",
            "layer_norm_epsilon": "layernorm_eps", "initializer_factor": "initializer_range", "eos_token_id": "eos_token_id",
            "pad_token_id": "pad_token_id", "decoder_start_token_id": "decoder_start_token_id",
            "tie_word_embeddings": "tie_word_embeddings",
        })
        # transformers >= 5 decouples the pre-head rescaling from weight tying (``scale_decoder_outputs``); older
        # config.json files only have ``tie_word_embeddings`` which implied it.  Our flag controls the rescaling
        # (whether the head is a separate matrix is decided by the checkpoint contents).
        if "scale_decoder_outputs" in cfg:
            self._update_cfg("tie_word_embeddings", bool(cfg["scale_decoder_outputs"]))
        gated = "gated" in str(cfg.get("feed_forward_proj", "relu"))
        self._update_cfg("model_type", "mt5" if gated else "t5")
        if "dropout_rate" in cfg:
            for key in ("hidden_dropout_prob", "attention_probs_dropout_prob", "embedding_dropout_prob"):
                self._update_cfg(key, cfg["dropout_rate"])


class T5LoaderLibai(ModelLoaderLiBai):
    def __init__(self, model, libai_cfg, pretrained_model_path, **kwargs):
        super().__init__(model, libai_cfg, pretrained_model_path, **kwargs)
        self.base_model_prefix_2 = "mt5_model"
