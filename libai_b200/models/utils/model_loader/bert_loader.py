"""BERT loaders (reference libai/models/utils/model_loader/bert_loader.py:22-263).

HF BERT is post-LN; the LiBai block is ``LN → attn → +res → LN → MLP → +res`` with
``apply_residual_post_layernorm=True``, so the LayerNorms shift by one position:
``embeddings.LayerNorm → encoders.0.input_layernorm``, layer *i* ``attention.output.LayerNorm →
encoders.i.post_attention_layernorm``, layer *i* ``output.LayerNorm → encoders.(i+1).input_layernorm``
(last one → ``final_layernorm``).
"""
import collections
import re

from .base_loader import ModelLoaderHuggerFace, ModelLoaderLiBai


class BertLoaderHuggerFace(ModelLoaderHuggerFace):
    hf_prefix = "bert"

    def __init__(self, model, libai_cfg, pretrained_model_path, **kwargs):
        super().__init__(model, libai_cfg, pretrained_model_path, **kwargs)
        self.base_model_prefix_1 = self.hf_prefix
        self.base_model_prefix_2 = "bert"

    def _convert_state_dict(self, sd, cfg):
        sd = collections.OrderedDict(sd)
        heads, hidden, layers = cfg.get("num_attention_heads"), cfg.get("hidden_size"), cfg.get("hidden_layers")
        head_size = hidden // heads
        hp = self.hf_prefix + "."
        has_prefix = any(k.startswith(hp) for k in sd)
        src = hp if has_prefix else ""
        dst = "bert." if has_prefix else ""
        for i in range(layers):
            base = f"{src}encoder.layer.{i}.attention.self"
            self._fuse_qkv(sd, f"{base}.query", f"{base}.key", f"{base}.value",
                           f"{dst}encoders.{i}.self_attention.query_key_value", head_size, heads)
        e, d = re.escape(src), dst
        rules = [
            (rf"^{e}embeddings\.word_embeddings\.", f"{d}embeddings.vocab_embeddings."),
            (rf"^{e}embeddings\.token_type_embeddings\.", f"{d}embeddings.tokentype_embeddings."),
            (rf"^{e}embeddings\.position_embeddings\.", f"{d}embeddings.position_embeddings."),
            (rf"^{e}embeddings\.LayerNorm\.", f"{d}encoders.0.input_layernorm."),
            (rf"^{e}encoder\.layer\.(\d+)\.attention\.output\.dense\.", rf"{d}encoders.\1.self_attention.dense."),
            (rf"^{e}encoder\.layer\.(\d+)\.attention\.output\.LayerNorm\.", rf"{d}encoders.\1.post_attention_layernorm."),
            (rf"^{e}encoder\.layer\.(\d+)\.intermediate\.dense\.", rf"{d}encoders.\1.mlp.dense_h_to_4h."),
            (rf"^{e}encoder\.layer\.(\d+)\.output\.dense\.", rf"{d}encoders.\1.mlp.dense_4h_to_h."),
            (rf"^{e}pooler\.dense\.", f"{d}pooler.dense."),
            (r"^cls\.predictions\.transform\.dense\.", "cls_head.predictions.dense."),
            (r"^cls\.predictions\.transform\.LayerNorm\.", "cls_head.predictions.layernorm."),
            (r"^cls\.predictions\.bias$", "cls_head.lm_logits.bias"),
            (r"^cls\.predictions\.decoder\.bias$", "cls_head.lm_logits.bias"),
            (r"^cls\.seq_relationship\.", "cls_head.seq_relationship."),
            (r"^lm_head\.dense\.", "lm_head.dense."),
            (r"^lm_head\.layer_norm\.", "lm_head.layernorm."),
            (r"^lm_head\.bias$", "lm_head.lm_logits.bias"),
            (r"^lm_head\.decoder\.bias$", "lm_head.lm_logits.bias"),
        ]
        sd = self._rename(sd, rules)
        out = collections.OrderedDict()
        pat = re.compile(rf"^{e}encoder\.layer\.(\d+)\.output\.LayerNorm\.(weight|bias)$")
        for k, v in sd.items():
            m = pat.match(k)
            if m:
                i = int(m.group(1))
                k = f"{d}final_layernorm.{m.group(2)}" if i == layers - 1 else f"{d}encoders.{i + 1}.input_layernorm.{m.group(2)}"
            if k.endswith("position_ids") or k in ("cls.predictions.decoder.weight", "lm_head.decoder.weight"):
                continue  # buffers / tied copies
            out[k] = v
        return out

    def _load_config_from_json(self, config_file):
        cfg = self._read_config_json()
        self._map_config(cfg, {
            "vocab_size": "vocab_size", "hidden_size": "hidden_size", "num_hidden_layers": "hidden_layers",
            "num_attention_heads": "num_attention_heads", "intermediate_size": "intermediate_size",
            "hidden_dropout_prob": "hidden_dropout_prob", "attention_probs_dropout_prob": "attention_probs_dropout_prob",
            "max_position_embeddings": "max_position_embeddings", "type_vocab_size": "num_tokentypes",
            "initializer_range": "initializer_range", "layer_norm_eps": "layernorm_eps",
        })
        self._update_cfg("apply_residual_post_layernorm", True)  # original BERT residual ordering


class BertLoaderLiBai(ModelLoaderLiBai):
    def __init__(self, model, libai_cfg, pretrained_model_path, **kwargs):
        super().__init__(model, libai_cfg, pretrained_model_path, **kwargs)
        self.base_model_prefix_2 = "bert"
