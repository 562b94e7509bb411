"""ViT loaders (reference libai/models/utils/model_loader/vit_loader.py:22-225)."""
import collections
import re

from .base_loader import ModelLoaderHuggerFace, ModelLoaderLiBai


class ViTLoaderHuggerFace(ModelLoaderHuggerFace):
    def __init__(self, model, libai_cfg, pretrained_model_path, **kwargs):
        super().__init__(model, libai_cfg, pretrained_model_path, **kwargs)
        self.base_model_prefix_1 = "vit"
        self.base_model_prefix_2 = ""

    def _convert_state_dict(self, sd, cfg):
        sd = collections.OrderedDict((k[4:] if k.startswith("vit.") else k, v) for k, v in sd.items())
        heads, hidden, depth = cfg.get("num_heads"), cfg.get("embed_dim"), cfg.get("depth")
        for i in range(depth):
            base = f"encoder.layer.{i}.attention.attention"
            self._fuse_qkv(sd, f"{base}.query", f"{base}.key", f"{base}.value",
                           f"blocks.{i}.self_attention.query_key_value", hidden // heads, heads)
        rules = [
            (r"^embeddings\.cls_token$", "cls_token"),
            (r"^embeddings\.position_embeddings$", "pos_embed"),
            (r"^embeddings\.patch_embeddings\.projection\.", "patch_embed.proj."),
            (r"^encoder\.layer\.(\d+)\.attention\.output\.dense\.", r"blocks.\1.self_attention.dense."),
            (r"^encoder\.layer\.(\d+)\.layernorm_before\.", r"blocks.\1.input_layernorm."),
            (r"^encoder\.layer\.(\d+)\.layernorm_after\.", r"blocks.\1.post_attention_layernorm."),
            (r"^encoder\.layer\.(\d+)\.intermediate\.dense\.", r"blocks.\1.mlp.dense_h_to_4h."),
            (r"^encoder\.layer\.(\d+)\.output\.dense\.", r"blocks.\1.mlp.dense_4h_to_h."),
            (r"^layernorm\.", "norm."),
            (r"^classifier\.", "head."),
        ]
        out = self._rename(sd, rules)
        return collections.OrderedDict((k, v) for k, v in out.items() if not k.startswith("pooler."))

    def _load_config_from_json(self, config_file):
        cfg = self._read_config_json()
        self._map_config(cfg, {
            "image_size": "img_size", "patch_size": "patch_size", "num_channels": "in_chans",
            "hidden_size": "embed_dim", "num_hidden_layers": "depth", "num_attention_heads": "num_heads",
            "attention_probs_dropout_prob": "attn_drop_rate", "hidden_dropout_prob": "drop_rate",
        })
        if "intermediate_size" in cfg:
            self._update_cfg("mlp_ratio", int(cfg["intermediate_size"] / cfg["hidden_size"]))
        if cfg.get("id2label"):
            self._update_cfg("num_classes", len(cfg["id2label"]))


class ViTLoaderLiBai(ModelLoaderLiBai):
    def __init__(self, model, libai_cfg, pretrained_model_path, **kwargs):
        super().__init__(model, libai_cfg, pretrained_model_path, **kwargs)
        self.base_model_prefix_2 = ""
