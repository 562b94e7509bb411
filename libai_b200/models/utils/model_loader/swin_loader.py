"""Swin loaders (reference libai/models/utils/model_loader/swin_loader.py:22-298).  The window attention keeps the
plain ``[q; k; v]`` row order (``reshape(B, N, 3, heads, d)``), so q/k/v are concatenated without re-ordering."""
import collections

import torch

from .base_loader import ModelLoaderHuggerFace, ModelLoaderLiBai


class SwinLoaderHuggerFace(ModelLoaderHuggerFace):
    hf_prefix = "swin"

    def __init__(self, model, libai_cfg, pretrained_model_path, **kwargs):
        super().__init__(model, libai_cfg, pretrained_model_path, **kwargs)
        self.base_model_prefix_1 = self.hf_prefix
        self.base_model_prefix_2 = ""

    def _fuse_window_qkv(self, sd, base, out):
        for suffix in ("weight", "bias"):
            keys = [f"{base}.{n}.{suffix}" for n in ("query", "key", "value")]
            if all(k in sd for k in keys):
                sd[f"{out}.{suffix}"] = torch.cat([sd.pop(k) for k in keys], dim=0)

    def _block_rules(self):
        return [
            (r"^embeddings\.patch_embeddings\.projection\.", "patch_embed.proj."),
            (r"^embeddings\.norm\.", "patch_embed.norm."),
            (r"^embeddings\.position_embeddings$", "absolute_pos_embed"),
            (r"^encoder\.layers\.(\d+)\.blocks\.(\d+)\.layernorm_before\.", r"layers.\1.blocks.\2.norm1."),
            (r"^encoder\.layers\.(\d+)\.blocks\.(\d+)\.layernorm_after\.", r"layers.\1.blocks.\2.norm2."),
            (r"^encoder\.layers\.(\d+)\.blocks\.(\d+)\.attention\.self\.relative_position_bias_table$",
             r"layers.\1.blocks.\2.attn.relative_position_bias_table"),
            (r"^encoder\.layers\.(\d+)\.blocks\.(\d+)\.attention\.self\.relative_position_index$",
             r"layers.\1.blocks.\2.attn.relative_position_index"),
            (r"^encoder\.layers\.(\d+)\.blocks\.(\d+)\.attention\.output\.dense\.", r"layers.\1.blocks.\2.attn.proj."),
            (r"^encoder\.layers\.(\d+)\.blocks\.(\d+)\.intermediate\.dense\.", r"layers.\1.blocks.\2.mlp.dense_h_to_4h."),
            (r"^encoder\.layers\.(\d+)\.blocks\.(\d+)\.output\.dense\.", r"layers.\1.blocks.\2.mlp.dense_4h_to_h."),
            (r"^encoder\.layers\.(\d+)\.downsample\.", r"layers.\1.downsample."),
            (r"^layernorm\.", "norm."),
            (r"^classifier\.", "head."),
        ]

    def _convert_state_dict(self, sd, cfg):
        hp = self.hf_prefix + "."
        sd = collections.OrderedDict((k[len(hp):] if k.startswith(hp) else k, v) for k, v in sd.items())
        for li, depth in enumerate(cfg.get("depths")):
            for bi in range(depth):
                self._fuse_window_qkv(sd, f"encoder.layers.{li}.blocks.{bi}.attention.self",
                                      f"layers.{li}.blocks.{bi}.attn.qkv")
        return self._rename(sd, self._block_rules())

    def _load_config_from_json(self, config_file):
        cfg = self._read_config_json()
        self._map_config(cfg, {
            "image_size": "img_size", "patch_size": "patch_size", "num_channels": "in_chans", "embed_dim": "embed_dim",
            "depths": "depths", "num_heads": "num_heads", "window_size": "window_size", "mlp_ratio": "mlp_ratio",
            "qkv_bias": "qkv_bias", "hidden_dropout_prob": "drop_rate", "drop_path_rate": "drop_path_rate",
            "use_absolute_embeddings": "ape",
        })
        if cfg.get("id2label"):
            self._update_cfg("num_classes", len(cfg["id2label"]))


class SwinLoaderLiBai(ModelLoaderLiBai):
    def __init__(self, model, libai_cfg, pretrained_model_path, **kwargs):
        super().__init__(model, libai_cfg, pretrained_model_path, **kwargs)
        self.base_model_prefix_2 = ""
