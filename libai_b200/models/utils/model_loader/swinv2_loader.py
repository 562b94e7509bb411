"""SwinV2 loaders (reference libai/models/utils/model_loader/swinv2_loader.py:22-316): q/v biases are separate
parameters (k has none), plus ``logit_scale`` and the continuous position bias MLP."""
import collections

import torch

from .base_loader import ModelLoaderLiBai
from .swin_loader import SwinLoaderHuggerFace


class SwinV2LoaderHuggerFace(SwinLoaderHuggerFace):
    hf_prefix = "swinv2"

    def _fuse_window_qkv(self, sd, base, out):
        keys = [f"{base}.{n}.weight" for n in ("query", "key", "value")]
        if all(k in sd for k in keys):
            sd[f"{out}.weight"] = torch.cat([sd.pop(k) for k in keys], dim=0)
        attn = out.rsplit(".", 1)[0]
        if f"{base}.query.bias" in sd:
            sd[f"{attn}.q_bias"] = sd.pop(f"{base}.query.bias")
        if f"{base}.value.bias" in sd:
            sd[f"{attn}.v_bias"] = sd.pop(f"{base}.value.bias")
        sd.pop(f"{base}.key.bias", None)

    def _block_rules(self):
        extra = [
            (r"^encoder\.layers\.(\d+)\.blocks\.(\d+)\.attention\.self\.logit_scale$", r"layers.\1.blocks.\2.attn.logit_scale"),
            (r"^encoder\.layers\.(\d+)\.blocks\.(\d+)\.attention\.self\.continuous_position_bias_mlp\.",
             r"layers.\1.blocks.\2.attn.cpb_mlp."),
            (r"^encoder\.layers\.(\d+)\.blocks\.(\d+)\.attention\.self\.relative_coords_table$",
             r"layers.\1.blocks.\2.attn.relative_coords_table"),
        ]
        return extra + super()._block_rules()

    def _convert_state_dict(self, sd, cfg):
        out = super()._convert_state_dict(sd, cfg)
        # non-persistent buffers of the model (recomputed from the window size)
        fixed = collections.OrderedDict()
        for k, v in out.items():
            if k.endswith(("relative_coords_table", "relative_position_index")):
                continue
            if k.endswith("attn.logit_scale") and v.dim() == 3:  # HF [heads, 1, 1] → [1, heads, 1, 1]
                v = v.unsqueeze(0)
            fixed[k] = v
        return fixed

    def _load_config_from_json(self, config_file):
        super()._load_config_from_json(config_file)
        cfg = self._read_config_json()
        self._map_config(cfg, {"pretrained_window_sizes": "pretrained_window_sizes"})


class SwinV2LoaderLiBai(ModelLoaderLiBai):
    def __init__(self, model, libai_cfg, pretrained_model_path, **kwargs):
        super().__init__(model, libai_cfg, pretrained_model_path, **kwargs)
        self.base_model_prefix_2 = ""
