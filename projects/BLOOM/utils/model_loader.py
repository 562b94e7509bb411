"""HF BLOOM loaders (reference projects/BLOOM/utils/model_loader.py): the HF layout already matches (per-head
interleaved qkv), only the ``transformer.`` prefix and the config keys differ."""
import collections

from libai_b200.models.utils.model_loader.base_loader import ModelLoaderHuggerFace, ModelLoaderLiBai


class BlooMLoaderHuggerFace(ModelLoaderHuggerFace):
    def __init__(self, model, libai_cfg, pretrained_model_path, **kwargs):
        super().__init__(model, libai_cfg, pretrained_model_path, **kwargs)
        self.base_model_prefix_1 = "transformer"
        self.base_model_prefix_2 = "transformer"

    def _convert_state_dict(self, sd, cfg):
        return collections.OrderedDict((k, v) for k, v in sd.items() if k != "lm_head.weight")

    def _load_config_from_json(self, config_file):
        cfg = self._read_config_json()
        self._map_config(cfg, {
            "vocab_size": "vocab_size", "hidden_size": "hidden_size", "n_embed": "hidden_size", "n_layer": "hidden_layers",
            "num_hidden_layers": "hidden_layers", "n_head": "n_head", "num_attention_heads": "n_head",
            "layer_norm_epsilon": "layer_norm_epsilon", "initializer_range": "initializer_range",
            "apply_residual_connection_post_layernorm": "apply_residual_connection_post_layernorm",
            "hidden_dropout": "hidden_dropout", "attention_dropout": "attention_dropout", "pad_token_id": "padding_idx",
            "bos_token_id": "bos_token_id", "eos_token_id": "eos_token_id",
        })


class BlooMLoaderLibai(ModelLoaderLiBai):
    def __init__(self, model, libai_cfg, pretrained_model_path, **kwargs):
        super().__init__(model, libai_cfg, pretrained_model_path, **kwargs)
        self.base_model_prefix_2 = "transformer"
