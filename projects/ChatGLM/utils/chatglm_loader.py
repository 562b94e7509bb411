"""ChatGLM2/3 checkpoint loaders (reference projects/ChatGLM/utils/chatglm_loader.py): the parameter names are the
original ones, only the rotary ``inv_freq`` buffer is dropped; ``config.json`` keys map 1:1."""
import collections

from libai_b200.models.utils.model_loader.base_loader import ModelLoaderHuggerFace, ModelLoaderLiBai


class ChatGLMLoaderHuggerFace(ModelLoaderHuggerFace):
    def __init__(self, model, libai_cfg, pretrained_model_path, **kwargs):
        super().__init__(model, libai_cfg, pretrained_model_path, **kwargs)
        self.base_model_prefix_1 = "transformer"
        self.base_model_prefix_2 = "transformer"

    def _convert_state_dict(self, sd, cfg):
        return collections.OrderedDict((k, v) for k, v in sd.items() if not k.endswith("rotary_pos_emb.inv_freq"))

    def _load_config_from_json(self, config_file):
        cfg = self._read_config_json()
        keys = ("add_bias_linear add_qkv_bias apply_query_key_layer_scaling apply_residual_connection_post_layernorm "
                "attention_dropout attention_softmax_in_fp32 ffn_hidden_size fp32_residual_connection hidden_dropout "
                "hidden_size kv_channels layernorm_epsilon multi_query_attention multi_query_group_num num_attention_heads "
                "num_layers padded_vocab_size post_layer_norm rmsnorm seq_length tie_word_embeddings eos_token_id "
                "pad_token_id").split()
        self._map_config(cfg, {k: k for k in keys})


class ChatGLMLoaderLiBai(ModelLoaderLiBai):
    def __init__(self, model, libai_cfg, pretrained_model_path, **kwargs):
        super().__init__(model, libai_cfg, pretrained_model_path, **kwargs)
        self.base_model_prefix_2 = "transformer"


class _LoraMixin:
    """Load the dense checkpoint, wrap the transformer in a :class:`LoraModel`, then (optionally) load adapter weights
    saved with ``LoraModel.lora_state_dict()`` (reference projects/ChatGLM/utils/chatglm_loader.py:77-150).

    Extra keyword arguments: ``lora_cfg`` (required) and ``lora_pretrained_model_path`` (a ``torch.save``d adapter
    state dict, or ``None`` for freshly initialised adapters)."""

    def _init_lora(self, kwargs):
        self.lora_cfg = kwargs.pop("lora_cfg")
        self.lora_pretrained_model_path = kwargs.pop("lora_pretrained_model_path", None)

    def load(self):
        import torch

        from projects.ChatGLM.lora.lora_model import LoraModel

        want_info = self.output_loading_info
        loaded = super().load()
        model, info = loaded if want_info else (loaded, None)
        model.transformer = LoraModel(model.transformer, self.lora_cfg, adapter_name="default")
        lora_info = {"missing_keys": [], "unexpected_keys": [], "mismatched_keys": [], "error_msgs": []}
        if self.lora_pretrained_model_path is not None:
            sd = torch.load(self.lora_pretrained_model_path, map_location="cpu", weights_only=True)
            own = model.transformer.state_dict()
            for k, v in sd.items():
                if k not in own:
                    lora_info["unexpected_keys"].append(k)
                elif tuple(own[k].shape) != tuple(v.shape):
                    lora_info["mismatched_keys"].append(k)
                else:
                    own[k].copy_(v.to(own[k].dtype))
            lora_info["missing_keys"] = [k for k in own if "lora_" in k and k not in sd]
        if want_info:
            return model, {k: list(info.get(k, [])) + lora_info[k] for k in lora_info}
        return model


class ChatGLMLoraLoaderHuggerFace(_LoraMixin, ChatGLMLoaderHuggerFace):
    def __init__(self, model, libai_cfg, pretrained_model_path, **kwargs):
        self._init_lora(kwargs)
        super().__init__(model, libai_cfg, pretrained_model_path, **kwargs)


class ChatGLMLoraLoaderLiBai(_LoraMixin, ChatGLMLoaderLiBai):
    def __init__(self, model, libai_cfg, pretrained_model_path, **kwargs):
        self._init_lora(kwargs)
        super().__init__(model, libai_cfg, pretrained_model_path, **kwargs)
