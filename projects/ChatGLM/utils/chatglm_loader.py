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
