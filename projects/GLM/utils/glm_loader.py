"""HF GLM → ``GLMForConditionalGeneration`` (reference projects/GLM/utils/glm_loader.py): ``[q; k; v]`` → per-head
interleaved rows, ``word_embeddings`` moved under ``embeddings``."""
import collections
import re

from libai_b200.models.utils.model_loader.base_loader import ModelLoaderHuggerFace, ModelLoaderLiBai


class GLMLoaderHuggerFace(ModelLoaderHuggerFace):
    def __init__(self, model, libai_cfg, pretrained_model_path, **kwargs):
        super().__init__(model, libai_cfg, pretrained_model_path, **kwargs)
        self.base_model_prefix_1 = "glm"
        self.base_model_prefix_2 = "glm"

    def _convert_state_dict(self, sd, cfg):
        heads, hidden = cfg.get("num_attention_heads"), cfg.get("hidden_size")
        out = collections.OrderedDict()
        for k, v in sd.items():
            k = k[4:] if k.startswith("glm.") else k
            if k.startswith("word_embeddings."):
                k = "embeddings." + k
            elif k.startswith("transformer.position_embeddings.") or k.startswith("transformer.block_position_embeddings."):
                k = "embeddings." + k[len("transformer."):]
            if re.search(r"attention\.query_key_value\.(weight|bias)$", k):
                v = self._fix_qkv_ordering(v, hidden // heads, heads)
            out["glm." + k] = v
        return out

    def _load_config_from_json(self, config_file):
        cfg = self._read_config_json()
        self._map_config(cfg, {
            "num_layers": "num_layers", "vocab_size": "vocab_size", "hidden_size": "hidden_size",
            "num_attention_heads": "num_attention_heads", "max_sequence_length": "max_sequence_length",
            "embedding_dropout_prob": "embedding_dropout_prob", "attention_dropout_prob": "attention_dropout_prob",
            "output_dropout_prob": "output_dropout_prob", "block_position_encoding": "block_position_encoding",
            "attention_scale": "attention_scale",
        })


class GLMLoaderLiBai(ModelLoaderLiBai):
    def __init__(self, model, libai_cfg, pretrained_model_path, **kwargs):
        super().__init__(model, libai_cfg, pretrained_model_path, **kwargs)
        self.base_model_prefix_2 = "glm"
